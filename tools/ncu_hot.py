"""Top stalled SASS lines of each kernel in an .ncu-rep (source page, needs -lineinfo).
usage: python tools/ncu_hot.py rep.ncu-rep [kernel-regex] [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "."
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
for b in blocks:
    h = b["hdr"]
    ci = {k: i for i, k in enumerate(h)}
    tot = sum(int(r[ci["# Samples"]] or 0) for r in b["rows"])
    texe = sum(int(r[ci["Instructions Executed"]] or 0) for r in b["rows"])
    print("=" * 110)
    print(b["name"], "| samples", tot, "| warp insts", texe)
    stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    rows = sorted(b["rows"], key=lambda r: -int(r[ci["# Samples"]] or 0))[:N]
    for r in rows:
        st = sorted(((int(r[ci[k]] or 0), k) for k in stall_cols), reverse=True)[:2]
        print(f'{r[ci["Address"]][-5:]:>6} {int(r[ci["# Samples"]] or 0):6d} {100.0*int(r[ci["# Samples"]] or 0)/max(tot,1):5.1f}% '
              f'exec {int(r[ci["Instructions Executed"]] or 0):9d} wf {r[ci["L1 Wavefronts Shared"]]:>9}/{r[ci["L1 Wavefronts Shared Ideal"]]:>9} '
              f'{st[0][1][6:]}:{st[0][0]} {st[1][1][6:]}:{st[1][0]} | {r[ci["Source"]][:70]}')
