"""One-line summary of bench.py's JSON line (tuning aid)."""
import json
import sys

for ln in sys.stdin:
    ln = ln.strip()
    if ln.startswith("{"):
        d = json.loads(ln)
        s = d.get("stages", {})
        print("ms_step", round(d["ms_per_step"], 4),
              "| fea", round(s["fea"]["ms"], 4), "ms", round(s["fea"]["gbs"]), "GB/s",
              "| adj", round(s["adj"]["ms"], 4), "ms", round(s["adj"]["gbs"]), "GB/s",
              "| gteps", round(d["value"], 2), "| e2e", round(d["e2e"]["value"], 3), d["e2e"].get("matches_resident"),
              "| launches", d.get("gpu_launches"))
    elif ln:
        print(ln)
