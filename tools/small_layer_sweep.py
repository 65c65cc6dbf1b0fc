"""Latency of one small layer (device-resident, back-to-back launches) for the launch strategies.
Each configuration runs in its own process because the tunables are read once per process."""
import os, subprocess, sys, json
sys.path.insert(0, ".")

def one(copies, P, fused):
    import numpy as np, torch
    import bench
    from sgracex1_b200 import _lib
    from sgracex1_b200.driver import DeviceLayer
    from sgracex1_b200.pynq_compat import MmultTop
    ip = MmultTop(0); ip.configure(mode=_lib.MODE_F32_FAST, staging=0, fused_small=fused)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ip.handle.set_stream(stream.cuda_stream)
    batch, _ = bench.make_cora_batch(copies, seed0=0, P=P)
    dl = DeviceLayer(ip.handle, _lib.MODE_F32_FAST)
    dl.load(N=batch.N, M=batch.M, P=batch.P, adj=(batch.adj_rowptr, batch.adj_col, batch.adj_val),
            fea=(batch.fea_rowptr, batch.fea_col, batch.fea_val), B=batch.B, relu=1)
    for _ in range(20): dl.run(sync=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ip.handle.launch_count()
    e0.record()
    for _ in range(200): dl.run(sync=False)
    e1.record(); torch.cuda.synchronize()
    D = dl.result("D")
    print(json.dumps({"us": round(e0.elapsed_time(e1) * 5, 2), "launches": (ip.handle.launch_count() - l0) / 200,
                      "sum": float(np.abs(D).sum())}))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])); sys.exit(0)
    cfgs = [("unfused", 0, {}), ("plain", 1 << 20, {"SGRACE_FUSED_SPLIT_ROWS": "0"}),
            ("split", 1 << 20, {"SGRACE_FUSED_SPLIT_ROWS": "10000000"}),
            ("split ph1", 1 << 20, {"SGRACE_FUSED_SPLIT_ROWS": "10000000", "SGRACE_FUSED_PHASES": "1"}),
            ("split ph3", 1 << 20, {"SGRACE_FUSED_SPLIT_ROWS": "10000000", "SGRACE_FUSED_PHASES": "3"}),
            ("split ph0", 1 << 20, {"SGRACE_FUSED_SPLIT_ROWS": "10000000", "SGRACE_FUSED_PHASES": "0"}),
            ("split 8/sm", 1 << 20, {"SGRACE_FUSED_SPLIT_ROWS": "10000000", "SGRACE_FUSED_CTAS_PER_SM": "8"})]
    for copies, P in ((1, 16), (1, 64), (4, 16), (16, 16), (64, 16)):
        for name, fused, env in cfgs:
            if copies > 4 and "ph" in name: continue
            r = subprocess.run([sys.executable, __file__, str(copies), str(P), str(fused)], env={**os.environ, **env},
                               capture_output=True, text=True)
            print(f"copies {copies:3d} P {P:3d} {name:12s}", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
