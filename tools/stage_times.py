"""Per-stage CUDA-event times of the cora_x1024 layer under a list of SGRACE_STREAM_* settings.
usage: python tools/stage_times.py "A=1 B=2" "C=3" ...   (each argument = one setting; "" = defaults)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sgracex1_b200 import _lib  # noqa: E402
from sgracex1_b200.driver import DeviceLayer  # noqa: E402
from sgracex1_b200.pynq_compat import MmultTop  # noqa: E402

os.environ["SGRACE_TUNE_LIVE"] = "1"      # the library re-reads SGRACE_STREAM_* at every launch
copies = int(os.environ.get("COPIES", "1024"))
batch, probs = bench.make_cora_batch(copies, seed0=0)
ip = MmultTop(0)
ip.configure(mode=_lib.MODE_F32_FAST, index_format=0, staging=0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ip.handle.set_stream(stream.cuda_stream)
dl = DeviceLayer(ip.handle, _lib.MODE_F32_FAST, device="cuda:0")
dl.load(N=batch.N, M=batch.M, P=batch.P, adj=(batch.adj_rowptr, batch.adj_col, batch.adj_val),
        fea=(batch.fea_rowptr, batch.fea_col, batch.fea_val), B=batch.B, relu=1)
d, xw = dl.desc, dl.t["XW"].data_ptr()
dl.run(sync=True)
ref = dl.result("D").copy()
settings = sys.argv[1:] or [""]
for s in settings:
    keys = []
    for kv in s.split():
        k, v = kv.split("=")
        os.environ[k] = v
        keys.append(k)
    try:
        for _ in range(3):
            dl.run(sync=False)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        f = a = 0.0
        reps = 10
        for _ in range(reps):
            e[0].record(); ip.handle.fea_run(d, xw); e[1].record(); ip.handle.adj_run(d, xw, batch.N); e[2].record()
            torch.cuda.synchronize()
            f += e[0].elapsed_time(e[1]) / reps
            a += e[1].elapsed_time(e[2]) / reps
        ok = np.array_equal(dl.result("D"), ref)
        print(f"{s or 'defaults':60s} fea {f:.4f} ms  adj {a:.4f} ms  layer {f + a:.4f}  same={ok}", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"{s:60s} FAILED {ex}", flush=True)
    for k in keys:
        os.environ.pop(k, None)
