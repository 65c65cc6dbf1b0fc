"""ADJ panel (shared-memory window) kernel vs the gather kernel on a block-diagonal Cora-shape batch:
bit-equality of D and the stage times of both.  usage: python tools/adj_panel_check.py [copies] [hidden]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgracex1_b200 import _lib, graphs as G  # noqa: E402
from sgracex1_b200.driver import DeviceLayer  # noqa: E402
from sgracex1_b200.pynq_compat import Overlay  # noqa: E402

copies = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
P = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ip = Overlay("gnn_all.bit", device=0).mmult_top_0
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ip.handle.set_stream(stream.cuda_stream)
probs = [G.cora_shape(seed=s, P=P) for s in range(min(16, copies))]
b = G.block_diagonal(probs, copies)
ip.configure(mode=_lib.MODE_F32_FAST, staging=0, index_format=0)
dl = DeviceLayer(ip.handle, _lib.MODE_F32_FAST, device="cuda:0")
dl.load(N=b.N, M=b.M, P=b.P, adj=(b.adj_rowptr, b.adj_col, b.adj_val), fea=(b.fea_rowptr, b.fea_col, b.fea_val), B=b.B, relu=1)
d, xw = dl.desc, dl.t["XW"].data_ptr()
adj_bytes = (b.N + 1) * 4 + b.nnz_adj * 8 + 2 * b.N * b.P * 4


def run(plan, reps=10):
    ip.configure(adj_plan=plan)
    dl.t["D"].zero_()
    for _ in range(3):
        ip.handle.fea_run(d, xw)
        ip.handle.adj_run(d, xw, b.N)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf, ta = [], []
    for _ in range(reps):
        e[0].record(); ip.handle.fea_run(d, xw); e[1].record(); ip.handle.adj_run(d, xw, b.N); e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1])); ta.append(e[1].elapsed_time(e[2]))
    return float(np.median(tf)), float(np.median(ta)), min(ta), dl.t["D"].clone()


f0, a0, a0m, D0 = run(0)
n0 = ip.handle.get_option(_lib.OPT_PANEL_LAUNCHES)
f1, a1, a1m, D1 = run(1)
n1 = ip.handle.get_option(_lib.OPT_PANEL_LAUNCHES)
print(f"N={b.N} nnz_adj={b.nnz_adj} P={b.P} plan builds={ip.handle.get_option(_lib.OPT_PLAN_BUILDS)} panel launches={n1 - n0}")
print(f"gather: FEA {f0:.4f} ms  ADJ {a0:.4f} ms (min {a0m:.4f})  {adj_bytes / a0 / 1e6:.0f} GB/s")
print(f"panel : FEA {f1:.4f} ms  ADJ {a1:.4f} ms (min {a1m:.4f})  {adj_bytes / a1 / 1e6:.0f} GB/s  [{'panel kernel' if n1 > n0 else 'FELL BACK to gather'}]")
print("bit-equal" if torch.equal(D0, D1) else f"DIFFER max {float((D0 - D1).abs().max())}")
