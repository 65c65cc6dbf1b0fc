"""Panel (shared-memory window) ADJ against the gather kernel on batches of SMALL graphs, where the window leaves room for
large CSR rings: half-Cora graphs (1354 nodes) x2048 at P = 16 and a molecule batch at P = 64.
usage: [SGRACE_PANEL_G=4 ...] python tools/panel_small_graphs.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgracex1_b200 import _lib, graphs as G
from sgracex1_b200.pynq_compat import Overlay
from tests.test_gpu_panel import adj_stage

ip = Overlay("gnn_all.bit", device=0).mmult_top_0
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ip.handle.set_stream(stream.cuda_stream)


def time_adj(adj, N, P, plan, reps=10):
    rng = np.random.default_rng(0)
    xw = rng.standard_normal((N, P)).astype(np.float32)
    D, n, (t, x, d, Dt) = adj_stage(ip, adj, xw, N, P, 1, plan=plan)
    ip.configure(mode=_lib.MODE_F32_FAST, staging=0, index_format=0, adj_plan=plan)
    for _ in range(3):
        ip.handle.adj_run(d, x.data_ptr(), N)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ts = []
    for _ in range(reps):
        e[0].record(); ip.handle.adj_run(d, x.data_ptr(), N); e[1].record(); torch.cuda.synchronize(); ts.append(e[0].elapsed_time(e[1]))
    ip.configure(adj_plan=0)
    return float(np.median(ts)), n, D


for name, probs, copies in (("half-cora x2048 P=16", [G.cora_shape(seed=s, n=1354, nnz_adj=6632, nnz_fea=24608) for s in range(8)], 2048),
                            ("quarter-cora x4096 P=16", [G.cora_shape(seed=s, n=677, nnz_adj=3316, nnz_fea=12304) for s in range(8)], 4096)):
    b = G.block_diagonal(probs, copies)
    adj = (b.adj_rowptr, b.adj_col, b.adj_val)
    t0, _, D0 = time_adj(adj, b.N, 16, 0)
    t1, n1, D1 = time_adj(adj, b.N, 16, 1)
    byt = (b.N + 1) * 4 + b.nnz_adj * 8 + 2 * b.N * 16 * 4
    print(f"{name}: N={b.N} nnz={b.nnz_adj}  gather {t0:.4f} ms ({byt / t0 / 1e6:.0f} GB/s)  panel {t1:.4f} ms ({byt / t1 / 1e6:.0f} GB/s) "
          f"[{'panel kernel' if n1 else 'fell back'}] {'bit-equal' if np.array_equal(D0, D1) else 'DIFFER'}")
mb, _, _ = G.molecule_batch(n_graphs=60000, seed=1, P=64)
adj = (mb.adj_rowptr, mb.adj_col, mb.adj_val)
t0, _, D0 = time_adj(adj, mb.N, 64, 0)
t1, n1, D1 = time_adj(adj, mb.N, 64, 1)
print(f"molecules x60000 P=64: N={mb.N} nnz={len(adj[1])}  gather {t0:.4f} ms  panel {t1:.4f} ms [{'panel kernel' if n1 else 'fell back'}] "
      f"{'bit-equal' if np.array_equal(D0, D1) else 'DIFFER'}")
