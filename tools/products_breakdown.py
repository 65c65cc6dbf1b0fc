"""Per-stage timing of the ogbn-products-shape layer on one GPU (tuning aid)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from sgracex1_b200 import _lib, dist as sdist, graphs as G

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
N = int(2_449_029 * scale)
M, P = 100, 256
dev = torch.device("cuda:0")
rp, ci, va = G.products_shape_rows(0, N, n_total=N)
deg = np.diff(rp)
print(f"N={N} nnz={len(ci)} max_deg={deg.max()} rows>512: {(deg > 512).sum()} nnz in them {deg[deg > 512].sum()}")
x = torch.randn(N, M, device=dev)
W = (torch.rand(M, P, device=dev) - 0.5) * 0.2
adj = tuple(torch.from_numpy(a).to(dev) for a in (rp, ci, va))
h = _lib.Handle(0)
h.set_option(_lib.OPT_STAGING, 0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
h.set_stream(st.cuda_stream)
fea, adjf = sdist.abi_fea_fn(h), sdist.abi_adj_fn(h)
xw = torch.empty(N, P, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(4):
    ev[0].record()
    keep = fea(x, W, xw)
    ev[1].record()
    D = adjf(adj, xw, 1)
    ev[2].record()
    torch.cuda.synchronize()
    tf, ta = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
fb = N * M * 4 + M * P * 4 + N * P * 4
ab = (N + 1) * 4 + len(ci) * 8 + 2 * N * P * 4
print(f"FEA dense {tf:.3f} ms  {fb / tf / 1e6:.0f} GB/s  {2.0 * N * M * P / tf / 1e9:.1f} TFLOP/s")
print(f"ADJ       {ta:.3f} ms  {ab / ta / 1e6:.0f} GB/s algorithmic, gather {len(ci) * P * 4 / ta / 1e6:.0f} GB/s, {len(ci) / ta / 1e6:.2f} GTEPS")
ref = torch.relu(torch.sparse_csr_tensor(adj[0].long(), adj[1].long(), adj[2], size=(N, N)) @ (x @ W))
err = (D - ref).abs().max().item() / ref.abs().max().item()
print("max rel err vs torch", err)

# aggregate-first order through the layer entry point
from sgracex1_b200.driver import DeviceLayer  # noqa: E402
h.set_option(_lib.OPT_AGG_FIRST, 1)
d = _lib.LayerDesc()
d.gemm_mode, d.relu, d.N_adj, d.M_adj, d.M_fea, d.P_w = 1, 1, N, N, M, P
Bt = W.t().contiguous()
D2 = torch.empty(N, P, device=dev)
d.values_fea, d.B, d.D = x.data_ptr(), Bt.data_ptr(), D2.data_ptr()
d.rowPtr_adj, d.columnIndex_adj, d.values_adj, d.nnz_adj = adj[0].data_ptr(), adj[1].data_ptr(), adj[2].data_ptr(), len(ci)
for it in range(4):
    ev[0].record()
    h.layer_run(d)
    ev[1].record()
    torch.cuda.synchronize()
print(f"aggregate-first layer {ev[0].elapsed_time(ev[1]):.3f} ms; max rel err vs torch {(D2 - ref).abs().max().item() / ref.abs().max().item():.2e}")
