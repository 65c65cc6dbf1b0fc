"""Device-resident stage times of the HALF C-simulation mode (binary16 buffers, FADD_LATENCY 4, bit-exact) on a
block-diagonal Cora-shape batch.  usage: python tools/half_bench.py [copies]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgracex1_b200 import _lib, graphs as G  # noqa: E402
from sgracex1_b200.driver import DeviceLayer  # noqa: E402
from sgracex1_b200.pynq_compat import Overlay  # noqa: E402

copies = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ip = Overlay("gnn_all.bit", device=0).mmult_top_0
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ip.handle.set_stream(stream.cuda_stream)
b = G.block_diagonal([G.cora_shape(seed=s) for s in range(min(16, copies))], copies)
ip.configure(mode=_lib.MODE_F16_CSIM, staging=0, index_format=0)
dl = DeviceLayer(ip.handle, _lib.MODE_F16_CSIM, device="cuda:0")
to16 = lambda a: np.asarray(a, np.float32).astype(np.float16).view(np.uint16)
dl.load(N=b.N, M=b.M, P=b.P, adj=(b.adj_rowptr, b.adj_col, to16(b.adj_val)), fea=(b.fea_rowptr, b.fea_col, to16(b.fea_val)),
        B=to16(b.B), relu=1)
d, xw = dl.desc, dl.t["XW"].data_ptr()
for _ in range(3):
    ip.handle.fea_run(d, xw); ip.handle.adj_run(d, xw, b.N)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf, ta = [], []
for _ in range(10):
    e[0].record(); ip.handle.fea_run(d, xw); e[1].record(); ip.handle.adj_run(d, xw, b.N); e[2].record()
    torch.cuda.synchronize()
    tf.append(e[0].elapsed_time(e[1])); ta.append(e[1].elapsed_time(e[2]))
ab = b.algorithmic_bytes(elt=2)
f, a = float(np.median(tf)), float(np.median(ta))
print(f"HALF C-sim cora_x{copies}: N={b.N}  FEA {f:.4f} ms {ab['fea'] / f / 1e6:.0f} GB/s   ADJ {a:.4f} ms {ab['adj'] / a / 1e6:.0f} GB/s  "
      f"{b.nnz_adj / a / 1e6:.1f} GTEPS")
