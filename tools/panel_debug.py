"""Which rows differ between the panel and the gather ADJ kernels (debug aid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgracex1_b200 import _lib
from sgracex1_b200.pynq_compat import MmultTop
from tests.test_gpu_panel import batched_adjacency, adj_stage
ip = MmultTop(0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(P)
sizes = rng.integers(40, 420 if P <= 64 else 300, size=400)
N, adj = batched_adjacency(sizes, rng, hub_every=37)
xw = rng.standard_normal((N, P)).astype(np.float32)
deg = np.diff(adj[0])
for relu in (1, 0):
    for long_row in (512, 64):
        want, n0, _ = adj_stage(ip, adj, xw, N, P, relu, plan=0, long_row=long_row)
        got, n1, _ = adj_stage(ip, adj, xw, N, P, relu, plan=1, long_row=long_row)
        bad = np.where((got != want).any(axis=1))[0]
        print(f"relu={relu} long_row={long_row} panel launches {n1}: {len(bad)} rows differ; degs {deg[bad][:12]} rows {bad[:12]}",
              "maxerr", float(np.abs(got - want).max()))
        if len(bad):
            r = bad[0]; print(" got", got[r][:4], "want", want[r][:4])
