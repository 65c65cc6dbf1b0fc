"""Graph preparation: host torch (the reference's code path) against the GPU entry points."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
from sgracex1_b200 import sgrace as S

def timeit(fn, reps, sync=False):
    fn()
    if sync: torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    if sync: torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

rng = np.random.default_rng(0)
for name, n, e in (("cora", 2708, 10556), ("pubmed", 19717, 88648), ("products/8", 306129, 7_700_000)):
    ei = torch.from_numpy(rng.integers(0, n, size=(2, e)).astype(np.int64))
    eid = ei.cuda()
    host = timeit(lambda: S.sym_norm2(ei, n, fill=1.0), 3 if e > 1e6 else 20)
    dev = timeit(lambda: S.sym_norm2_device(eid, n, fill=1.0), 20, sync=True)
    print(json.dumps({"op": "sym_norm2", "graph": name, "edges": e, "host_ms": round(host, 3), "gpu_ms": round(dev, 3)}), flush=True)
for name, n, m, dens in (("cora X", 2708, 1433, 0.0127), ("citeseer X", 3327, 3703, 0.0085)):
    x = (rng.random((n, m)) < dens).astype(np.float32)
    xt = torch.from_numpy(x); xd = xt.cuda()
    host = timeit(lambda: xt.to_sparse(), 10)
    dev = timeit(lambda: S.to_sparse_device(xd), 20, sync=True)
    print(json.dumps({"op": "to_sparse", "matrix": name, "host_ms": round(host, 3), "gpu_ms": round(dev, 3)}), flush=True)
