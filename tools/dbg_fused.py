import sys; sys.path.insert(0, ".")
import numpy as np, torch, time
from sgracex1_b200 import _lib, graphs as G
from sgracex1_b200.driver import DeviceLayer
from sgracex1_b200.pynq_compat import MmultTop
ip = MmultTop(0); ip.configure(mode=_lib.MODE_F32_FAST, staging=0)
p = G.cora_shape(seed=0)
dl = DeviceLayer(ip.handle, _lib.MODE_F32_FAST)
dl.load(N=p.N, M=p.M, P=p.P, adj=(p.adj_rowptr, p.adj_col, p.adj_val), fea=(p.fea_rowptr, p.fea_col, p.fea_val), B=p.B, relu=1)
for fs in (65536, 0):
    ip.configure(fused_small=fs)
    l0 = ip.handle.launch_count(); dl.run(); print("fused_small", fs, "launches per layer", ip.handle.launch_count() - l0)
    D = dl.result("D").copy()
    for _ in range(20): dl.run(sync=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): dl.run(sync=False)
    e1.record(); torch.cuda.synchronize()
    print("  layer us", e0.elapsed_time(e1) * 1e3 / 200, "checksum", float(np.abs(D).sum()))
def timeit(d, tag):
    for _ in range(20): d.run(sync=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(200): d.run(sync=False)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print("  ", tag, "layer us", e0.elapsed_time(e1) * 1e3 / 200, "cpu issue us", (t1 - t0) * 1e6 / 200)
print("--- with a torch side stream")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ip.handle.set_stream(stream.cuda_stream)
ip.configure(fused_small=65536)
l0 = ip.handle.launch_count(); dl.run(); print("launches per layer", ip.handle.launch_count() - l0, ip.handle.last_error() if hasattr(ip.handle, "last_error") else "")
timeit(dl, "side stream")
print("--- after a big batch on the same handle")
import bench
batch, probs = bench.make_cora_batch(64, seed0=0, P=16)
dlb = DeviceLayer(ip.handle, _lib.MODE_F32_FAST)
dlb.load(N=batch.N, M=batch.M, P=batch.P, adj=(batch.adj_rowptr, batch.adj_col, batch.adj_val), fea=(batch.fea_rowptr, batch.fea_col, batch.fea_val), B=batch.B, relu=1)
dlb.run(); dlb.run()
p1 = probs[0]
dl1 = DeviceLayer(ip.handle, _lib.MODE_F32_FAST)
dl1.load(N=p1.N, M=p1.M, P=p1.P, adj=(p1.adj_rowptr, p1.adj_col, p1.adj_val), fea=(p1.fea_rowptr, p1.fea_col, p1.fea_val), B=p1.B, relu=1)
l0 = ip.handle.launch_count(); dl1.run(); print("launches per layer", ip.handle.launch_count() - l0)
ip.configure(index_format=0)
l0 = ip.handle.launch_count(); dl1.run(); print("launches per layer (index_format=0)", ip.handle.launch_count() - l0)
timeit(dl1, "probs[0]")
timeit(dl, "cora_shape again")
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for _ in range(5):
    ip.handle.fea_run(dlb.desc, dlb.t["XW"].data_ptr()); ip.handle.adj_run(dlb.desc, dlb.t["XW"].data_ptr(), batch.N)
torch.cuda.synchronize()
timeit(dl1, "probs[0] after fea_run/adj_run")
