"""ADJ stage on block-diagonal adjacencies with the XW window in shared memory (csrc/sgrace_spmm_panel.cuh):
bit-equal to the gather kernel, 1e-5 against the oracle, correct under a stale plan and on graphs it does not fit.
Reference semantics: loop_adj, kernelMatrixmult_all.cpp:3339-3627 (CSR order within a row, ReLU, row-major D)."""
import numpy as np
import pytest

from oracle import oracle as O
from sgracex1_b200 import _lib
from tests import util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ip():
    from sgracex1_b200.pynq_compat import MmultTop
    return MmultTop(0)


def batched_adjacency(sizes, rng, avg_deg=4, hub_every=0, empty_frac=0.1):
    """Block-diagonal CSR adjacency of len(sizes) random graphs; some rows empty, an occasional hub row."""
    rows, off = [], 0
    for gi, n in enumerate(sizes):
        for r in range(n):
            if rng.random() < empty_frac:
                rows.append(np.zeros(0, np.int64))
                continue
            d = min(n, 1 + rng.poisson(avg_deg))
            if hub_every and gi % hub_every == 0 and r == n // 2:
                d = n
            rows.append(off + np.sort(rng.choice(n, size=d, replace=False)))
        off += n
    deg = np.array([len(x) for x in rows])
    rp = np.zeros(off + 1, np.int32)
    np.cumsum(deg, out=rp[1:])
    ci = np.concatenate(rows).astype(np.int32)
    av = rng.uniform(-0.5, 0.5, size=len(ci)).astype(np.float32)
    return off, (rp, ci, av)


def adj_stage(ip, adj, xw, N, P, relu, plan, long_row=512):
    """One ADJ launch on device-resident buffers; returns D and how many panel-kernel launches it made."""
    import torch
    ip.configure(mode=_lib.MODE_F32_FAST, staging=0, index_format=0, adj_plan=plan, long_row=long_row)
    t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in adj]
    x = torch.from_numpy(xw).cuda()
    D = torch.full((N, P), -7.0, dtype=torch.float32, device="cuda")
    d = _lib.LayerDesc()
    d.N_adj, d.M_adj, d.M_fea, d.P_w, d.relu, d.nnz_adj = N, N, 1, P, relu, len(adj[1])
    d.rowPtr_adj, d.columnIndex_adj, d.values_adj, d.D = t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), D.data_ptr()
    before = ip.handle.get_option(_lib.OPT_PANEL_LAUNCHES)
    ip.handle.adj_run(d, x.data_ptr(), N)
    torch.cuda.synchronize()
    n = ip.handle.get_option(_lib.OPT_PANEL_LAUNCHES) - before
    ip.configure(adj_plan=0, long_row=512)
    return D.cpu().numpy(), n, (t, x, d, D)


def spmm_ref(adj, xw, relu):
    import scipy.sparse as sp
    rp, ci, av = adj
    A = sp.csr_matrix((av.astype(np.float64), ci, rp), shape=(len(rp) - 1, xw.shape[0]))
    out = A @ xw.astype(np.float64)
    return np.maximum(out, 0) if relu else out


@pytest.mark.parametrize("P", [16, 64, 8, 128, 4, 256])
def test_panel_adj_bit_equal_to_gather(ip, P):
    rng = np.random.default_rng(P)
    sizes = rng.integers(40, 420 if P <= 64 else (300 if P <= 128 else 150), size=400 if P <= 128 else 700)
    N, adj = batched_adjacency(sizes, rng, hub_every=37)
    xw = rng.standard_normal((N, P)).astype(np.float32)
    for relu in (1, 0):
        for long_row in (512, 64):
            want, n0, _ = adj_stage(ip, adj, xw, N, P, relu, plan=0, long_row=long_row)
            got, n1, _ = adj_stage(ip, adj, xw, N, P, relu, plan=1, long_row=long_row)
            assert n0 == 0 and n1 == 1, (n0, n1)
            assert np.array_equal(got, want), f"P={P} relu={relu} long_row={long_row}"
    ref = spmm_ref(adj, xw, 0)
    U.assert_close_f32(got, ref.astype(np.float32), what=f"panel ADJ P={P}")


def test_panel_layer_matches_oracle(ip):
    """Whole layer (FEA -> ADJ) through the register-free entry on a batch of graphs, against the CPU oracle."""
    from sgracex1_b200 import graphs as G
    from sgracex1_b200.driver import DeviceLayer
    import torch
    probs = [G.cora_shape(seed=s, n=300, m=96, nnz_adj=1500, nnz_fea=2400) for s in range(4)]
    b = G.block_diagonal(probs, 160)
    ip.configure(mode=_lib.MODE_F32_FAST, staging=0, index_format=0, adj_plan=1, fused_small=0)
    dl = DeviceLayer(ip.handle, _lib.MODE_F32_FAST, device="cuda:0")
    adj, fea = (b.adj_rowptr, b.adj_col, b.adj_val), (b.fea_rowptr, b.fea_col, b.fea_val)
    dl.load(N=b.N, M=b.M, P=b.P, adj=adj, fea=fea, B=b.B, relu=1)
    before = ip.handle.get_option(_lib.OPT_PANEL_LAUNCHES)
    ip.handle.layer_run(dl.desc)
    torch.cuda.synchronize()
    assert ip.handle.get_option(_lib.OPT_PANEL_LAUNCHES) == before + 1
    got = dl.t["D"].cpu().numpy().reshape(b.N, b.P)
    ip.configure(fused_small=65536, adj_plan=0)
    ref = O.layer(dtype=O.F32, N=b.N, M_fea=b.M, P=b.P, adj=adj, fea=fea, B=b.B, relu=1)
    U.assert_close_f32(got, ref, what="panel layer")


def test_stale_plan_and_unfit_graphs_stay_correct(ip):
    import torch
    rng = np.random.default_rng(5)
    sizes = rng.integers(100, 300, size=300)
    N, adj = batched_adjacency(sizes, rng)
    P = 16
    xw = rng.standard_normal((N, P)).astype(np.float32)
    got, n, (t, x, d, D) = adj_stage(ip, adj, xw, N, P, 1, plan=1)
    assert n == 1
    # same buffers, new contents: columns now cross the blocks the cached plan was made for
    ci2 = adj[1].copy()
    hit = rng.random(len(ci2)) < 0.2
    ci2[hit] = rng.integers(0, N, size=int(hit.sum()))
    t[1].copy_(torch.from_numpy(ci2))
    ip.configure(mode=_lib.MODE_F32_FAST, staging=0, index_format=0, adj_plan=1)
    before = ip.handle.get_option(_lib.OPT_PANEL_LAUNCHES)
    ip.handle.adj_run(d, x.data_ptr(), N)
    torch.cuda.synchronize()
    assert ip.handle.get_option(_lib.OPT_PANEL_LAUNCHES) == before + 1      # the cached plan was used
    ip.configure(adj_plan=0)
    U.assert_close_f32(D.cpu().numpy(), spmm_ref((adj[0], ci2, adj[2]), xw, 1).astype(np.float32), what="stale plan")
    # adj_plan = 2: the plan is analysed at every launch (the new contents are block-diagonal again after the restore)
    t[1].copy_(torch.from_numpy(adj[1]))
    builds = ip.handle.get_option(_lib.OPT_PLAN_BUILDS)
    ip.configure(mode=_lib.MODE_F32_FAST, staging=0, index_format=0, adj_plan=2)
    for _ in range(2):
        ip.handle.adj_run(d, x.data_ptr(), N)
    torch.cuda.synchronize()
    ip.configure(adj_plan=0)
    assert ip.handle.get_option(_lib.OPT_PLAN_BUILDS) == builds + 2
    assert np.array_equal(D.cpu().numpy(), got)
    # one connected random graph: no diagonal blocks -> the plan is not usable and the gather kernel runs
    N2 = 20000
    rp = np.arange(0, 4 * N2 + 1, 4, dtype=np.int32)
    ci = rng.integers(0, N2, size=4 * N2).astype(np.int32)
    ci.reshape(N2, 4).sort(axis=1)
    av = rng.uniform(-1, 1, size=4 * N2).astype(np.float32)
    xw2 = rng.standard_normal((N2, P)).astype(np.float32)
    got2, n2, _ = adj_stage(ip, (rp, ci, av), xw2, N2, P, 0, plan=1)
    assert n2 == 0
    U.assert_close_f32(got2, spmm_ref((rp, ci, av), xw2, 0).astype(np.float32), what="unfit graph")
