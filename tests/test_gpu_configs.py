"""GPU parity at the shapes BASELINE.json names (configs[1], configs[2]) and size-independent
properties at benchmark scale."""
import numpy as np
import pytest

from oracle import oracle as O
from sgracex1_b200 import _lib, graphs as G, quant as Q
from tests import util as U
from tests.test_gpu_parity import run_full, run_host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ip():
    from sgracex1_b200.pynq_compat import MmultTop
    return MmultTop(0)


def pruned(p, keep=0.5, seed=1):
    """Random edge drop (keeps the diagonal) on top of the quantiser's own pruning of small entries."""
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(p.N), np.diff(p.adj_rowptr))
    keep_mask = (rng.random(len(rows)) < keep) | (rows == p.adj_col)
    rp = np.zeros(p.N + 1, np.int32)
    np.cumsum(np.bincount(rows[keep_mask], minlength=p.N), out=rp[1:])
    return rp, p.adj_col[keep_mask], p.adj_val[keep_mask]


@pytest.mark.parametrize("shape", ["citeseer", "pubmed"])
@pytest.mark.parametrize("qbits", [8, 4])
def test_config2_quantised_gat_on_pruned_adjacency(ip, shape, qbits):
    p = G.citeseer_shape() if shape == "citeseer" else G.pubmed_shape()
    adj = pruned(p)
    fea = (p.fea_rowptr, p.fea_col, p.fea_val)
    rng = np.random.default_rng(7)
    bound = 1.414 * np.sqrt(6.0 / (2 * p.P + 1))                  # xavier_uniform(gain 1.414) on (2P, 1)
    att = rng.uniform(-bound, bound, size=2 * p.P).astype(np.float32)
    consts = Q.layer_constants(qbits)
    B = O.weights_to_B(p.W)
    for gat in (1, 0):
        ref = O.qlayer(N=p.N, M_fea=p.M, P=p.P, adj=adj, fea=fea, B=B, attention=att, relu=1, gat=gat, qbits=qbits,
                       consts=consts, return_all=True)
        D, E, S, max_fea = run_full(ip, qbits=qbits, gat=gat, N=p.N, M=p.M, P=p.P, adj=adj, fea=fea, B=B, attention=att, relu=1)
        assert max_fea == ref["max_fea"]
        if gat:
            U.assert_close_f32(D, ref["D"], what=f"{shape} GAT q{qbits}")
            np.testing.assert_allclose(E, ref["E"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(S, ref["S"], rtol=1e-5, atol=1e-7)
        else:
            assert np.array_equal(D, ref["D"]), f"{shape} quantised GCN q{qbits} not bit-exact"


def test_config1_cora_shape_half_csim_bit_exact(ip):
    p = G.cora_shape(seed=3)
    a16 = (p.adj_rowptr, p.adj_col, O.to_storage(p.adj_val, O.F16))
    f16 = (p.fea_rowptr, p.fea_col, O.to_storage(p.fea_val, O.F16))
    B16 = O.to_storage(p.B, O.F16)
    for knobs in (dict(spmm_block=1), dict(spmm_block=4, lat_fea=4, lat_adj=4)):
        ref = O.layer(dtype=O.F16, N=p.N, M_fea=p.M, P=p.P, adj=a16, fea=f16, B=B16, relu=1, **knobs)
        got = run_host(ip, _lib.MODE_F16_CSIM, N=p.N, M=p.M, P=p.P, adj=a16, fea=f16, B=B16, relu=1, **knobs)
        assert np.array_equal(got, ref), knobs


def test_benchmark_scale_properties(ip):
    """cora_x64 through the device-resident path: (i) every replica of a graph gives the same rows,
    (ii) positive homogeneity act(A (2X) W) == 2 act(A X W) bit for bit (scaling by 2 is exact),
    (iii) the first graph matches the oracle."""
    from sgracex1_b200.driver import DeviceLayer
    probs = [G.cora_shape(seed=s) for s in range(4)]
    b = G.block_diagonal(probs, 64)
    ip.configure(staging=0)
    dl = DeviceLayer(ip.handle, _lib.MODE_F32_FAST)
    adj, fea = (b.adj_rowptr, b.adj_col, b.adj_val), (b.fea_rowptr, b.fea_col, b.fea_val)
    dl.load(N=b.N, M=b.M, P=b.P, adj=adj, fea=fea, B=b.B, relu=1)
    dl.run()
    D = dl.result("D").copy()
    n = probs[0].N
    for k in range(4, 64):
        assert np.array_equal(D[k * n:(k + 1) * n], D[(k % 4) * n:(k % 4 + 1) * n]), k
    dl.load(N=b.N, M=b.M, P=b.P, adj=adj, fea=(fea[0], fea[1], 2.0 * fea[2]), B=b.B, relu=1)
    dl.run()
    assert np.array_equal(dl.result("D"), 2.0 * D)
    p0 = probs[0]
    ref = O.layer(dtype=O.F32, N=p0.N, M_fea=p0.M, P=p0.P, adj=(p0.adj_rowptr, p0.adj_col, p0.adj_val),
                  fea=(p0.fea_rowptr, p0.fea_col, p0.fea_val), B=p0.B, relu=1)
    U.assert_close_f32(D[:n], ref, what="first graph of the batch")
    ip.configure(staging=1)


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("banded", [True, False])
def test_overlapped_staging_equals_the_serial_path(ip, dense, banded):
    """sgrace_start with host buffers on a layer whose D exceeds 16 MB (SGRACE_OPT_OVERLAP).  A block-diagonal batch is
    pipelined in row chunks (features up -> FEA -> the adjacency panels whose columns are covered -> D down); a graph
    whose rows reference columns anywhere takes the panelled adjacency upload / download after one feature stage.
    Both give the bits of the serial path, twice in a row (buffers and events are reused)."""
    from sgracex1_b200.driver import HostLayer
    probs = [G.cora_shape(seed=s, n=2708, m=96, nnz_fea=9000) for s in range(3)]
    b = G.block_diagonal(probs, 120)
    assert b.N * b.P * 4 >= 16 << 20
    rng = np.random.default_rng(3)
    adj, fea = (b.adj_rowptr, b.adj_col, b.adj_val), (b.fea_rowptr, b.fea_col, b.fea_val)
    if not banded:
        ci = rng.integers(0, b.N, size=b.nnz_adj).astype(np.int32)       # columns anywhere (order within a row is free)
        adj = (b.adj_rowptr, ci, b.adj_val)
    xd = rng.standard_normal((b.N, 8)).astype(np.float32) if dense else None
    M = 8 if dense else b.M
    B = (rng.uniform(-0.3, 0.3, size=(b.P, 8)).astype(np.float32).reshape(-1)) if dense else b.B
    hl = HostLayer(ip, _lib.MODE_F32_FAST, N=b.N, M=M, P=b.P, nnz_adj=b.nnz_adj, nnz_fea=0 if dense else b.nnz_fea, dense=dense)
    outs = []
    for overlap in (0, 1, 1):
        ip.configure(mode=_lib.MODE_F32_FAST, staging=1, index_format=0, overlap=overlap)
        before = ip.handle.get_option(_lib.OPT_OVERLAPPED_STARTS)
        before_p = ip.handle.get_option(_lib.OPT_PIPELINED_STARTS)
        if dense:
            hl.load(N=b.N, M=M, P=b.P, adj=adj, x_dense=xd, B=B, relu=1)
        else:
            hl.load(N=b.N, M=M, P=b.P, adj=adj, fea=fea, B=B, relu=1)
        hl.D[:] = -3.0
        outs.append(hl.run().copy())
        assert ip.handle.get_option(_lib.OPT_OVERLAPPED_STARTS) - before == overlap
        assert ip.handle.get_option(_lib.OPT_PIPELINED_STARTS) - before_p == (overlap if banded else 0)
    hl.free()
    ip.configure(overlap=1)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    if not dense:
        rows = np.r_[0:300, b.N - 300:b.N]
        ref = O.layer(dtype=O.F32, N=b.N, M_fea=M, P=b.P, adj=adj, fea=fea, B=B, relu=1)
        U.assert_close_f32(outs[1][rows], ref[rows], what="overlapped staging")
