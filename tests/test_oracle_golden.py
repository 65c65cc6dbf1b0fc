"""CPU tests: the oracle against every golden vector the reference holds for this path.

  * csim log (Citeseer, HALF, SPMM_BLOCK=4 build): 42 values, bit-exact
  * on-board notebook output (row 0, 16 fp16 values): bit-exact
  * notebook scipy float32 A@(X@W) row 0 (21 values): 1e-5
  * the reference HLS source compiled natively (oracle/_ref, when built): whole matrices, bit-exact
  * fixtures of that compiled reference on seeded inputs (tests/golden/ref_hls.npz): bit-exact
  * the reference's own sgrace.py emulation (tests/golden/qlayer_*.npz): <= 2 ulp of float32
  * 4x4 toy fixtures: hand-derived answers
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from sgracex1_b200 import quant as Q
from tests import util as U


def _run_half(adj, fea, w16, n, P, sb, lat=4, relu=0):
    B = np.ascontiguousarray(w16[:, :P].T).view(np.uint16).reshape(-1)
    return O.layer(dtype=O.F16, N=n, M_fea=w16.shape[0], P=P, adj=adj, fea=fea, B=B, spmm_block=sb,
                   lat_fea=lat, lat_adj=lat, relu=relu)


def test_half_helpers_match_numpy():
    L = O.lib()
    rng = np.random.default_rng(0)
    f = np.concatenate([rng.standard_normal(4000).astype(np.float32) * s for s in (1e-8, 1e-5, 1e-3, 1, 1e3, 7e4)])
    f = np.concatenate([f, np.array([0.0, -0.0, 65504, 65519.9, 65520, 2 ** -24, 2 ** -25, 2 ** -25 * 1.0001, np.inf, -np.inf], np.float32)])
    want = f.astype(np.float16).view(np.uint16)
    got = np.array([L.sgo_f32_to_f16(float(x)) for x in f], np.uint16)
    assert np.array_equal(got, want)
    h = np.arange(0, 0x7C00, 7, dtype=np.uint16)
    back = np.array([L.sgo_f16_to_f32(int(x)) for x in h], np.float32)
    assert np.array_equal(back, h.view(np.float16).astype(np.float32))


def test_csim_log_42_values_bit_exact():
    g, n, adj, fea, w16 = U.citeseer_half()
    D = _run_half(adj, fea, w16, n, 21, sb=4).view(np.float16)
    want = g["csim_vals"].astype(np.float32).astype(np.float16)
    got = D[g["csim_rows"], g["csim_cols"]]
    assert np.array_equal(got.view(np.uint16), want.view(np.uint16))


def test_onboard_notebook_row0_bit_exact():
    g, n, adj, fea, w16 = U.citeseer_half()
    D = _run_half(adj, fea, w16, n, 16, sb=1).view(np.float16)
    want = g["nb37_row0"].astype(np.float32).astype(np.float16)
    assert np.array_equal(D[0].view(np.uint16), want.view(np.uint16))


def test_scipy_float_row0():
    g, n, adj, fea, w16 = U.citeseer_half()
    a = (adj[0], adj[1], adj[2].view(np.float16).astype(np.float32))
    f = (fea[0], fea[1], np.ones(len(fea[1]), np.float32))
    B = O.weights_to_B(w16.astype(np.float32))
    D = O.layer(dtype=O.F32, N=n, M_fea=w16.shape[0], P=21, adj=a, fea=f, B=B)
    np.testing.assert_allclose(D[0], g["nb55_row0"], rtol=1e-5)


def test_reference_hls_fixtures_bit_exact():
    g = np.load(os.path.join(U.GOLDEN, "ref_hls.npz"))
    N, M = int(g["N"]), int(g["M"])
    for kind, dt in (("half", O.F16), ("float", O.F32)):
        adj = (g["adj_rowptr"], g["adj_col"], O.to_storage(g["adj_val"], dt))
        fea = (g["fea_rowptr"], g["fea_col"], O.to_storage(g["fea_val"], dt))
        for P in (16, 7):
            for relu in (0, 1):
                B = O.to_storage(O.weights_to_B(g["W"][:, :P]), dt)
                D = O.layer(dtype=dt, N=N, M_fea=M, P=P, adj=adj, fea=fea, B=B, relu=relu, spmm_block=1,
                            lat_fea=4, lat_adj=4)
                assert np.array_equal(D.view(np.uint8), g[f"{kind}_sparse_P{P}_relu{relu}"].view(np.uint8))
        Bd = O.to_storage(O.weights_to_B(g["w_dense"]), dt)
        D = O.layer(dtype=dt, N=N, M_fea=24, P=10, adj=adj, x_dense=O.to_storage(g["x_dense"], dt), B=Bd, relu=1,
                    spmm_block=1, lat_fea=4, lat_adj=4)
        assert np.array_equal(D.view(np.uint8), g[f"{kind}_dense_P10_relu1"].view(np.uint8))


def test_reference_hls_fix16_fixtures_bit_exact():
    """FIX16 (ap_fixed<16,2>): the oracle against outputs of the reference's own source compiled in its EIGHTBIT
    configuration (oracle/hls_shim/ref_kernel_eightbit.cpp) -- the dataflow and the order of operations are the
    reference's, the arithmetic type is the stand-in of oracle/hls_shim/ap_int.h (AP_TRN / AP_WRAP, the documented
    defaults; the Xilinx header is not available).  One set stays in range, one wraps around."""
    g = np.load(os.path.join(U.GOLDEN, "ref_hls_fix16.npz"))
    for tag in ("small", "wrap"):
        N, M = int(g[f"{tag}_N"]), int(g[f"{tag}_M"])
        adj = (g[f"{tag}_adj_rowptr"], g[f"{tag}_adj_col"], g[f"{tag}_adj_val"])
        fea = (g[f"{tag}_fea_rowptr"], g[f"{tag}_fea_col"], g[f"{tag}_fea_val"])
        for P in (16, 7):
            for relu in (0, 1):
                D = O.layer(dtype=O.FIX16, N=N, M_fea=M, P=P, adj=adj, fea=fea, B=g[f"{tag}_B_P{P}"], relu=relu, spmm_block=1,
                            lat_fea=1, lat_adj=1)
                assert np.array_equal(D, g[f"{tag}_sparse_P{P}_relu{relu}"]), (tag, P, relu)
        D = O.layer(dtype=O.FIX16, N=N, M_fea=24, P=10, adj=adj, x_dense=g[f"{tag}_x_dense"], B=g[f"{tag}_B_dense"], relu=1,
                    spmm_block=1, lat_fea=1, lat_adj=1)
        assert np.array_equal(D, g[f"{tag}_dense_P10_relu1"]), tag
    # the wrap-around set really wraps: its exact result leaves [-2, 2)
    import scipy.sparse as sp
    N, M = int(g["wrap_N"]), int(g["wrap_M"])
    A = sp.csr_matrix((g["wrap_adj_val"].astype(np.float64) / 16384, g["wrap_adj_col"], g["wrap_adj_rowptr"]), shape=(N, N))
    X = sp.csr_matrix((g["wrap_fea_val"].astype(np.float64) / 16384, g["wrap_fea_col"], g["wrap_fea_rowptr"]), shape=(N, M))
    W = g["wrap_B_P16"].reshape(16, M).T.astype(np.float64) / 16384
    assert (np.abs(A @ (X @ W)) >= 2).sum() > 0


@pytest.mark.skipif(not O.ref_available("fix16"), reason="oracle/_ref EIGHTBIT build of the reference source not built")
def test_against_compiled_reference_fix16():
    for seed, scale in ((5, 0.5), (7, 1.5), (8, 3.0)):
        pr = U.random_problem(seed, n=257, m=40, p=7, val_scale=scale)
        a, f, B, xd = U.to_storage_problem(pr, O.FIX16)
        for relu in (0, 1):
            ref = O.ref_layer(kind="fix16", N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, fea=f, B=B, relu=relu)
            mine = O.layer(dtype=O.FIX16, N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, fea=f, B=B, relu=relu, spmm_block=1,
                           lat_fea=1, lat_adj=1)
            assert np.array_equal(ref, mine)
            ref = O.ref_layer(kind="fix16", N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, x_dense=xd, B=B, relu=relu)
            mine = O.layer(dtype=O.FIX16, N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, x_dense=xd, B=B, relu=relu, spmm_block=1,
                           lat_fea=1, lat_adj=1)
            assert np.array_equal(ref, mine)


@pytest.mark.skipif(not (O.ref_available("half") and O.ref_available("float")),
                    reason="oracle/_ref (reference HLS source compiled natively) not built")
def test_against_compiled_reference_full_matrix():
    g, n, adj, fea, w16 = U.citeseer_half()
    for P, relu in ((16, 0), (21, 1)):
        B = np.ascontiguousarray(w16[:, :P].T).view(np.uint16).reshape(-1)
        ref = O.ref_layer(kind="half", N=n, M_fea=w16.shape[0], P=P, adj=adj, fea=fea, B=B, relu=relu)
        mine = _run_half(adj, fea, w16, n, P, sb=1, relu=relu)
        assert np.array_equal(ref, mine)
    pr = U.random_problem(5, n=300, m=50, p=12)
    for kind, dt in (("half", O.F16), ("float", O.F32)):
        a, f, B, xd = U.to_storage_problem(pr, dt)
        for relu in (0, 1):
            ref = O.ref_layer(kind=kind, N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, fea=f, B=B, relu=relu)
            mine = O.layer(dtype=dt, N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, fea=f, B=B, relu=relu,
                           spmm_block=1, lat_fea=4, lat_adj=4)
            assert np.array_equal(ref.view(np.uint8), mine.view(np.uint8))
            ref = O.ref_layer(kind=kind, N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, x_dense=xd, B=B, relu=relu)
            mine = O.layer(dtype=dt, N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, x_dense=xd, B=B, relu=relu,
                           spmm_block=1, lat_fea=4, lat_adj=4)
            assert np.array_equal(ref.view(np.uint8), mine.view(np.uint8))


def test_toy_fixtures_known_answers():
    g = np.load(os.path.join(U.GOLDEN, "toy.npz"))
    # A = row0 [.5 .5 0 0]; X = row0 [1 2 0 0]; W 8x2 with W[0,0]=1  ->  XW[0]=[1,0], D[0]=[.5,0]
    adj = (g["test_adj_rowptr"], g["test_adj_col"], g["test_adj_val"])
    fea = (g["test_feat_rowptr"], g["test_feat_col"], g["test_feat_val"])
    W = g["test_weights"][:4]
    D, XW = O.layer(dtype=O.F32, N=4, M_fea=4, P=2, adj=adj, fea=fea, B=O.weights_to_B(W), return_xw=True)
    assert np.array_equal(XW, np.array([[1, 0], [0, 0], [0, 0], [0, 0]], np.float32))
    assert np.array_equal(D, np.array([[0.5, 0], [0, 0], [0, 0], [0, 0]], np.float32))
    # A2 = row0 all ones; X2 = all ones; W2 = [[1,.5],[1,-.5],[1,.5],[1,-.5]] -> XW rows [4,0]; D[0]=[16,0]
    adj = (g["test_adj2_rowptr"], g["test_adj2_col"], g["test_adj2_val"])
    fea = (g["test_feat2_rowptr"], g["test_feat2_col"], g["test_feat2_val"])
    D = O.layer(dtype=O.F32, N=4, M_fea=4, P=2, adj=adj, fea=fea, B=O.weights_to_B(g["test_weights2"]))
    assert np.array_equal(D, np.array([[16, 0], [0, 0], [0, 0], [0, 0]], np.float32))
    # the same in Q2.14: 16 wraps to 0 in ap_fixed<16,2>; with A scaled it stays in range
    D = O.layer(dtype=O.FIX16, N=4, M_fea=4, P=2, adj=(adj[0], adj[1], O.to_storage(adj[2] * 0.0625, O.FIX16)),
                fea=(fea[0], fea[1], O.to_storage(fea[2] * 0.25, O.FIX16)),
                B=O.to_storage(O.weights_to_B(g["test_weights2"]), O.FIX16))
    assert np.array_equal(O.from_storage(D, O.FIX16), np.array([[0.25, 0], [0, 0], [0, 0], [0, 0]], np.float32))


def test_fix16_wrap_and_truncation():
    # 1.5 * 1.5 = 2.25 wraps to -1.75 in ap_fixed<16,2>; -(2^-14) * 0.5 truncates toward -inf to -(2^-14)
    one = np.array([0, 1], np.int32), np.array([0], np.int32)
    def run(a, x, w):
        return O.from_storage(O.layer(dtype=O.FIX16, N=1, M_fea=1, P=1, adj=(one[0], one[1], O.to_storage([a], O.FIX16)),
                                      fea=(one[0], one[1], O.to_storage([x], O.FIX16)), B=O.to_storage([w], O.FIX16)), O.FIX16)[0, 0]
    assert run(1.0, 1.5, 1.5) == -1.75
    assert run(1.0, -2.0 ** -14, 0.5) == -2.0 ** -14
    assert run(1.0, 2.0 ** -14, 0.5) == 0.0


def test_knobs_change_half_results_but_not_fix16():
    pr = U.random_problem(3, n=120, m=48, p=8)
    a, f, B, _ = U.to_storage_problem(pr, O.F16)
    base = O.layer(dtype=O.F16, N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, fea=f, B=B, spmm_block=1)
    other = O.layer(dtype=O.F16, N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, fea=f, B=B, spmm_block=4)
    assert not np.array_equal(base, other)          # SPMM_BLOCK is a numerical parameter in HALF
    a, f, B, _ = U.to_storage_problem({**pr, "adj": (pr["adj"][0], pr["adj"][1], pr["adj"][2] * 0.2)}, O.FIX16)
    base = O.layer(dtype=O.FIX16, N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, fea=f, B=B, spmm_block=1)
    for sb, ft, at in ((4, 1, 1), (2, 2, 4), (8, 4, 2)):
        other = O.layer(dtype=O.FIX16, N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, fea=f, B=B, spmm_block=sb,
                        fea_threads=ft, adj_threads=at)
        assert np.array_equal(base, other)          # integer adds are associative: order-free


def test_relu_and_use_sblocks():
    pr = U.random_problem(4, n=64, m=32, p=5)
    a, f, B, _ = U.to_storage_problem(pr, O.F32)
    kw = dict(dtype=O.F32, N=pr["N"], M_fea=pr["M"], P=pr["P"], adj=a, fea=f, B=B)
    raw = O.layer(relu=0, **kw)
    assert (raw < 0).any()
    assert np.array_equal(O.layer(relu=1, **kw), np.maximum(raw, 0))
    assert np.array_equal(O.layer(relu=1, use_sblocks=1, **kw), raw)   # USE_SBLOCKS=1 drops the ReLU


@pytest.mark.parametrize("qbits", [8, 4, 2, 1])
@pytest.mark.parametrize("gat", [0, 1])
def test_qlayer_against_reference_emulation(qbits, gat):
    g = np.load(os.path.join(U.GOLDEN, f"qlayer_q{qbits}_gat{gat}.npz"))
    x, w, att = g["x"], g["w"], g["attention"]
    n, m = x.shape
    p = w.shape[1]
    consts = Q.layer_constants(qbits)
    for k in ("w_s", "a_s", "f_s", "deq_o", "scale_fea", "internal_quantization", "w_z", "a_z", "f_z"):
        assert float(g["c_" + k]) == float(consts[k]), k      # host constant tables match init_SGRACE
    adj = U.coo_to_csr(g["edge_index"], g["norm"], n)
    for relu in (0, 1):
        for dense in (0, 1):
            kw = dict(x_dense=x) if dense else dict(fea=U.dense_to_csr(x))
            out = O.qlayer(N=n, M_fea=m, P=p, adj=adj, B=O.weights_to_B(w), attention=att, relu=relu, gat=gat,
                           qbits=qbits, consts=consts, **kw)
            ref = g[f"out_relu{relu}_dense{dense}"]
            # integer stages are exact; the float adds of torch's sparse matmul have no defined
            # order, so agreement is to a couple of float32 ulps of the row magnitude
            tol = 4 * np.finfo(np.float32).eps * np.abs(ref).max()
            assert np.abs(out - ref).max() <= tol
            if qbits == 1 and not gat:
                assert np.array_equal(out, ref)


def test_qlayer_empty_row_gets_column_mean():
    # a row whose every edge quantises to zero: the dense emulation's softmax is uniform over all
    # nodes (sgrace.py:638-641), i.e. the column mean of Wh
    n, m, p = 6, 4, 3
    rng = np.random.default_rng(0)
    x = rng.random((n, m)).astype(np.float32)
    w = rng.uniform(-1, 1, (m, p)).astype(np.float32)
    att = rng.uniform(-1, 1, (2 * p, 1)).astype(np.float32)
    rp = np.array([0, 1, 2, 3, 4, 5, 6], np.int32)
    ci = np.arange(n, dtype=np.int32)
    av = np.full(n, 0.5, np.float32)
    av[2] = 0.001                                     # quantises to code 0 at 8 bits
    r = O.qlayer(N=n, M_fea=m, P=p, adj=(rp, ci, av), B=O.weights_to_B(w), x_dense=x, attention=att, gat=1,
                 qbits=8, consts=Q.layer_constants(8), return_all=True)
    want = r["Wh"].astype(np.float64).mean(0) * Q.layer_constants(8)["deq_o"]
    np.testing.assert_allclose(r["D"][2], want, rtol=1e-6)
    assert r["S"][2] == 0 and r["E"][2] == 0
    np.testing.assert_allclose(r["S"][[0, 1, 3, 4, 5]], 1.0)
