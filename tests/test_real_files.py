"""The reference's own data files (BASELINE configs[0] and SURVEY 8d "also run on the real files"):
data/matrices/mol_*.txt (main_float.cpp:40-51), cora_*.txt (main_float.cpp:73-82) and the MUTAG raw/ batch of the
molecule notebook (cells 10, 16-18).  tests/golden/real_*.npz hold the inputs and the outputs of the reference HLS
source compiled natively (oracle/_ref, HALF and FLOAT builds; tests/golden/make_golden.py real).

CPU: the oracle restatement reproduces those outputs bit for bit.
GPU: the C-simulation modes reproduce them bit for bit through the register map; the float32 fast mode stays
within 1e-5 (row-normalised, and the plain element-wise relative error on elements above 1e-3 of their row maximum
is reported and bounded too)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import util as U

KINDS = (("half", O.F16), ("float", O.F32))


def _mol():
    g = np.load(os.path.join(U.GOLDEN, "real_mol.npz"))
    adj = (g["adj_rowptr"], g["adj_col"].astype(np.int32), g["adj_val"])
    fea = (g["fea_rowptr"], g["fea_col"].astype(np.int32), g["fea_val"])
    n = len(adj[0]) - 1
    xd = np.zeros((n, 7), np.float32)
    xd[np.repeat(np.arange(n), np.diff(fea[0])), fea[1]] = fea[2]
    return g, n, adj, fea, xd


def _cora():
    g = np.load(os.path.join(U.GOLDEN, "real_cora.npz"))
    adj = (g["adj_rowptr"], g["adj_col"].astype(np.int32), g["adj_val"])
    fv = np.ones(int(g["fea_rowptr"][-1]), np.float32) if bool(g["fea_all_ones"]) else g["fea_val"]
    fea = (g["fea_rowptr"], g["fea_col"].astype(np.int32), fv)
    return g, len(adj[0]) - 1, adj, fea


def _mutag():
    g = np.load(os.path.join(U.GOLDEN, "real_mutag.npz"))
    edges = g["edges"].astype(np.int64)
    n = len(g["graph_indicator"])
    order = np.lexsort((edges[:, 1], edges[:, 0]))
    e = edges[order]
    rp = np.zeros(n + 1, np.int32)
    np.cumsum(np.bincount(e[:, 0], minlength=n), out=rp[1:])
    adj = (rp, e[:, 1].astype(np.int32), np.ones(len(e), np.float32))
    fea = (np.arange(n + 1, dtype=np.int32), g["node_labels"].astype(np.int32), np.ones(n, np.float32))
    return g, n, adj, fea


def _st(triple, dt):
    return (triple[0], triple[1], O.to_storage(triple[2], dt))


def elementwise_rel(got, want, floor=1e-3):
    """max |got - want| / |want| over the elements whose magnitude is at least `floor` of their row maximum"""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    big = np.abs(want) >= floor * np.maximum(np.abs(want).max(axis=-1, keepdims=True), 1e-30)
    return float((np.abs(got - want)[big] / np.abs(want)[big]).max()) if big.any() else 0.0


# ------------------------------------------------------------------------------------------------------
# CPU: oracle vs the compiled reference source on the real files
# ------------------------------------------------------------------------------------------------------
def test_oracle_reproduces_reference_on_mol_files():
    g, n, adj, fea, xd = _mol()
    for kind, dt in KINDS:
        for P in (16, 32):
            B1 = O.to_storage(O.weights_to_B(g["w"][:, :P]), dt)
            h1 = O.layer(dtype=dt, N=n, M_fea=7, P=P, adj=_st(adj, dt), fea=_st(fea, dt), B=B1, relu=1, lat_fea=4, lat_adj=4)
            assert np.array_equal(h1.view(np.uint8), g[f"{kind}_P{P}_layer1"].view(np.uint8)), (kind, P)
            h1d = O.layer(dtype=dt, N=n, M_fea=7, P=P, adj=_st(adj, dt), x_dense=O.to_storage(xd, dt), B=B1, relu=1,
                          lat_fea=4, lat_adj=4)
            assert np.array_equal(h1d.view(np.uint8), g[f"{kind}_P{P}_layer1_dense"].view(np.uint8)), (kind, P)
            B2 = O.to_storage(O.weights_to_B(g[f"w2_P{P}"]), dt)
            h2 = O.layer(dtype=dt, N=n, M_fea=P, P=P, adj=_st(adj, dt), x_dense=h1, B=B2, relu=0, lat_fea=4, lat_adj=4)
            assert np.array_equal(h2.view(np.uint8), g[f"{kind}_P{P}_layer2"].view(np.uint8)), (kind, P)


def test_oracle_reproduces_reference_on_cora_files():
    g, n, adj, fea = _cora()
    for kind, dt in KINDS:
        B = O.to_storage(O.weights_to_B(g["w"]), dt)
        for relu in (0, 1):
            d = O.layer(dtype=dt, N=n, M_fea=1433, P=16, adj=_st(adj, dt), fea=_st(fea, dt), B=B, relu=relu, lat_fea=4, lat_adj=4)
            assert np.array_equal(d.view(np.uint8), g[f"{kind}_relu{relu}"].view(np.uint8)), (kind, relu)


def test_oracle_reproduces_reference_on_mutag_batch():
    g, n, adj, fea = _mutag()
    assert (n, len(adj[1])) == (3371, 7442)
    for kind, dt in KINDS:
        h1 = O.layer(dtype=dt, N=n, M_fea=7, P=64, adj=_st(adj, dt), fea=_st(fea, dt), B=O.to_storage(O.weights_to_B(g["w1"]), dt),
                     relu=1, lat_fea=4, lat_adj=4)
        if kind == "half":
            assert np.array_equal(h1, g["half_layer1"])
        h2 = O.layer(dtype=dt, N=n, M_fea=64, P=64, adj=_st(adj, dt), x_dense=h1, B=O.to_storage(O.weights_to_B(g["w2"]), dt),
                     relu=0, lat_fea=4, lat_adj=4)
        assert np.array_equal(h2.view(np.uint8), g[f"{kind}_layer2"].view(np.uint8)), kind


# ------------------------------------------------------------------------------------------------------
# GPU: the library through the register map
# ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ip():
    from sgracex1_b200.pynq_compat import MmultTop
    return MmultTop(0)


def _run(ip, mode, **kw):
    from tests.test_gpu_parity import run_host
    return run_host(ip, mode, lat_fea=4, lat_adj=4, **kw)


@pytest.mark.gpu
def test_gpu_two_layer_mol_forward_matches_reference(ip):
    """BASELINE configs[0]: 2-layer GCN forward on mol_adj.txt (layer 1 sparse + ReLU, layer 2 gemm_mode on its output)"""
    from sgracex1_b200 import _lib
    from tests.test_gpu_parity import MODE_OF
    g, n, adj, fea, xd = _mol()
    for kind, dt in KINDS:
        for P in (16, 32):
            B1 = O.to_storage(O.weights_to_B(g["w"][:, :P]), dt)
            B2 = O.to_storage(O.weights_to_B(g[f"w2_P{P}"]), dt)
            h1 = _run(ip, MODE_OF[dt], N=n, M=7, P=P, adj=_st(adj, dt), fea=_st(fea, dt), B=B1, relu=1)
            assert np.array_equal(h1.view(np.uint8), g[f"{kind}_P{P}_layer1"].view(np.uint8)), (kind, P)
            h1d = _run(ip, MODE_OF[dt], N=n, M=7, P=P, adj=_st(adj, dt), x_dense=O.to_storage(xd, dt), B=B1, relu=1)
            assert np.array_equal(h1d.view(np.uint8), g[f"{kind}_P{P}_layer1_dense"].view(np.uint8)), (kind, P)
            h2 = _run(ip, MODE_OF[dt], N=n, M=P, P=P, adj=_st(adj, dt), x_dense=h1, B=B2, relu=0)
            assert np.array_equal(h2.view(np.uint8), g[f"{kind}_P{P}_layer2"].view(np.uint8)), (kind, P)
    # float32 fast mode against the FLOAT build
    for P in (16, 32):
        B1, B2 = O.weights_to_B(g["w"][:, :P]), O.weights_to_B(g[f"w2_P{P}"])
        for fused in (65536, 0):          # one cooperative launch / the streaming kernels
            h1 = _run(ip, _lib.MODE_F32_FAST, N=n, M=7, P=P, adj=adj, fea=fea, B=B1, relu=1, fused_small=fused)
            U.assert_close_f32(h1, g[f"float_P{P}_layer1"], what=f"mol layer 1 P={P}")
            assert elementwise_rel(h1, g[f"float_P{P}_layer1"]) < 2e-4
            h2 = _run(ip, _lib.MODE_F32_FAST, N=n, M=P, P=P, adj=adj, x_dense=h1, B=B2, relu=0, fused_small=fused)
            U.assert_close_f32(h2, g[f"float_P{P}_layer2"], rtol=2e-5, what=f"mol layer 2 P={P}")


@pytest.mark.gpu
def test_gpu_real_cora_layer_matches_reference(ip):
    from sgracex1_b200 import _lib
    from tests.test_gpu_parity import MODE_OF
    g, n, adj, fea = _cora()
    for kind, dt in KINDS:
        B = O.to_storage(O.weights_to_B(g["w"]), dt)
        for relu in (0, 1):
            d = _run(ip, MODE_OF[dt], N=n, M=1433, P=16, adj=_st(adj, dt), fea=_st(fea, dt), B=B, relu=relu)
            assert np.array_equal(d.view(np.uint8), g[f"{kind}_relu{relu}"].view(np.uint8)), (kind, relu)
    B = O.weights_to_B(g["w"])
    for fused in (65536, 0):
        for relu in (0, 1):
            d = _run(ip, _lib.MODE_F32_FAST, N=n, M=1433, P=16, adj=adj, fea=fea, B=B, relu=relu, fused_small=fused)
            U.assert_close_f32(d, g[f"float_relu{relu}"], what=f"cora relu={relu} fused={fused}")
            assert elementwise_rel(d, g[f"float_relu{relu}"]) < 2e-4, (relu, fused)


@pytest.mark.gpu
def test_gpu_mutag_notebook_layers_match_reference(ip):
    """The notebook's batch (188 MUTAG graphs, raw/ files), fp16 buffers as on the board: both GCN layers bit-exact"""
    from sgracex1_b200 import _lib
    from tests.test_gpu_parity import MODE_OF
    g, n, adj, fea = _mutag()
    for kind, dt in KINDS:
        h1 = _run(ip, MODE_OF[dt], N=n, M=7, P=64, adj=_st(adj, dt), fea=_st(fea, dt), B=O.to_storage(O.weights_to_B(g["w1"]), dt), relu=1)
        if kind == "half":
            assert np.array_equal(h1, g["half_layer1"])
        h2 = _run(ip, MODE_OF[dt], N=n, M=64, P=64, adj=_st(adj, dt), x_dense=h1, B=O.to_storage(O.weights_to_B(g["w2"]), dt), relu=0)
        assert np.array_equal(h2.view(np.uint8), g[f"{kind}_layer2"].view(np.uint8)), kind
    h1 = _run(ip, _lib.MODE_F32_FAST, N=n, M=7, P=64, adj=adj, fea=fea, B=O.weights_to_B(g["w1"]), relu=1)
    h2 = _run(ip, _lib.MODE_F32_FAST, N=n, M=64, P=64, adj=adj, x_dense=h1, B=O.weights_to_B(g["w2"]), relu=0)
    U.assert_close_f32(h2, g["float_layer2"], rtol=2e-5, what="MUTAG layer 2 (float32 fast)")
