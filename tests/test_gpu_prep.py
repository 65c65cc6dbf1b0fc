"""Graph preparation on the GPU against the host mirror of the reference's code:
sym_norm2 (demo/sgrace_lib/sgrace.py:18-51) bit for bit, dense X -> CSR exactly."""
import os

import numpy as np
import pytest
import torch

from sgracex1_b200 import sgrace as S
from tests import util as U

pytestmark = pytest.mark.gpu


def _check_sym_norm(ei, n, w=None, fill=0.0):
    want_ei, want_norm = S.sym_norm2(ei, n, edge_weight=w, fill=fill)
    got_ei, got_norm = S.sym_norm2_device(ei.cuda(), n, edge_weight=None if w is None else w.cuda(), fill=fill)
    assert got_ei.dtype == ei.dtype
    assert torch.equal(got_ei.cpu(), want_ei)
    assert np.array_equal(got_norm.cpu().numpy().view(np.uint32), want_norm.numpy().view(np.uint32)), \
        f"max diff {(got_norm.cpu() - want_norm).abs().max()}"


def test_sym_norm_reference_fixture():
    g = np.load(os.path.join(U.GOLDEN, "sym_norm2.npz"))
    ei = torch.from_numpy(g["edge_index_in"])
    _check_sym_norm(ei, int(g["num_nodes"]), fill=float(g["fill"]))
    # and straight against the reference's own recorded output
    got_ei, got_norm = S.sym_norm2_device(ei.cuda(), int(g["num_nodes"]), fill=float(g["fill"]))
    assert np.array_equal(got_ei.cpu().numpy(), g["edge_index_out"])
    np.testing.assert_allclose(got_norm.cpu().numpy(), g["norm_out"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("fill", [0.0, 1.0, 2.0])
@pytest.mark.parametrize("weighted", [False, True])
def test_sym_norm_random_graphs(fill, weighted):
    rng = np.random.default_rng(11)
    n, e = 700, 6000
    row = rng.integers(0, n, size=e)
    col = rng.integers(0, n, size=e)
    col[::17] = row[::17]                         # existing self-loops (some nodes several times)
    row[-50:] = row[:50]; col[-50:] = col[:50]    # duplicate edges
    keep = (row != 5) & (col != 5) & (row != 6)   # node 5 isolated, node 6 without out-edges
    ei = torch.from_numpy(np.stack([row[keep], col[keep]]).astype(np.int64))
    w = torch.from_numpy(rng.uniform(0.1, 2.0, size=ei.shape[1]).astype(np.float32)) if weighted else None
    _check_sym_norm(ei, n, w=w, fill=fill)


def test_sym_norm_edge_cases():
    _check_sym_norm(torch.zeros((2, 0), dtype=torch.int64), 4, fill=1.0)          # no edges: identity
    _check_sym_norm(torch.tensor([[0, 1, 2], [0, 1, 2]]), 3, fill=0.0)           # only self-loops
    ei = torch.tensor([[0, 1], [1, 0]])
    _check_sym_norm(ei, 3, fill=0.0)                                             # fill 0: isolated node -> inf -> 0
    with pytest.raises(Exception):
        S.sym_norm2_device(torch.tensor([[0, 7], [1, 0]]).cuda(), 3)             # index out of range is reported


@pytest.mark.parametrize("shape", [(1, 1), (37, 7), (300, 100), (64, 1433), (5, 33)])
def test_dense_to_csr(shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    x = rng.standard_normal(shape).astype(np.float32)
    x[rng.random(shape) < 0.8] = 0.0
    if shape[0] > 3:
        x[2] = 0.0                                # an empty row
    rp, ci, va = S.to_sparse_device(torch.from_numpy(x).cuda())
    r, c = np.nonzero(x)
    want_rp = np.zeros(shape[0] + 1, np.int64)
    np.cumsum(np.bincount(r, minlength=shape[0]), out=want_rp[1:])
    assert np.array_equal(rp.cpu().numpy(), want_rp)
    assert np.array_equal(ci.cpu().numpy(), c)
    assert np.array_equal(va.cpu().numpy(), x[r, c])


def test_dense_to_csr_all_zero_and_capacity_retry():
    rp, ci, va = S.to_sparse_device(torch.zeros(10, 12).cuda())
    assert rp.cpu().tolist() == [0] * 11 and ci.numel() == 0 and va.numel() == 0
    x = torch.ones(2048, 1024)                    # 2 M non-zeros: larger than the first capacity guess
    rp, ci, va = S.to_sparse_device(x.cuda())
    assert int(rp[-1]) == 2048 * 1024 and ci.numel() == 2048 * 1024
    assert torch.equal(ci[:1024].cpu(), torch.arange(1024, dtype=torch.int32))


def test_prepared_graph_feeds_the_layer():
    """sym_norm on the GPU -> CSR row pointer -> ADJ stage, all device-resident, against the host pipeline."""
    from sgracex1_b200 import _lib
    from tests.test_dist_cpu import torch_adj
    rng = np.random.default_rng(3)
    n, p = 500, 16
    ei = torch.from_numpy(rng.integers(0, n, size=(2, 3000)).astype(np.int64))
    hei, hnorm = S.sym_norm2(ei, n, fill=1.0)
    dei, dnorm = S.sym_norm2_device(ei.cuda(), n, fill=1.0)
    xw = torch.from_numpy(rng.standard_normal((n, p)).astype(np.float32))
    rp = torch.zeros(n + 1, dtype=torch.int64)
    rp[1:] = torch.cumsum(torch.bincount(hei[0], minlength=n), 0)
    want = torch_adj((rp.int(), hei[1].int(), hnorm), xw, 1).numpy()
    h = _lib.Handle(0)
    h.set_option(_lib.OPT_MODE, _lib.MODE_F32_FAST)
    h.set_option(_lib.OPT_STAGING, 0)
    h.set_option(_lib.OPT_INDEX_FORMAT, 1)        # COO row indices in the rowPtr buffer, as the full design takes them
    d = _lib.LayerDesc()
    rows32, cols32 = dei[0].int().contiguous(), dei[1].int().contiguous()
    out = torch.empty(n, p, device="cuda")
    xwd = xw.cuda()
    d.N_adj, d.M_adj, d.P_w, d.relu, d.nnz_adj = n, n, p, 1, int(cols32.numel())
    d.rowPtr_adj, d.columnIndex_adj, d.values_adj, d.D = rows32.data_ptr(), cols32.data_ptr(), dnorm.data_ptr(), out.data_ptr()
    h.adj_run(d, xwd.data_ptr(), n)
    h.wait()
    U.assert_close_f32(out.cpu().numpy(), want, what="ADJ on the GPU-prepared graph")


@pytest.mark.parametrize("qbits,gat", [(4, 0), (4, 1), (8, 1), (2, 0)])
def test_prune_adjacency_compaction_keeps_the_layer_bit_for_bit(qbits, gat):
    """Adaptive pruning made explicit (sgrace.py:626-629): dropping the zero-coded adjacency entries once leaves D
    unchanged (GCN: bit for bit; GAT: to rounding, logits bit for bit through the `kept` map); the surviving set
    equals the numpy restatement."""
    import torch
    from sgracex1_b200 import _lib, quant as Q
    from sgracex1_b200.driver import DeviceLayer
    from sgracex1_b200.pynq_compat import MmultTop
    from sgracex1_b200 import graphs as G
    p = G.pubmed_shape(seed=5)
    rng = np.random.default_rng(3)
    # a wide spread of edge weights so that a good share of them quantise to code 0
    val = (p.adj_val * rng.choice([1.0, 0.3, 0.05, 0.01], size=len(p.adj_val))).astype(np.float32)
    c = Q.layer_constants(qbits)
    inv_as = np.float32(1.0 / c["a_s"])
    codes = np.clip(np.rint(inv_as * val), 0, 2 ** qbits - 1)
    keep = codes != 0
    assert 0.05 < 1.0 - keep.mean() < 0.95
    ip = MmultTop(0)
    ip.configure(mode=_lib.MODE_FULL, qbits=qbits, staging=0, index_format=0)
    dev = "cuda:0"
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    rp, ci, va = t(p.adj_rowptr, np.int32), t(p.adj_col, np.int32), t(val, np.float32)
    nnz = len(val)
    rp2, ci2, va2, kept = torch.zeros_like(rp), torch.zeros_like(ci), torch.zeros_like(va), torch.zeros_like(ci)
    n2 = ip.handle.prune_adjacency(rp.data_ptr(), ci.data_ptr(), va.data_ptr(), p.N, nnz, float(inv_as), qbits,
                                   rp2.data_ptr(), ci2.data_ptr(), va2.data_ptr(), kept.data_ptr())
    assert n2 == int(keep.sum())
    assert np.array_equal(kept[:n2].cpu().numpy(), np.nonzero(keep)[0])
    assert np.array_equal(ci2[:n2].cpu().numpy(), p.adj_col[keep]) and np.array_equal(va2[:n2].cpu().numpy(), val[keep])
    want_rp = np.concatenate([[0], np.cumsum(np.add.reduceat(keep.astype(np.int64), p.adj_rowptr[:-1]) *
                                             (np.diff(p.adj_rowptr) > 0))]).astype(np.int32)
    assert np.array_equal(rp2.cpu().numpy(), want_rp)
    att = rng.uniform(-0.6, 0.6, size=2 * p.P).astype(np.float32)
    outs = []
    for adj in ((p.adj_rowptr, p.adj_col, val), (rp2.cpu().numpy(), ci2[:n2].cpu().numpy(), va2[:n2].cpu().numpy())):
        dl = DeviceLayer(ip.handle, _lib.MODE_FULL, device=dev)
        dl.load(N=p.N, M=p.M, P=p.P, adj=adj, fea=(p.fea_rowptr, p.fea_col, p.fea_val), B=p.B, relu=1, attention=att,
                gat_mode=gat, consts=c, want_es=True)
        dl.run()
        outs.append((dl.result("D"), dl.result("E"), dl.result("S")))
    (d0, e0, s0), (d1, e1, s1) = outs
    if not gat:
        # the GCN sum of a row runs over the survivors in their order either way: identical bits
        assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
    else:
        # the softmax sums are grouped by the position of an edge in its row, which compaction changes: the logits
        # are identical, weights and output agree to rounding
        k = np.nonzero(keep)[0]
        assert np.array_equal(e0[k].view(np.uint32), e1[:n2].view(np.uint32))
        np.testing.assert_allclose(s1[:n2], s0[k], rtol=2e-6, atol=1e-9)
        U.assert_close_f32(d1, d0, rtol=2e-6, what="GAT on the compacted adjacency")
        assert not e0[~keep].any() and not s0[~keep].any()
