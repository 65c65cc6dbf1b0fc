"""GPU tests of the molecule-GCN drop-in (the reference notebook's layer code on the B200 backend)
and of the multi-GPU drivers at world size 1."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from sgracex1_b200 import _lib, dist as sdist, graphs as G
from sgracex1_b200 import molecule_gcn as MG
from sgracex1_b200.pynq_compat import Overlay
from tests import util as U
from tests.test_dist_cpu import torch_adj, torch_layer

pytestmark = pytest.mark.gpu


class Data:
    pass


def molecule_data(n_graphs=24, seed=3):
    prob, batch, y = G.molecule_batch(n_graphs=n_graphs, seed=seed, P=16)
    d = Data()
    d.x = torch.zeros(prob.N, prob.M)
    d.x[torch.arange(prob.N), torch.from_numpy(prob.fea_col.astype(np.int64))] = 1.0
    rows = np.repeat(np.arange(prob.N), np.diff(prob.adj_rowptr))
    d.edge_index = torch.from_numpy(np.stack([rows, prob.adj_col]).astype(np.int64))
    d.batch = torch.from_numpy(batch)
    d.y = torch.from_numpy(y.astype(np.int64))
    return prob, d


def test_fpynq_half_buffers_bit_exact_with_csim_oracle():
    prob, d = molecule_data()
    ol = Overlay("gnn_all.bit")
    ip = ol.mmult_top_0
    bufs = MG.NotebookBuffers(ip, prob.N, prob.nnz_adj, prob.N * 16, 16, dtype=np.float16)
    try:
        model = MG.GCN_PYNQ(16, ip)
        adj = MG.to_dense_adj(d.edge_index, prob.N)
        csr = adj.to_sparse_csr()
        bufs.rowPtr_adj_buffer[:prob.N + 1] = csr.crow_indices().numpy()
        bufs.columnIndex_adj_buffer[:prob.nnz_adj] = csr.col_indices().numpy()
        bufs.values_adj_buffer[:prob.nnz_adj] = csr.values().numpy().astype(np.float16)
        xc = d.x.to_sparse_csr()
        bufs.rowPtr_fea_buffer[:prob.N + 1] = xc.crow_indices().numpy()
        bufs.columnIndex_fea_buffer[:prob.N] = xc.col_indices().numpy()
        bufs.values_fea_buffer[:prob.N] = xc.values().numpy().astype(np.float16)
        out = model.conv1(1, 0, 1, d.x, adj, *bufs.as_args())
        W = model.conv1.weight.detach().numpy()
        B16 = O.to_storage(O.weights_to_B(W), O.F16)
        a16 = (prob.adj_rowptr, prob.adj_col, O.to_storage(prob.adj_val, O.F16))
        f16 = (prob.fea_rowptr, prob.fea_col, O.to_storage(prob.fea_val, O.F16))
        ref = O.layer(dtype=O.F16, N=prob.N, M_fea=prob.M, P=16, adj=a16, fea=f16, B=B16, relu=1)
        assert out.dtype == torch.float16
        assert np.array_equal(out.detach().numpy().view(np.uint16), ref)
    finally:
        bufs.free()


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_gcn_pynq_forward_backward_matches_torch_reference(dtype):
    prob, d = molecule_data(n_graphs=20, seed=5)
    ol = Overlay("gnn_all.bit")
    ip = ol.mmult_top_0
    bufs = MG.NotebookBuffers(ip, prob.N, prob.nnz_adj, prob.N * 16, 16, dtype=dtype)
    try:
        model = MG.GCN_PYNQ(16, ip)
        model.eval()
        out = model(0, d.x, d.edge_index, d.batch, *bufs.as_args())
        loss = torch.nn.CrossEntropyLoss()(out, d.y)
        loss.backward()
        got = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        # plain torch model with the same weights: adj @ x @ W, relu, adj @ h @ W2, mean pool, linear
        adj = MG.to_dense_adj(d.edge_index, prob.N)
        W1 = model.conv1.weight.detach().clone().requires_grad_()
        W2 = model.conv2.weight.detach().clone().requires_grad_()
        h = (adj @ d.x @ W1).relu()
        h = adj @ h @ W2
        ref = model.lin(MG.global_mean_pool(h, d.batch))
        tol = 1e-4 if dtype == np.float32 else 3e-2
        assert torch.allclose(out, ref, rtol=tol, atol=tol), float((out - ref).abs().max())
        ref_loss = torch.nn.CrossEntropyLoss()(ref, d.y)
        g1, g2 = torch.autograd.grad(ref_loss, [W1, W2])
        assert torch.allclose(got["conv1.weight"], g1, rtol=10 * tol, atol=tol)
        assert torch.allclose(got["conv2.weight"], g2, rtol=10 * tol, atol=tol)
    finally:
        bufs.free()


def test_gcn_b200_device_resident_matches_stand_in():
    prob, d = molecule_data(n_graphs=40, seed=9)
    dev = torch.device("cuda:0")
    handle = _lib.Handle(0)
    handle.set_option(_lib.OPT_STAGING, 0)
    adj_dev = tuple(torch.from_numpy(a).to(dev) for a in (prob.adj_rowptr, prob.adj_col, prob.adj_val))
    adj_cpu = tuple(torch.from_numpy(a) for a in (prob.adj_rowptr, prob.adj_col, prob.adj_val))
    m_gpu = MG.GCN_B200(16, handle).to(dev).eval()
    m_cpu = MG.GCN_B200(16, None, layer_fn=torch_layer).eval()
    m_cpu.load_state_dict({k: v.cpu() for k, v in m_gpu.state_dict().items()})
    pool = MG.pooling_csr(d.batch.to(dev), 40)
    out_g = m_gpu(d.x.to(dev), adj_dev, d.batch.to(dev), 40, pool)
    out_c = m_cpu(d.x, adj_cpu, d.batch, 40)
    assert torch.allclose(out_g.cpu(), out_c, rtol=1e-4, atol=1e-5)
    torch.nn.CrossEntropyLoss()(out_g, d.y.to(dev)).backward()
    torch.nn.CrossEntropyLoss()(out_c, d.y).backward()
    for (n, pg), (_, pc) in zip(m_gpu.named_parameters(), m_cpu.named_parameters()):
        if pc.grad is not None:
            assert torch.allclose(pg.grad.cpu(), pc.grad, rtol=1e-3, atol=1e-5), n


def test_row_partitioned_layer_world1_and_stage_split():
    pr = U.random_problem(11, n=1000, m=40, p=64)
    dev = torch.device("cuda:0")
    handle = _lib.Handle(0)
    handle.set_option(_lib.OPT_STAGING, 0)
    handle.set_option(_lib.OPT_MODE, _lib.MODE_F32_FAST)
    layer = sdist.RowPartitionedLayer(1000, 64, 0, 1, dev, sdist.abi_fea_fn(handle), sdist.abi_adj_fn(handle))
    adj = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in pr["adj"])
    D = layer.forward(torch.from_numpy(pr["x"]).to(dev), torch.from_numpy(pr["W"]).to(dev), adj, 1)
    handle.wait()
    adj_c = tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in pr["adj"])
    want = torch_adj(adj_c, torch.from_numpy(pr["x"] @ pr["W"]), 1)
    U.assert_close_f32(D.cpu().numpy(), want.numpy(), what="row-partitioned layer")
    # two emulated ranks on one GPU: slices of A against the full XW give the rows of the full result
    xw = torch.from_numpy(pr["x"] @ pr["W"]).to(dev)
    rp, ci, va = pr["adj"]
    for r in range(2):
        lo, hi = sdist.row_range(1000, r, 2)
        loc = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in sdist.csr_row_slice(rp, ci, va, lo, hi))
        part = sdist.abi_adj_fn(handle)(loc, xw, 1)
        handle.wait()
        U.assert_close_f32(part.cpu().numpy(), want.numpy()[lo:hi], what=f"rank {r} rows")


def test_peer_gather_layer_two_emulated_ranks_on_one_gpu():
    """The NVLink peer-gather ADJ with its row-partitioned operand split over two buffers of the same
    GPU (what two ranks would map from each other): every rank's rows must equal the single-GPU layer."""
    pr = U.random_problem(17, n=1501, m=100, p=256, avg_deg=6)
    rp, ci, va = pr["adj"]
    # one very long row so that the segmented long-row kernel runs with the peer table too
    rng = np.random.default_rng(3)
    deg = np.diff(rp).copy()
    rows = [ci[rp[r]:rp[r + 1]] for r in range(1501)]
    rows[700] = np.arange(1501, dtype=np.int32)
    deg[700] = 1501
    rp = np.zeros(1502, np.int32)
    np.cumsum(deg, out=rp[1:])
    ci = np.concatenate(rows).astype(np.int32)
    va = rng.uniform(-0.2, 0.2, size=len(ci)).astype(np.float32)
    dev = torch.device("cuda:0")
    handle = _lib.Handle(0)
    handle.set_option(_lib.OPT_STAGING, 0)
    handle.set_option(_lib.OPT_MODE, _lib.MODE_F32_FAST)
    x, W = pr["x"], pr["W"]
    adj_c = (torch.from_numpy(rp), torch.from_numpy(ci), torch.from_numpy(va))
    want = torch_adj(adj_c, torch.from_numpy(x @ W), 1).numpy()
    world = 2
    lay = sdist.PeerGatherLayer(handle, 1501, 100, 0, 1, dev)     # re-pointed at the two emulated ranks below
    block = sdist.row_block(1501, world)
    bufs = []
    for r in range(world):
        addr, _ = handle.peer_alloc(block * 100 * 4)
        t = torch.as_tensor(sdist._RawCuda(addr, (block, 100)), device=dev)
        lo, hi = sdist.row_range(1501, r, world)
        t.zero_()
        t[:hi - lo] = torch.from_numpy(x[lo:hi]).to(dev)
        bufs.append((addr, t))
    torch.cuda.synchronize()
    Wd = torch.from_numpy(W).to(dev)
    for r in range(world):
        lo, hi = sdist.row_range(1501, r, world)
        loc = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in sdist.csr_row_slice(rp, ci, va, lo, hi))
        lay.world, lay.block, lay.bases = world, block, [b[0] for b in bufs]
        out, keep = lay.forward_agg_first(loc, Wd, 1)
        handle.wait()
        U.assert_close_f32(out.cpu().numpy(), want[lo:hi], what=f"peer-gather rank {r}")
    handle.peer_release()


def test_halo_layer_three_emulated_ranks_on_one_gpu():
    """Halo exchange + split aggregation + tensor-core dense stage: every emulated rank's rows must
    equal the single-GPU layer act(A (X W))."""
    n, m, p = 3000, 100, 256
    pr = U.random_problem(23, n=n, m=m, p=p, avg_deg=8)
    rp, ci, va = pr["adj"]
    rows = [ci[rp[r]:rp[r + 1]] for r in range(n)]
    rows[1234] = np.arange(0, n, 2, dtype=np.int32)          # a hub row with many remote columns (long-row path)
    deg = np.array([len(r) for r in rows])
    rp = np.zeros(n + 1, np.int32)
    np.cumsum(deg, out=rp[1:])
    ci = np.concatenate(rows).astype(np.int32)
    va = np.random.default_rng(5).uniform(-0.2, 0.2, size=len(ci)).astype(np.float32)
    x, W = pr["x"], pr["W"]
    adj_c = (torch.from_numpy(rp), torch.from_numpy(ci), torch.from_numpy(va))
    want = torch_adj(adj_c, torch.from_numpy(x @ W), 1).numpy()
    dev = torch.device("cuda:0")
    world = 3
    hm, hh = _lib.Handle(0), _lib.Handle(0)
    for h in (hm, hh):
        h.set_option(_lib.OPT_STAGING, 0)
        h.set_option(_lib.OPT_MODE, _lib.MODE_F32_FAST)
    hm.set_stream(torch.cuda.current_stream().cuda_stream or 1)
    layers = []
    for r in range(world):
        lo, hi = sdist.row_range(n, r, world)
        loc = sdist.csr_row_slice(rp, ci, va, lo, hi)
        lay = sdist.HaloLayer(hm, hh, loc, n, m, r, world, dev, exchange="defer")
        lay.local[:hi - lo].copy_(torch.from_numpy(x[lo:hi]).to(dev))
        layers.append(lay)
    for lay in layers:
        lay.bases = [l2.addr for l2 in layers]
    torch.cuda.synchronize()
    Wd = torch.from_numpy(W).to(dev)
    for r, lay in enumerate(layers):
        out, keep = lay.forward(Wd, 1)
        torch.cuda.synchronize()
        lo, hi = sdist.row_range(n, r, world)
        assert lay.n_halo > 0 and lay.nnz_remote > 0
        U.assert_close_f32(out.cpu().numpy(), want[lo:hi], what=f"halo layer rank {r}")
    hm.peer_release()


@pytest.mark.parametrize("exchange", ["dma", "pushf"])
def test_halo_layer_copy_engine_exchange_three_emulated_ranks(monkeypatch, exchange):
    """The exchanges that signal with flag words and stream waits -- "dma": pack, copy-engine transfers (no SM);
    "pushf": SM push straight into the peers' halo regions -- with the owned columns aggregated while the halo
    travels: three emulated ranks on one GPU, each with its own streams and handles, two layers back to back (the
    flags carry an epoch)."""
    monkeypatch.setenv("SGRACE_HALO_EXCHANGE", exchange)
    n, m, p = 3000, 100, 256
    pr = U.random_problem(29, n=n, m=m, p=p, avg_deg=8)
    rp, ci, va = pr["adj"]
    x, W = pr["x"], pr["W"]
    adj_c = (torch.from_numpy(rp), torch.from_numpy(ci), torch.from_numpy(va))
    dev = torch.device("cuda:0")
    world = 3
    layers, streams, handles = [], [], []
    for r in range(world):
        st = torch.cuda.Stream()
        hm, hh = _lib.Handle(0), _lib.Handle(0)
        for h in (hm, hh):
            h.set_option(_lib.OPT_STAGING, 0)
            h.set_option(_lib.OPT_MODE, _lib.MODE_F32_FAST)
        hm.set_stream(st.cuda_stream)
        lo, hi = sdist.row_range(n, r, world)
        loc = sdist.csr_row_slice(rp, ci, va, lo, hi)
        with torch.cuda.stream(st):
            lay = sdist.HaloLayer(hm, hh, loc, n, m, r, world, dev, exchange="defer")
        layers.append(lay); streams.append(st); handles.append((hm, hh))
    wants = [lay._halo_wants(lay.halo_rows_np)[1] for lay in layers]
    for lay in layers:
        lay.bases = [l2.addr for l2 in layers]
        lay.flag_bases = [l2.flag_addr for l2 in layers]
        lay._exchange_push_lists(lay.halo_rows_np, dev, gather=lambda w: wants)
    Wd = torch.from_numpy(W).to(dev)
    torch.cuda.synchronize()
    for rep in range(2):
        xs = x * (rep + 1)
        want = torch_adj(adj_c, torch.from_numpy(xs @ W), 1).numpy()
        for r, lay in enumerate(layers):
            lo, hi = sdist.row_range(n, r, world)
            lay.local[:hi - lo].copy_(torch.from_numpy(xs[lo:hi]).to(dev))
        torch.cuda.synchronize()
        outs, states = [], []
        for r, lay in enumerate(layers):          # every rank sends and signals ...
            with torch.cuda.stream(streams[r]):
                states.append(lay.forward_begin())
        for r, lay in enumerate(layers):          # ... before any rank's stream starts waiting for its flags
            with torch.cuda.stream(streams[r]):
                outs.append(lay.forward_end(states[r], Wd, 1)[0])
        torch.cuda.synchronize()
        for r, out in enumerate(outs):
            lo, hi = sdist.row_range(n, r, world)
            U.assert_close_f32(out.cpu().numpy(), want[lo:hi], what=f"dma halo layer rank {r} rep {rep}")
    for hm, _ in handles:
        hm.peer_release()


def test_quantised_gat_layer_row_partition_two_emulated_ranks():
    """SURVEY 8e for the full design: FEA on each rank's rows into its slot of the gathered Wh, then the GAT
    stage on the local adjacency rows against the full Wh (the scores are recomputed from it).  Per-row work
    is deterministic, so every rank's D / E / S must equal the one-GPU layer bit for bit."""
    from oracle import oracle as O
    from sgracex1_b200 import quant as Q
    n, m, p, world = 600, 48, 16, 2
    pr = U.random_problem(31, n=n, m=m, p=p, avg_deg=5)
    rp, ci, va = pr["adj"]
    va = np.abs(va).astype(np.float32)
    frp, fci, fva = pr["fea"]
    fva = np.abs(fva).astype(np.float32)
    c = Q.layer_constants(8)
    att = np.random.default_rng(3).uniform(-0.5, 0.5, size=2 * p).astype(np.float32)
    dev = torch.device("cuda:0")
    h = _lib.Handle(0)
    for k, v in ((_lib.OPT_MODE, _lib.MODE_FULL), (_lib.OPT_QBITS, 8), (_lib.OPT_STAGING, 0), (_lib.OPT_INDEX_FORMAT, 0)):
        h.set_option(k, v)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    Bd, attd = up(O.weights_to_B(pr["W"]).astype(np.float32)), up(att)

    def desc(fea, adj, n_rows, D, E, S, XW=None):
        d = _lib.LayerDesc()
        d.gemm_mode, d.relu, d.gat_mode = 0, 1, 1
        d.N_adj, d.M_adj, d.M_fea, d.P_w = n_rows, n, m, p
        d.scale_fea, d.internal_quantization = c["scale_fea"], c["internal_quantization"]
        d.qscale_fea, d.qscale_w, d.qscale_adj, d.deq_factor = 1 / c["f_s"], 1 / c["w_s"], 1 / c["a_s"], c["deq_o"]
        d.rowPtr_fea, d.columnIndex_fea, d.values_fea = (t.data_ptr() for t in fea)
        d.rowPtr_adj, d.columnIndex_adj, d.values_adj = (t.data_ptr() for t in adj)
        d.nnz_fea, d.nnz_adj = int(fea[1].numel()), int(adj[1].numel())
        d.B, d.attention, d.D, d.E, d.S = Bd.data_ptr(), attd.data_ptr(), D.data_ptr(), E.data_ptr(), S.data_ptr()
        if XW is not None:
            d.XW = XW.data_ptr()
        return d

    fea_full, adj_full = tuple(up(a) for a in (frp, fci, fva)), tuple(up(a) for a in (rp, ci, va))
    D0, E0, S0 = torch.empty(n, p, device=dev), torch.empty(len(ci), device=dev), torch.empty(len(ci), device=dev)
    h.layer_run(desc(fea_full, adj_full, n, D0, E0, S0))
    h.wait()
    Wh = torch.zeros(n, p, device=dev)            # the all-gathered buffer; each rank fills its slot
    parts = []
    for r in range(world):
        lo, hi = sdist.row_range(n, r, world)
        fea_loc = tuple(up(a) for a in sdist.csr_row_slice(frp, fci, fva, lo, hi))
        adj_loc = tuple(up(a) for a in sdist.csr_row_slice(rp, ci, va, lo, hi))
        D, E, S = torch.empty(hi - lo, p, device=dev), torch.empty(adj_loc[1].numel(), device=dev), torch.empty(adj_loc[1].numel(), device=dev)
        d = desc(fea_loc, adj_loc, hi - lo, D, E, S)
        h.fea_run(d, Wh.data_ptr() + lo * p * 4)
        parts.append((lo, hi, d, D, E, S, fea_loc, adj_loc))
    for lo, hi, d, D, E, S, _, _ in parts:          # after the "all-gather": every slot is filled
        h.set_option(_lib.OPT_ROW_OFFSET, lo)       # the destination row's own score is read at its global index
        h.adj_run(d, Wh.data_ptr(), n)
    h.set_option(_lib.OPT_ROW_OFFSET, 0)
    h.wait()
    for lo, hi, d, D, E, S, _, _ in parts:
        assert torch.equal(D, D0[lo:hi]), f"rows {lo}:{hi}"
        assert torch.equal(E, E0[rp[lo]:rp[hi]]) and torch.equal(S, S0[rp[lo]:rp[hi]])
    assert float(D0.abs().sum()) > 0
