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
    out_g = m_gpu(d.x.to(dev), adj_dev, d.batch.to(dev), 40)
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
