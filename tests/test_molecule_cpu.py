"""CPU tests of the molecule-GCN host mirror (jupyter/molecule_gcn/Graph_Classification.ipynb):
the reference's class names exist with the reference's signatures, and the torch_geometric
stand-ins compute what PyG computes."""
import inspect

import numpy as np
import torch

from sgracex1_b200 import molecule_gcn as MG


def test_reference_names_and_signatures():
    for name in ("RPYNQ", "FPYNQ", "Relu_pynq", "GraphConvolution_pynq", "GCN_PYNQ"):
        assert hasattr(MG, name)
    sig = list(inspect.signature(MG.GraphConvolution_pynq.forward).parameters)
    assert sig == ["self", "acc", "dense", "relu", "input", "adj", "rowPtr_fea_buffer", "columnIndex_fea_buffer",
                   "values_fea_buffer", "rowPtr_adj_buffer", "columnIndex_adj_buffer", "values_adj_buffer", "B_buffer",
                   "D_buffer"]
    sig = list(inspect.signature(MG.GCN_PYNQ.forward).parameters)
    assert sig[:5] == ["self", "acc", "x", "edge_index", "batch"] and sig[-1] == "D_buffer"
    assert list(inspect.signature(MG.FPYNQ.forward).parameters) == ["ctx", "my_ip", "adj", "input", "weights"]


def test_rpynq_masks_where_the_accelerator_output_is_zero():
    x = torch.tensor([[0.0, 1.5, -2.0], [0.0, 0.0, 3.0]], requires_grad=True)
    y = MG.RPYNQ.apply(x)
    assert torch.equal(y, x)
    y.backward(torch.ones_like(x))
    assert torch.equal(x.grad, torch.tensor([[0.0, 1.0, 1.0], [0.0, 0.0, 1.0]]))


def test_pyg_stand_ins():
    ei = torch.tensor([[0, 1, 1, 2, 2, 0, 0], [1, 0, 2, 1, 0, 2, 1]])        # edge (0,1) twice
    adj = MG.to_dense_adj(ei, 4)
    want = torch.zeros(4, 4)
    for a, b in ei.t().tolist():
        want[a, b] += 1
    assert torch.equal(adj, want)
    rp, ci, va = MG.edge_index_to_csr(ei, 4)
    csr = adj.to_sparse_csr()
    assert np.array_equal(rp, csr.crow_indices().numpy()) and np.array_equal(ci, csr.col_indices().numpy())
    assert np.array_equal(va, csr.values().numpy())
    x = torch.arange(12, dtype=torch.float32).reshape(6, 2)
    batch = torch.tensor([0, 0, 1, 1, 1, 3])
    got = MG.global_mean_pool(x, batch, 4)
    assert torch.allclose(got[0], x[:2].mean(0)) and torch.allclose(got[1], x[2:5].mean(0))
    assert torch.equal(got[2], torch.zeros(2)) and torch.allclose(got[3], x[5])


def test_graphconvolution_cpu_branch_matches_reference_formula():
    torch.manual_seed(0)
    conv = MG.GraphConvolution_pynq(5, 3, None)
    stdv = 1.0 / np.sqrt(3)
    assert float(conv.weight.abs().max()) <= stdv + 1e-6
    x, adj = torch.rand(4, 5), torch.rand(4, 4)
    out = conv(0, 0, 0, x, adj, *([None] * 8))
    assert torch.allclose(out, adj @ x @ conv.weight)
