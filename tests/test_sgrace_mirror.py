"""Tests of the full-design driver mirror (sgracex1_b200/sgrace.py vs the reference's
demo/sgrace_lib/sgrace.py): CPU part = names, signatures, sym_norm2 against the fixture produced by
the reference's own sym_norm2; GPU part = GATConv_SGRACE through the register map against the
fixtures produced by the reference's own emulation."""
import inspect
import os

import numpy as np
import pytest
import torch

from tests import util as U


def test_reference_names_and_signatures():
    from sgracex1_b200 import sgrace as S
    for name in ("sym_norm2", "quantization", "quantization_b", "quantization_uqbits", "quantization_qbits",
                 "generate_quantization_constants", "generate_quantization_uqbits_constants",
                 "generate_quantization_qbits_constants", "fake_quantization", "fake_quantization_b",
                 "quantization_fbits", "quantization_ufbits", "RPYNQ", "FPYNQ_GAT", "Relu_SGRACE", "GATConv_SGRACE",
                 "init_SGRACE"):
        assert hasattr(S, name), name
    assert list(inspect.signature(S.GATConv_SGRACE.forward).parameters) == [
        "self", "compute_attention", "dense", "relu", "input", "edge_index", "norm", "adj"]
    assert list(inspect.signature(S.FPYNQ_GAT.forward).parameters) == [
        "ctx", "my_ip", "self", "adj", "nnz_adj", "input", "weights", "attention", "out_features", "dropout", "relu"]
    assert list(inspect.signature(S.GATConv_SGRACE.__init__).parameters) == [
        "self", "in_features", "out_features", "nheads", "bias", "dropout", "alpha", "concat"]
    from sgracex1_b200 import config
    for flag in ("acc", "accb", "fake_quantization", "compute_attention", "w_qbits", "hidden_channels", "N_adj",
                 "NNZ_adj", "NNZ_fea", "P_w", "float_type", "device", "profiling", "min_output"):
        assert hasattr(config, flag), flag


def test_sym_norm2_matches_the_reference_fixture():
    from sgracex1_b200 import sgrace as S
    g = np.load(os.path.join(U.GOLDEN, "sym_norm2.npz"))
    ei, norm = S.sym_norm2(torch.from_numpy(g["edge_index_in"]), int(g["num_nodes"]), fill=float(g["fill"]))
    assert np.array_equal(ei.numpy(), g["edge_index_out"])
    np.testing.assert_allclose(norm.numpy(), g["norm_out"], rtol=1e-6, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("qbits,gat", [(8, 0), (4, 0), (8, 1), (4, 1)])
def test_gatconv_sgrace_against_reference_emulation(qbits, gat):
    from sgracex1_b200 import config, sgrace as S
    g = np.load(os.path.join(U.GOLDEN, f"qlayer_q{qbits}_gat{gat}.npz"))
    x, w, att = g["x"], g["w"], g["attention"]
    n, m = x.shape
    p = w.shape[1]
    config.w_qbits, config.compute_attention, config.acc, config.accb = qbits, gat, 1, 0
    config.N_adj, config.P_w, config.NNZ_adj, config.NNZ_fea = n, p, len(g["norm"]) + 8, n * m + 8
    S.init_SGRACE()
    try:
        layer = S.GATConv_SGRACE(m, p)
        with torch.no_grad():
            layer.weight.copy_(torch.from_numpy(w))
            layer.attention.copy_(torch.from_numpy(att.reshape(-1, 1)))
        ei = torch.from_numpy(g["edge_index"].astype(np.int64))
        norm = torch.from_numpy(g["norm"])
        adj = torch.sparse_coo_tensor(ei, norm, (n, n))
        xt = torch.from_numpy(x)
        for relu in (0, 1):
            for dense in (0, 1):
                out = layer(gat, dense, relu, xt, ei, norm, adj)
                U.assert_close_f32(out.detach().numpy(), g[f"out_relu{relu}_dense{dense}"],
                                   what=f"GATConv_SGRACE q={qbits} gat={gat} relu={relu} dense={dense}")
        # the software backward (accb = 0): GCN formulas grad_X = A g W^T, grad_W = X^T A g  (sgrace.py:1100-1103)
        xt2 = xt.clone().requires_grad_()
        out = layer(gat, 1, 0, xt2, ei, norm, adj)
        gout = torch.ones_like(out)
        out.backward(gout)
        if not gat:
            A = adj.to_dense()
            np.testing.assert_allclose(layer.weight.grad.numpy(), (xt.t() @ (A @ gout)).numpy(), rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(xt2.grad.numpy(), ((A @ gout) @ layer.weight.detach().t()).numpy(), rtol=1e-5, atol=1e-6)
        else:
            assert layer.attention.grad is not None and torch.isfinite(layer.attention.grad).all()
        assert S.cur_max_fea > 0
    finally:
        S.free_SGRACE()


@pytest.mark.parametrize("gat", [0, 1])
def test_software_backward_matches_the_reference_emulation(gat):
    """FPYNQ_GAT.backward with accb = 0 (sgrace.py:880-1110): the tensors the reference's forward saved and a seeded
    grad_output go through the mirror's backward; grad_input, grad_weights and the attention gradient equal the ones the
    imported reference returned (tests/golden/make_golden.py:backward_emulation).  CPU only: this is host code."""
    import types
    from sgracex1_b200 import config, sgrace as S
    g = np.load(os.path.join(U.GOLDEN, "backward_emulation.npz"))
    n = int(g["n"])
    adj = torch.sparse_coo_tensor(torch.from_numpy(g["edge_index"].astype(np.int64)), torch.from_numpy(g["norm"]), (n, n))
    t = lambda k: torch.from_numpy(g[f"gat{gat}_{k}"])
    ctx = types.SimpleNamespace(saved_tensors=(adj, t("input"), t("weights"), t("e"), t("attentions"), t("output")),
                                alpha=float(g[f"gat{gat}_alpha"]), nheads=1)
    old = (config.accb, config.compute_attention)
    config.accb, config.compute_attention = 0, gat
    try:
        grads = S.FPYNQ_GAT.backward(ctx, t("grad_output"))
    finally:
        config.accb, config.compute_attention = old
    grad_input, grad_weights, grad_attention = grads[4], grads[5], grads[6]
    for name, got in (("grad_input", grad_input), ("grad_weights", grad_weights), ("grad_attention", grad_attention)):
        want = g[f"gat{gat}_{name}"]
        got = got.detach().numpy().reshape(want.shape)
        scale = max(float(np.abs(want).max()), 1e-30)
        assert np.abs(got - want).max() <= 1e-5 * scale, (name, float(np.abs(got - want).max()), scale)
