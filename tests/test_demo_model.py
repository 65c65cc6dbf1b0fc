"""the reference's demo model (demo/emulation/demo_sgrace.py:271-401, GAT_PYNQ) on the reference's own
Cora files.  tests/golden/demo_model_cora.npz was produced by the reference's unmodified sgrace.py pieces in the
emulation mode (make_golden.py demo)."""
import os

import numpy as np
import pytest
import torch

from tests import util as U


def _fixture():
    g = np.load(os.path.join(U.GOLDEN, "demo_model_cora.npz"))
    n, m = 2708, 1433
    x = np.zeros((n, m), np.float32)
    x[np.repeat(np.arange(n), np.diff(g["fea_rowptr"])), g["fea_col"].astype(np.int64)] = 1.0
    return g, n, m, x


def test_state_dict_keys_are_the_reference_checkpoint_keys():
    """demo/zcu104/model_Photo_8bit.ptx is a state_dict of the reference model: same keys, same shapes rule"""
    from sgracex1_b200 import config
    from sgracex1_b200.demo_sgrace import GAT_PYNQ
    g, n, m, _ = _fixture()
    acc0 = config.acc
    config.acc = 0                      # constructing the modules needs no accelerator
    try:
        model = GAT_PYNQ(16, 1, 745, 8, 31.8)           # the Photo checkpoint: 745 features, 8 classes
    finally:
        config.acc = acc0
    sd = model.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["state_dict_keys"]]
    assert tuple(sd["att2.weight"].shape) == (745, 16) and tuple(sd["att2.attention"].shape) == (32, 1)
    assert tuple(sd["conv22.weight"].shape) == (16, 16) and tuple(sd["lin.weight"].shape) == (8, 16)


@pytest.mark.gpu
@pytest.mark.parametrize("qbits,gat", [(8, 0), (8, 1), (4, 0), (4, 1)])
def test_gat_pynq_forward_on_real_cora_matches_reference_emulation(qbits, gat):
    from sgracex1_b200 import config, sgrace as S
    from sgracex1_b200.demo_sgrace import GAT_PYNQ
    g, n, m, x = _fixture()
    ei = torch.from_numpy(g["edge_index"].astype(np.int64))
    config.w_qbits, config.compute_attention, config.acc, config.accb = qbits, gat, 1, 0
    config.N_adj, config.P_w, config.NNZ_adj, config.NNZ_fea = n, 16, ei.shape[1] + n + 8, n * 16 + int(g["fea_rowptr"][-1]) + 8
    S.init_SGRACE()
    try:
        model = GAT_PYNQ(16, 1, m, 7, float(g["average_node_degree"]))
        sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("p_")}
        missing = model.load_state_dict(sd, strict=False)
        assert set(missing.missing_keys) <= {"att2.bias", "conv22.bias"} and not missing.unexpected_keys
        model.eval()
        with torch.no_grad():
            logits = model(torch.from_numpy(x), ei).numpy()
        want = g[f"q{qbits}_gat{gat}_logits"]
        # two quantised layers: a last-bit difference in layer 1 can move a layer-2 feature code by one step, so the bar
        # is 1e-4 of the logit range on every element and 1e-5 relative on all but a handful
        scale = np.abs(want).max()
        err = np.abs(logits - want)
        assert err.max() <= 2e-3 * scale, (err.max(), scale)
        assert (err > 1e-5 * np.maximum(np.abs(want), scale)).mean() < 0.01
        assert (logits.argmax(1) == want.argmax(1)).mean() > 0.999
    finally:
        S.free_SGRACE()
