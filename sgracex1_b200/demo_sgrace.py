"""Host mirror of the reference's demo model (demo/emulation/demo_sgrace.py) on libsgrace_b200.

    GAT_PYNQ        :271-401   two SGRACE layers (att2: sparse features + fused ReLU, conv22: dense features),
                               Relu_SGRACE between them, dropout, Linear read-out; sym_norm2 per forward
    train / test    :476-560   the script's loops, as functions of (model, loader, ...)

The reference class reads the module globals `dataset`, `average_node_degree`, `config`; here the three numbers it
takes from them are constructor arguments.  Parameter names -- and therefore state_dict keys -- are the reference's
(att2.weight / att2.attention / att2.bias / conv22.* / lin.*: the keys of demo/zcu104/model_Photo_8bit.ptx), so a
checkpoint written by the reference loads with load_state_dict and vice versa.
"""
from __future__ import annotations

import math
import time

import torch
import torch.nn.functional as F
from torch.nn import Linear

from . import config
from . import sgrace
from .sgrace import GATConv_SGRACE, Relu_SGRACE


class GAT_PYNQ(torch.nn.Module):
    def __init__(self, hidden_channels, head_count, num_node_features, num_classes, average_node_degree):
        super().__init__()
        self.att2 = GATConv_SGRACE(num_node_features, hidden_channels, head_count, dropout=0.1, alpha=0.2, concat=False)
        self.conv22 = GATConv_SGRACE(hidden_channels * head_count, hidden_channels, 1)
        self.reluh = Relu_SGRACE()
        self.lin = Linear(hidden_channels, num_classes)
        # demo_sgrace.py:246 -- self-loop weight = trunc(log2(average degree)) (the script computes it from `data`)
        self.fill_value = math.trunc(math.log2(average_node_degree))

    def forward(self, x, edge_index):
        ptime = time.time()
        edge_index, norm = sgrace.sym_norm2(edge_index, x.size(0), fill=self.fill_value)
        adj = torch.sparse_coo_tensor(edge_index, norm, (x.size(0), x.size(0)))
        # layer 1: sparse features, ReLU merged into the accelerator call (:330); Relu_SGRACE only masks the gradient
        x = self.att2(config.compute_attention, 0, 1, x, edge_index, norm, adj)
        x = self.reluh(x)
        # layer 2: dense features (gemm_mode 1), no ReLU (:362-376)
        x = self.conv22(config.compute_attention, 1, 0, x, edge_index, norm, adj)
        x = x.float()
        x = F.dropout(x, p=0.5, training=self.training)
        x = self.lin(x)
        if config.profiling == 1:
            print("Model time {:.5f}ms".format(1000 * (time.time() - ptime)))
        return x


def train(model, loader, optimizer, criterion):
    """demo_sgrace.py:476-507: one pass over the loader (batches carry x, edge_index, y, train_mask)."""
    model.train()
    for batch in loader:
        out = model(batch.x, batch.edge_index)
        loss = criterion(out[batch.train_mask], batch.y[batch.train_mask])
        loss.backward()
        optimizer.step()
        optimizer.zero_grad()


@torch.no_grad()
def test(model, loader, split):
    """demo_sgrace.py:509-560: accuracy over the batches' `split`_mask nodes."""
    model.eval()
    correct = total = 0
    for batch in loader:
        pred = model(batch.x, batch.edge_index).argmax(dim=1)
        mask = getattr(batch, f"{split}_mask")
        correct += int((pred[mask] == batch.y[mask]).sum())
        total += int(mask.sum())
    return correct / max(total, 1)
