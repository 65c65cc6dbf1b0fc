"""Host mirror of the reference's full-design driver `demo/sgrace_lib/sgrace.py` on libsgrace_b200.

Same public names and argument meaning as the reference (line numbers are the reference's):
    sym_norm2                    :18-51     D^-1/2 (A + fill I) D^-1/2 as sorted COO
    quantization* / fake_*       :53-265    (re-exported from .quant)
    RPYNQ                        :267-294
    FPYNQ_GAT                    :298-1126  forward = accelerator path (:321-559), backward = the
                                            software backward (:886-1126, `accb = 0`)
    Relu_SGRACE, GATConv_SGRACE  :1146-1265
    init_SGRACE                  :1271-1896 Overlay + buffers at config maxima + per-bit-width constants
The module keeps the reference's habit of module-level globals (`my_ip`, the `*_buffer`s, the
per-layer constants, `layern` toggling 1 <-> 2 per call) so that scripts written against the
reference (demo/emulation/demo_sgrace.py) find what they expect.  torch_geometric / torch_scatter
are not needed: the three PyG helpers sym_norm2 uses are restated in torch.
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch
from scipy.sparse import coo_matrix
from torch.nn import LeakyReLU, init
from torch.nn.modules.module import Module
from torch.nn.parameter import Parameter

from . import config
from .pynq_compat import Overlay, allocate
from .quant import (fake_quantization, fake_quantization_b, fake_quantization_b2, float_bits,  # noqa: F401
                    generate_quantization_constants, generate_quantization_qbits_constants,
                    generate_quantization_uqbits_constants, layer_constants, quantization, quantization_b,
                    quantization_fbits, quantization_qbits, quantization_ufbits, quantization_uqbits)

my_ip = None
ol = None
layern = 1
cur_max_fea = 0.0
cur_max_fea2 = 0.0
frac_bits_o = 16


# ------------------------------------------------------------------------------------------
# sgrace.py:18-51
# ------------------------------------------------------------------------------------------
def sym_norm2(edge_index, num_nodes, edge_weight=None, fill=0, dtype=None):
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
    # add_remaining_self_loops: keep the weight of existing self-loops, add `fill` for the others
    row, col = edge_index
    mask = row != col
    loop_weight = torch.full((num_nodes,), float(fill), dtype=edge_weight.dtype, device=edge_index.device)
    inv = ~mask
    if inv.any():
        loop_weight[row[inv]] = edge_weight[inv]
    loop_index = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    edge_index = torch.cat([edge_index[:, mask], torch.stack([loop_index, loop_index])], dim=1)
    edge_weight = torch.cat([edge_weight[mask], loop_weight])
    # sort_edge_index: by row, then column
    order = torch.argsort(edge_index[0] * num_nodes + edge_index[1], stable=True)
    edge_index, edge_weight = edge_index[:, order], edge_weight[order]
    row, col = edge_index
    deg = torch.zeros(num_nodes, dtype=edge_weight.dtype, device=edge_index.device).index_add_(0, row, edge_weight)
    deg_inv_sqrt = deg.pow(-0.5)
    deg_inv_sqrt[deg_inv_sqrt == float('inf')] = 0
    return edge_index, deg_inv_sqrt[row] * edge_weight * deg_inv_sqrt[col]


def _prep_handle(device):
    """One library handle per device for the preparation calls (created on first use)."""
    from . import _lib
    idx = device.index if device.index is not None else torch.cuda.current_device()
    h = _PREP_HANDLES.get(idx)
    if h is None:
        h = _PREP_HANDLES[idx] = _lib.Handle(idx)
    h.set_stream(torch.cuda.current_stream(device).cuda_stream or 1)
    return h


_PREP_HANDLES = {}


def sym_norm2_device(edge_index, num_nodes, edge_weight=None, fill=0, dtype=None):
    """sym_norm2 on the GPU (`sgrace_sym_norm`): same signature, same values bit for bit, for CUDA tensors.
    The reference runs this on the host before every forward."""
    if not edge_index.is_cuda:
        raise ValueError("sym_norm2_device needs CUDA tensors; sym_norm2 is the host version")
    dev = edge_index.device
    nnz = int(edge_index.size(1))
    row = edge_index[0].to(torch.int32).contiguous()
    col = edge_index[1].to(torch.int32).contiguous()
    w = None if edge_weight is None else edge_weight.to(torch.float32).contiguous()
    cap = nnz + int(num_nodes)
    out_row = torch.empty(cap, dtype=torch.int32, device=dev)
    out_col = torch.empty(cap, dtype=torch.int32, device=dev)
    out_val = torch.empty(cap, dtype=torch.float32, device=dev)
    h = _prep_handle(dev)
    n = h.sym_norm(row.data_ptr(), col.data_ptr(), w.data_ptr() if w is not None else 0, nnz, int(num_nodes), float(fill), cap,
                   out_row.data_ptr(), out_col.data_ptr(), out_val.data_ptr())
    ei = torch.stack([out_row[:n], out_col[:n]]).to(edge_index.dtype)
    return ei, out_val[:n]


def to_sparse_device(x):
    """CSR of a dense feature matrix on the GPU (`sgrace_dense_to_csr`): (rowptr, col, val) int32/int32/float32
    CUDA tensors -- what `input.to_sparse()` feeds the accelerator buffers with (sgrace.py:1218-1227)."""
    if not x.is_cuda:
        raise ValueError("to_sparse_device needs a CUDA tensor")
    x = x.to(torch.float32).contiguous()
    n, m = x.shape
    dev = x.device
    h = _prep_handle(dev)
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    cap = max(1, min(n * m, 1 << 20))
    while True:
        col = torch.empty(cap, dtype=torch.int32, device=dev)
        val = torch.empty(cap, dtype=torch.float32, device=dev)
        rc, nnz = h.dense_to_csr(x.data_ptr(), n, m, cap, rowptr.data_ptr(), col.data_ptr(), val.data_ptr())
        if rc == 0:
            return rowptr, col[:nnz], val[:nnz]
        cap = nnz                      # SGRACE_EBOUNDS: the call reported the size it needs


# ------------------------------------------------------------------------------------------
# sgrace.py:267-294
# ------------------------------------------------------------------------------------------
class RPYNQ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return input.clone()

    @staticmethod
    def backward(ctx, grad_output):
        input, = ctx.saved_tensors
        grad_input = grad_output.clone()
        grad_input[input == 0] = 0       # the ReLU ran inside the accelerator
        return grad_input


# ------------------------------------------------------------------------------------------
# sgrace.py:298-1126
# ------------------------------------------------------------------------------------------
class FPYNQ_GAT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, my_ip, self, adj, nnz_adj, input, weights, attention, out_features, dropout, relu):
        global layern, cur_max_fea, cur_max_fea2
        if config.acc != 1:
            raise RuntimeError("config.acc == 0 selects the reference's software emulation, which this "
                               "package does not ship (there is no CPU compute path); set config.acc = 1")
        rm = my_ip.register_map
        c = _consts
        # per-layer constants: the driver alternates between the layer-1 and layer-2 tables (:327-359)
        rm.scale_fea = c["scale_fea"]
        rm.deq_factor = float_bits(c["deq_o"])
        rm.quantization_scale_fea = float_bits(1 / c["f_s"])
        rm.quantization_scale_w = float_bits(1 / c["w_s"])
        cur_layer = layern
        layern = 2 if layern == 1 else 1
        rm.quantization_scale_adj = float_bits(1 / c["a_s"])
        for i in "1234":
            setattr(rm, f"rowPtr_adj{i}_offset_1", rowPtr_adj_buffer.physical_address)
            setattr(rm, f"columnIndex_adj{i}_offset_1", columnIndex_adj_buffer.physical_address)
            setattr(rm, f"values_adj{i}_offset_1", values_adj_buffer.physical_address)
            setattr(rm, f"rowPtr_fea{i}_offset_1", rowPtr_fea_buffer.physical_address)
            setattr(rm, f"columnIndex_fea{i}_offset_1", columnIndex_fea_buffer.physical_address)
            setattr(rm, f"values_fea{i}_offset_1", values_fea_buffer.physical_address)
            setattr(rm, f"D{i}_offset_1", D_buffer.physical_address)
        rm.N_adj = input.shape[0]
        rm.M_adj = input.shape[0]
        rm.M_fea = input.shape[1]
        rm.P_w = weights.shape[1]
        rm.E1_offset_1 = E_buffer.physical_address
        rm.S1_offset_1 = S_buffer.physical_address
        rm.B_offset_1 = B_buffer.physical_address
        rm.ate_m_offset_1 = attention_buffer.physical_address
        support = torch.transpose(weights, 0, 1)         # the B buffer holds W transposed
        support_pynq_q = support.data.numpy().reshape(1, weights.shape[0] * weights.shape[1])
        B_buffer[0:(weights.shape[0] * weights.shape[1])] = support_pynq_q.astype(config.float_type)
        if config.compute_attention == 1:
            attention_q = attention.reshape(1, attention.shape[0] * attention.shape[1]).detach().numpy()
            attention_buffer[0:(attention.shape[0] * attention.shape[1])] = attention_q.astype(config.float_type)
        rm.quantized_multiplier = c["internal_quantization"]
        amult = time.time()
        rm.CTRL.AP_START = 1
        kernel_done = rm.CTRL.AP_DONE
        while kernel_done == 0:
            kernel_done = rm.CTRL.AP_DONE
        my_ip.handle.wait()
        if config.profiling == 1:
            print('Accelerator forward kernel mult time: {:.5f}ms'.format(1000 * (time.time() - amult)))
        output_acc = np.array(D_buffer[0:input.shape[0] * weights.shape[1]])
        max_fea_float = float(rm.max_fea) / (2 ** frac_bits_o)
        if cur_layer == 1:
            cur_max_fea = max(cur_max_fea, max_fea_float)
        else:
            cur_max_fea2 = max(cur_max_fea2, max_fea_float)
        output_acc = torch.from_numpy(output_acc.reshape(input.shape[0], weights.shape[1])).float()
        ctx.nheads, ctx.alpha = self.nheads, self.alpha
        if config.compute_attention == 1:
            n = input.shape[0]
            rindex = np.array(rowPtr_adj_buffer[0:nnz_adj])
            cindex = np.array(columnIndex_adj_buffer[0:nnz_adj])
            output_e_val = np.array(E_buffer[0:nnz_adj]).astype(config.float_type)
            output_s_val = np.array(S_buffer[0:nnz_adj]).astype(config.float_type)
            output_s = torch.from_numpy(np.asarray(coo_matrix((output_s_val, (rindex, cindex)), shape=(n, n)).todense())).float()
            output_e = torch.from_numpy(np.asarray(coo_matrix((output_e_val, (rindex, cindex)), shape=(n, n)).todense())).float()
            ctx.save_for_backward(adj, input, weights, output_e, output_s, output_acc)
        else:
            ctx.save_for_backward(adj, input, weights, adj, adj, output_acc)
        return output_acc

    @staticmethod
    def backward(ctx, grad_output):
        if config.accb == 1:
            raise RuntimeError("config.accb == 1 (hardware backward, gemm_mode 2) is not available; use accb = 0")
        adj, input, weights, e, attentions, output = ctx.saved_tensors
        alpha = ctx.alpha
        input_t, weights_t = input.t(), weights.t()
        if config.compute_attention == 1:
            support = torch.mm(weights_t, input_t)
            softmax_out = torch.mm(grad_output, support)
            dx = attentions * softmax_out
            s = dx.sum(axis=dx.ndim - 1, keepdims=True)
            soft_gradient = dx - attentions * s
            adj_d = adj.to_dense()
            soft_gradient = torch.where(adj_d > 0, soft_gradient, torch.zeros_like(soft_gradient))
            soft_gradient = ((e > 0) + alpha * (e <= 0)) * soft_gradient
            torch_ones = torch.ones(input.shape[0])
            grad_attention1 = torch.matmul(torch.mm(support, soft_gradient), torch_ones)
            grad_attention2 = torch.matmul(torch_ones, torch.mm(soft_gradient, torch.mm(input, weights)))
            output_attention = torch.cat((grad_attention1, grad_attention2)).unsqueeze(1)
            support = torch.mm(grad_output, weights_t)
            output_input = torch.mm(attentions, support)
            support = torch.mm(attentions, grad_output)
        else:
            output_attention = torch.zeros(size=(weights.shape[1] * 2, 1))
            support = torch.mm(grad_output, weights_t)
            output_input = torch.mm(adj, support)
            support = torch.mm(adj, grad_output)
        output_weights = torch.mm(input_t, support)
        return None, None, None, None, output_input, output_weights, output_attention, None, None, None


# ------------------------------------------------------------------------------------------
# sgrace.py:1146-1265
# ------------------------------------------------------------------------------------------
class Relu_SGRACE(Module):
    def __init__(self):
        super().__init__()
        self.fn = RPYNQ.apply

    def forward(self, x):
        return self.fn(x)


class GATConv_SGRACE(Module):
    def __init__(self, in_features, out_features, nheads=1, bias=True, dropout=0.2, alpha=0.2, concat=False):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.alpha, self.dropout = alpha, dropout
        self.weight = Parameter(torch.FloatTensor(in_features, out_features * nheads))
        init.xavier_uniform_(self.weight.data, gain=1.414)
        self.attention = Parameter(torch.empty(size=(2 * out_features * nheads, 1)))
        init.xavier_uniform_(self.attention.data, gain=1.414)
        self.leakyrelu = LeakyReLU(self.alpha)
        self.nheads, self.concat = nheads, concat
        self.fn = FPYNQ_GAT.apply
        self.my_ip = my_ip if config.acc == 1 else None
        if bias:
            self.bias = Parameter(torch.FloatTensor(out_features))
        else:
            self.register_parameter('bias', None)

    def run_kernel(self):
        self.my_ip.register_map.CTRL.AP_START = 1
        kernel_done = self.my_ip.register_map.CTRL.AP_DONE
        while kernel_done == 0:
            kernel_done = self.my_ip.register_map.CTRL.AP_DONE

    def forward(self, compute_attention, dense, relu, input, edge_index, norm, adj):
        if self.my_ip is None:
            self.my_ip = my_ip
        rm = self.my_ip.register_map
        rm.relu = relu
        rm.gemm_mode = dense
        rm.gat_mode = compute_attention
        self.my_ip.configure(leaky_alpha=self.alpha)
        if dense == 0:
            pynq_features = input.detach().to_sparse()           # COO: row indices go into the rowPtr buffer
            nnz_fea = len(pynq_features.values())
            rm.nnz_fea1 = nnz_fea
            rowPtr_fea_buffer[0:nnz_fea] = pynq_features.indices()[0].numpy()
            columnIndex_fea_buffer[0:nnz_fea] = pynq_features.indices()[1].numpy()
            values_fea_buffer[0:nnz_fea] = pynq_features.values().numpy()
        else:
            xaux = input.detach().numpy()
            values_fea_buffer[0:xaux.shape[0] * xaux.shape[1]] = xaux.reshape(1, xaux.shape[0] * xaux.shape[1])
        nnz_adj = len(norm)
        rowPtr_adj_buffer[0:nnz_adj] = edge_index[0].numpy()
        values_adj_buffer[0:nnz_adj] = norm.detach().numpy()
        columnIndex_adj_buffer[0:nnz_adj] = edge_index[1].numpy()
        rm.nnz_adj1 = nnz_adj
        return self.fn(self.my_ip, self, adj, nnz_adj, input, self.weight, self.attention, self.out_features,
                       self.dropout, relu)

    def __repr__(self):
        return f"{self.__class__.__name__} ({self.in_features} -> {self.out_features})"


# ------------------------------------------------------------------------------------------
# sgrace.py:1271-1896
# ------------------------------------------------------------------------------------------
def init_SGRACE(device: int = 0):
    """Overlay, buffers sized by the config maxima, quantisation constants for config.w_qbits,
    constant registers.  Returns the IP handle (also left in the module global `my_ip`)."""
    global ol, my_ip, _consts, layern
    global attention_buffer, bias_buffer, profiling_buffer, rowPtr_fea_buffer, columnIndex_fea_buffer
    global values_fea_buffer, rowPtr_adj_buffer, columnIndex_adj_buffer, values_adj_buffer, B_buffer, D_buffer
    global E_buffer, S_buffer
    ol = Overlay("gat_all_unsigned.bit", device=device)
    my_ip = ol.mmult_top_0
    my_ip.configure(qbits=config.w_qbits if config.hardware_quantize and config.fake_quantization else 0)
    _consts = layer_constants(config.w_qbits)
    layern = 1
    ft = config.float_type
    attention_buffer = allocate(config.P_w * 2, dtype=ft)
    bias_buffer = allocate(1024, dtype=np.int32)
    profiling_buffer = allocate(16, dtype=np.int64)
    rowPtr_fea_buffer = allocate(config.NNZ_fea, dtype=np.int32)
    columnIndex_fea_buffer = allocate(config.NNZ_fea, dtype=np.int32)
    values_fea_buffer = allocate(config.NNZ_fea, dtype=ft)
    rowPtr_adj_buffer = allocate(config.NNZ_adj, dtype=np.int32)
    columnIndex_adj_buffer = allocate(config.NNZ_adj, dtype=np.int32)
    values_adj_buffer = allocate(config.NNZ_adj, dtype=ft)
    B_buffer = allocate(config.N_adj * config.P_w * config.head_count, dtype=ft)
    D_buffer = allocate(config.N_adj * config.P_w * config.head_count, dtype=ft)
    E_buffer = allocate(config.NNZ_adj, dtype=ft)
    S_buffer = allocate(config.NNZ_adj, dtype=ft)
    rm = my_ip.register_map
    rm.f_align = _consts["f_align"]
    rm.beta_qu = _consts["beta_qu"]
    rm.load_weights = config.load_weights
    rm.gat_mode = config.compute_attention
    rm.layer_count = config.layer_count
    rm.stream_mode = config.stream_mode
    rm.profiling_offset_1 = profiling_buffer.physical_address
    rm.bias_offset_1 = bias_buffer.physical_address
    for name in ("E2", "E3", "E4", "S2", "S3", "S4"):        # written by the reference, absent from the register map
        setattr(rm, f"{name}_offset_1", 0)
    return my_ip


def free_SGRACE():
    """Release the buffers init_SGRACE allocated (the reference never does; tests need to)."""
    for name in ("attention_buffer", "bias_buffer", "profiling_buffer", "rowPtr_fea_buffer", "columnIndex_fea_buffer",
                 "values_fea_buffer", "rowPtr_adj_buffer", "columnIndex_adj_buffer", "values_adj_buffer", "B_buffer",
                 "D_buffer", "E_buffer", "S_buffer"):
        b = globals().get(name)
        if b is not None:
            b.freebuffer()
            globals()[name] = None
