"""Host mirror of the reference's molecule-GCN notebook layer code, running on libsgrace_b200.

Reference: jupyter/molecule_gcn/Graph_Classification.ipynb (code cells, 0-based):
    cell 4   buffer allocation + pointer registers          -> `NotebookBuffers`
    cell 7   RPYNQ, FPYNQ (autograd Functions)              -> `RPYNQ`, `FPYNQ`
    cell 8   Relu_pynq, GraphConvolution_pynq               -> same names
    cell 9   GCN_PYNQ                                       -> same name
    cell 10  train() / test()                               -> `train_epoch`, `evaluate`
Same names, argument order and meaning.  What differs from the notebook:
  * `from pynq import allocate, Overlay` is `sgracex1_b200.pynq_compat`;
  * the notebook's module-level buffer globals live in a `NotebookBuffers` object hung on `my_ip`;
  * torch_geometric is not needed: `to_dense_adj` / `global_mean_pool` are restated in torch;
  * `GCN_B200` is the same model with every buffer resident in HBM (no PCIe per layer), used by the
    data-parallel trainer; its parameters have the notebook's names (state_dict compatible).
The accelerator does the forward of both GCN layers; the backward is the notebook's saved-tensor
backward (`accb = 0` branch): grad_W = X^T (A g), grad_X = (A g) W^T  -- A is NOT transposed.
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch
import torch.nn.functional as F
from torch.nn import Linear
from torch.nn.modules.module import Module
from torch.nn.parameter import Parameter

from . import _lib
from .pynq_compat import Overlay, allocate

VERBOSE = False      # the notebook prints a timing line per kernel call


# ------------------------------------------------------------------------------------------
# torch_geometric stand-ins (the notebook imports these from PyG)
# ------------------------------------------------------------------------------------------
def to_dense_adj(edge_index, num_nodes=None):
    """torch_geometric.utils.to_dense_adj for one graph batch: duplicate edges add up."""
    n = int(num_nodes) if num_nodes is not None else (int(edge_index.max()) + 1 if edge_index.numel() else 0)
    adj = torch.zeros(n, n, dtype=torch.float32)
    if edge_index.numel():
        adj.index_put_((edge_index[0], edge_index[1]), torch.ones(edge_index.shape[1]), accumulate=True)
    return adj


def edge_index_to_csr(edge_index, num_nodes):
    """What `to_dense_adj(edge_index)._to_sparse_csr()` yields, without the N x N detour."""
    ei = edge_index.cpu().numpy().astype(np.int64)
    key = ei[0] * num_nodes + ei[1]
    uniq, counts = np.unique(key, return_counts=True)
    rows, cols = uniq // num_nodes, uniq % num_nodes
    rp = np.zeros(num_nodes + 1, np.int32)
    np.cumsum(np.bincount(rows, minlength=num_nodes), out=rp[1:])
    return rp, cols.astype(np.int32), counts.astype(np.float32)


def global_mean_pool(x, batch, num_graphs=None):
    g = int(num_graphs) if num_graphs is not None else int(batch.max()) + 1
    out = torch.zeros(g, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, batch, x)
    cnt = torch.zeros(g, dtype=x.dtype, device=x.device).index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return out / cnt.clamp(min=1).unsqueeze(1)


# ------------------------------------------------------------------------------------------
# notebook cell 4
# ------------------------------------------------------------------------------------------
class NotebookBuffers:
    """The notebook's global buffers; `dtype` np.float16 selects the HALF C-simulation mode (what
    the notebook's bitstream computes), np.float32 the float32 fast path."""

    def __init__(self, my_ip, N_adj, NNZ_adj, NNZ_fea, P_w, dtype=np.float16):
        self.my_ip, self.dtype = my_ip, np.dtype(dtype)
        half = self.dtype == np.float16
        my_ip.configure(mode=_lib.MODE_F16_CSIM if half else _lib.MODE_F32_FAST, index_format=0, staging=1)
        al = lambda n, dt: allocate(int(n), dtype=dt, target=my_ip)
        self.quantized_multiplier_buffer = al(1024, np.int32)
        self.bias_buffer = al(1024, np.int32)
        self.shift_buffer = al(1024, np.int32)
        self.profiling_buffer = al(16, np.int64)
        self.rowPtr_fea_buffer = al(N_adj + 1, np.int32)
        self.columnIndex_fea_buffer = al(NNZ_fea, np.int32)
        self.values_fea_buffer = al(NNZ_fea, dtype)
        self.rowPtr_adj_buffer = al(N_adj + 1, np.int32)
        self.columnIndex_adj_buffer = al(NNZ_adj, np.int32)
        self.values_adj_buffer = al(NNZ_adj, dtype)
        self.B_buffer = al(N_adj * P_w, dtype)
        self.D_buffer = al(N_adj * P_w, dtype)
        rm = my_ip.register_map
        rm.B_offset_1 = self.B_buffer.physical_address
        for i in "1234":
            setattr(rm, f"rowPtr_fea{i}_offset_1", self.rowPtr_fea_buffer.physical_address)
            setattr(rm, f"columnIndex_fea{i}_offset_1", self.columnIndex_fea_buffer.physical_address)
            setattr(rm, f"values_fea{i}_offset_1", self.values_fea_buffer.physical_address)
            setattr(rm, f"rowPtr_adj{i}_offset_1", self.rowPtr_adj_buffer.physical_address)
            setattr(rm, f"columnIndex_adj{i}_offset_1", self.columnIndex_adj_buffer.physical_address)
            setattr(rm, f"values_adj{i}_offset_1", self.values_adj_buffer.physical_address)
        rm.quantized_multiplier_offset_1 = self.quantized_multiplier_buffer.physical_address
        rm.bias_offset_1 = self.bias_buffer.physical_address
        rm.shift_offset_1 = self.shift_buffer.physical_address
        rm.profiling_offset_1 = self.profiling_buffer.physical_address
        my_ip.buffers = self

    def as_args(self):
        """The eight buffers GCN_PYNQ.forward takes, in the notebook's order."""
        return (self.rowPtr_fea_buffer, self.columnIndex_fea_buffer, self.values_fea_buffer, self.rowPtr_adj_buffer,
                self.columnIndex_adj_buffer, self.values_adj_buffer, self.B_buffer, self.D_buffer)

    def free(self):
        for v in vars(self).values():
            if hasattr(v, "freebuffer"):
                v.freebuffer()


def _to_buffer_dtype(t, dtype):
    a = t.detach().cpu().numpy()
    return a.astype(dtype)


# ------------------------------------------------------------------------------------------
# notebook cell 7
# ------------------------------------------------------------------------------------------
class RPYNQ(torch.autograd.Function):
    """Identity forward (the ReLU ran inside the accelerator); backward zeroes the gradient where
    the accelerator's output is exactly 0."""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return input.clone()

    @staticmethod
    def backward(ctx, grad_output):
        input, = ctx.saved_tensors
        grad_input = grad_output.clone()
        grad_input[input == 0] = 0
        return grad_input


class FPYNQ(torch.autograd.Function):
    """One layer on the accelerator through the register map (host buffers, AP_START/AP_DONE)."""

    @staticmethod
    def forward(ctx, my_ip, adj, input, weights):
        b = my_ip.buffers
        rm = my_ip.register_map
        rm.N_adj = adj.shape[0]
        rm.M_adj = adj.shape[0]
        rm.M_fea = input.shape[1]
        rm.P_w = weights.shape[1]
        for i in "1234":
            setattr(rm, f"D{i}_offset_1", b.D_buffer.physical_address)
        rm.values_fea1_offset_1 = b.values_fea_buffer.physical_address
        rm.B_offset_1 = b.B_buffer.physical_address
        support = torch.transpose(weights, 0, 1)          # the B buffer holds W transposed
        support_pynq = support.data.numpy().reshape(1, weights.shape[0] * weights.shape[1])
        b.B_buffer[0:(weights.shape[0] * weights.shape[1])] = support_pynq
        amult = time.time()
        rm.CTRL.AP_START = 1
        kernel_done = rm.CTRL.AP_DONE
        while kernel_done == 0:
            kernel_done = rm.CTRL.AP_DONE
        my_ip.handle.wait()
        if VERBOSE:
            print('acc forward kernel mult: {:.5f}s'.format(time.time() - amult))
        output_acc = np.array(b.D_buffer[0:adj.shape[0] * weights.shape[1]])
        output_acc = torch.from_numpy(output_acc.reshape(adj.shape[0], weights.shape[1]))
        ctx.save_for_backward(adj, input, weights, output_acc)
        return output_acc

    @staticmethod
    def backward(ctx, grad_output):
        adj, input, weights, output = ctx.saved_tensors
        input = input.float()
        grad_output = grad_output.float()
        tmult = time.time()
        ag = adj @ grad_output
        grad_weights = input.t() @ ag
        grad_input = ag @ weights.t()
        if VERBOSE:
            print('CPU backward kernel 2xmult: {:.5f}s'.format(time.time() - tmult))
        return None, None, grad_input, grad_weights


# ------------------------------------------------------------------------------------------
# notebook cell 8
# ------------------------------------------------------------------------------------------
class Relu_pynq(Module):
    def __init__(self):
        super().__init__()
        self.fn = RPYNQ.apply

    def forward(self, x):
        return self.fn(x)


class GraphConvolution_pynq(Module):
    def __init__(self, in_features, out_features, my_ip, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = Parameter(torch.FloatTensor(in_features, out_features))
        self.fn = FPYNQ.apply
        self.my_ip = my_ip
        if bias:
            self.bias = Parameter(torch.FloatTensor(out_features))      # created, never added (as in the notebook)
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def run_kernel(self):
        self.my_ip.register_map.CTRL.AP_START = 1
        kernel_done = self.my_ip.register_map.CTRL.AP_DONE
        while kernel_done == 0:
            kernel_done = self.my_ip.register_map.CTRL.AP_DONE

    def forward(self, acc, dense, relu, input, adj, rowPtr_fea_buffer, columnIndex_fea_buffer, values_fea_buffer,
                rowPtr_adj_buffer, columnIndex_adj_buffer, values_adj_buffer, B_buffer, D_buffer):
        if acc == 1:
            self.my_ip.register_map.relu = relu
            self.my_ip.register_map.gemm_mode = dense
            return self.fn(self.my_ip, adj, input, self.weight)
        input = input.float()
        return adj @ input @ self.weight

    def __repr__(self):
        return f"{self.__class__.__name__} ({self.in_features} -> {self.out_features})"


# ------------------------------------------------------------------------------------------
# notebook cell 9
# ------------------------------------------------------------------------------------------
class GCN_PYNQ(torch.nn.Module):
    def __init__(self, hidden_channels, my_ip, num_node_features=7, num_classes=2):
        super().__init__()
        torch.manual_seed(12345)
        self.conv1 = GraphConvolution_pynq(num_node_features, hidden_channels, my_ip)
        self.conv2 = GraphConvolution_pynq(hidden_channels, hidden_channels, my_ip)
        self.reluh = Relu_pynq()
        self.lin = Linear(hidden_channels, num_classes)

    def forward(self, acc, x, edge_index, batch, rowPtr_fea_buffer, columnIndex_fea_buffer, values_fea_buffer,
                rowPtr_adj_buffer, columnIndex_adj_buffer, values_adj_buffer, B_buffer, D_buffer):
        n = x.shape[0]
        adj = torch.squeeze(to_dense_adj(edge_index, n))
        pynq_adj = adj.to_sparse_csr()
        pynq_features = x.to_sparse_csr()
        vt = values_adj_buffer.dtype
        rowPtr_adj_buffer[0:len(pynq_adj.crow_indices())] = pynq_adj.crow_indices().numpy()
        columnIndex_adj_buffer[0:len(pynq_adj.col_indices())] = pynq_adj.col_indices().numpy()
        values_adj_buffer[0:len(pynq_adj.values())] = pynq_adj.values().numpy().astype(vt)
        rowPtr_fea_buffer[0:len(pynq_features.crow_indices())] = pynq_features.crow_indices().numpy()
        columnIndex_fea_buffer[0:len(pynq_features.col_indices())] = pynq_features.col_indices().numpy()
        values_fea_buffer[0:len(pynq_features.values())] = pynq_features.values().numpy().astype(vt)

        dense, relu, acc = 0, 1, 1                      # the notebook overrides `acc` here (cell 9)
        x = self.conv1(acc, dense, relu, x, adj, rowPtr_fea_buffer, columnIndex_fea_buffer, values_fea_buffer,
                       rowPtr_adj_buffer, columnIndex_adj_buffer, values_adj_buffer, B_buffer, D_buffer)
        x = self.reluh(x)
        dense = 1                                        # layer 2 streams the dense activations
        xaux = x.detach().numpy()
        values_fea_buffer[0:(x.shape[0] * x.shape[1])] = xaux.reshape(1, x.shape[0] * x.shape[1])
        relu = 0
        x = self.conv2(acc, dense, relu, x, adj, rowPtr_fea_buffer, columnIndex_fea_buffer, values_fea_buffer,
                       rowPtr_adj_buffer, columnIndex_adj_buffer, values_adj_buffer, B_buffer, D_buffer)
        x = x.float()
        x = global_mean_pool(x, batch)
        x = F.dropout(x, p=0.5, training=self.training)
        return self.lin(x)


# ------------------------------------------------------------------------------------------
# device-resident variant: same model, buffers in HBM, sgrace_layer_run per layer
# ------------------------------------------------------------------------------------------
class _DeviceGraphLayer(torch.autograd.Function):
    """D = act(A (X W)) with every operand a CUDA tensor; forward = one sgrace_layer_run, backward =
    the notebook's saved-tensor formulas with A g computed by the ADJ kernel."""

    @staticmethod
    def forward(ctx, handle, adj_csr, x, weight, relu, layer_fn):
        N, M = x.shape
        P = weight.shape[1]
        out = layer_fn(handle, adj_csr, x, weight, bool(relu))
        ctx.handle, ctx.adj_csr, ctx.layer_fn = handle, adj_csr, layer_fn
        ctx.save_for_backward(x, weight)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        x, weight = ctx.saved_tensors
        ag = ctx.layer_fn(ctx.handle, ctx.adj_csr, grad_output.contiguous(), None, False)   # A g (no FEA stage)
        h = ctx.handle
        if h is not None and ag.is_cuda:
            # grad_W = X^T (A g): tall-skinny reduction kernel; grad_X = (A g) W^T: the dense FEA stage with B = W
            N, M = x.shape
            P = weight.shape[1]
            xc, wc = x.contiguous(), weight.detach().contiguous()
            grad_w = torch.empty(M, P, dtype=torch.float32, device=ag.device)
            h.xty_run(xc.data_ptr(), ag.data_ptr(), grad_w.data_ptr(), N, M, P)
            grad_x = torch.empty(N, M, dtype=torch.float32, device=ag.device)
            h.dense_run(ag.data_ptr(), wc.data_ptr(), grad_x.data_ptr(), N, P, M, 0)
            ctx.keep = (xc, wc)
        else:
            grad_w = x.t() @ ag
            grad_x = ag @ weight.t()
        return None, None, grad_x, grad_w, None, None


def sgrace_layer_device(handle, adj_csr, x, weight, relu):
    """Run one layer (or, with weight=None, only the ADJ stage on `x`) on device tensors through the
    C ABI.  adj_csr = (rowptr int32, col int32, val float32) CUDA tensors."""
    rp, ci, va = adj_csr
    N = rp.numel() - 1
    d = _lib.LayerDesc()
    d.N_adj = d.M_adj = N
    d.rowPtr_adj, d.columnIndex_adj, d.values_adj = rp.data_ptr(), ci.data_ptr(), va.data_ptr()
    d.nnz_adj = int(ci.numel())
    handle.set_option(_lib.OPT_MODE, _lib.MODE_F32_FAST)
    handle.set_option(_lib.OPT_INDEX_FORMAT, 0)
    # the legacy default stream has handle 0, which the ABI reads as "your own stream": name it (cudaStreamLegacy = 1)
    handle.set_stream(torch.cuda.current_stream(x.device).cuda_stream or 1)
    x = x.contiguous()
    if weight is None:
        P = x.shape[1]
        out = torch.empty(N, P, dtype=torch.float32, device=x.device)
        d.P_w, d.relu, d.D = P, int(relu), out.data_ptr()
        handle.adj_run(d, x.data_ptr(), x.shape[0])
        return out
    M, P = weight.shape
    B = weight.detach().t().contiguous()                 # W transposed, P x M
    out = torch.empty(N, P, dtype=torch.float32, device=x.device)
    xw = torch.empty(N, P, dtype=torch.float32, device=x.device)
    d.M_fea, d.P_w, d.relu, d.gemm_mode = M, P, int(relu), 1
    d.values_fea, d.B, d.D, d.XW = x.data_ptr(), B.data_ptr(), out.data_ptr(), xw.data_ptr()
    handle.layer_run(d)
    return out


class _DevicePool(torch.autograd.Function):
    """global_mean_pool as two CSR products on the accelerator's ADJ kernel: pooled = P h with
    P[g, n] = 1/|graph g| (graphs are contiguous node ranges), grad_h = P^T grad_pooled."""

    @staticmethod
    def forward(ctx, handle, pool_csr, pool_t_csr, h, layer_fn):
        ctx.handle, ctx.pool_t_csr, ctx.layer_fn = handle, pool_t_csr, layer_fn
        return layer_fn(handle, pool_csr, h.contiguous(), None, False)

    @staticmethod
    def backward(ctx, grad_out):
        return None, None, None, ctx.layer_fn(ctx.handle, ctx.pool_t_csr, grad_out.contiguous(), None, False), None


def pooling_csr(batch, num_graphs):
    """(P, P^T) as CSR triples for a sorted `batch` vector (node -> graph)."""
    n = batch.numel()
    dev = batch.device
    counts = torch.bincount(batch, minlength=num_graphs)
    rp = torch.zeros(num_graphs + 1, dtype=torch.int32, device=dev)
    rp[1:] = torch.cumsum(counts, 0).to(torch.int32)
    inv = (1.0 / counts.clamp(min=1).to(torch.float32))
    val = inv[batch].contiguous()
    p = (rp, torch.arange(n, dtype=torch.int32, device=dev), val)
    pt = (torch.arange(n + 1, dtype=torch.int32, device=dev), batch.to(torch.int32).contiguous(), val)
    return p, pt


class GCN_B200(torch.nn.Module):
    """GCN_PYNQ with device-resident buffers.  `layer_fn` is the accelerator call (default: the C
    ABI); the CPU tests of the distributed logic inject a torch stand-in."""

    def __init__(self, hidden_channels, handle=None, num_node_features=7, num_classes=2, layer_fn=sgrace_layer_device):
        super().__init__()
        torch.manual_seed(12345)
        self.handle, self.layer_fn = handle, layer_fn
        self.conv1 = GraphConvolution_pynq(num_node_features, hidden_channels, None)
        self.conv2 = GraphConvolution_pynq(hidden_channels, hidden_channels, None)
        self.lin = Linear(hidden_channels, num_classes)

    def forward(self, x, adj_csr, batch, num_graphs, pool=None):
        """`pool` = pooling_csr(batch, num_graphs) computed once per batch (None: torch index_add pooling)."""
        h = _DeviceGraphLayer.apply(self.handle, adj_csr, x, self.conv1.weight, 1, self.layer_fn)
        h = RPYNQ.apply(h)
        h = _DeviceGraphLayer.apply(self.handle, adj_csr, h, self.conv2.weight, 0, self.layer_fn)
        if pool is not None:
            h = _DevicePool.apply(self.handle, pool[0], pool[1], h, self.layer_fn)
        else:
            h = global_mean_pool(h, batch, num_graphs)
        h = F.dropout(h, p=0.5, training=self.training)
        return self.lin(h)


# ------------------------------------------------------------------------------------------
# notebook cell 10
# ------------------------------------------------------------------------------------------
def train_epoch(model, loader, optimizer, criterion, buffers):
    model.train()
    for data in loader:
        out = model(0, data.x, data.edge_index, data.batch, *buffers.as_args())
        loss = criterion(out, data.y)
        loss.backward()
        optimizer.step()
        optimizer.zero_grad()


def evaluate(model, loader, buffers):
    model.eval()
    correct = total = 0
    for data in loader:
        out = model(0, data.x, data.edge_index, data.batch, *buffers.as_args())
        correct += int((out.argmax(dim=1) == data.y).sum())
        total += int(data.y.numel())
    return correct / max(total, 1)


# ------------------------------------------------------------------------------------------
# bench.py --workload molecule : data-parallel training, graphs/s
# ------------------------------------------------------------------------------------------
def molecule_record(steps, warmup, rank, world, local, graphs_per_rank=188 * 64, hidden=64):
    """Molecule-GCN training (two accelerator layers forward, saved-tensor backward, DP gradient all-reduce, Adam),
    graphs sharded over the ranks.  The process group must exist when world > 1.  Returns the record on rank 0."""
    import os

    import torch.distributed as dist

    from . import dist as sdist
    from . import graphs as G

    dev = torch.device(f"cuda:{local}")
    prob, batch_np, y_np = G.molecule_batch(n_graphs=graphs_per_rank, seed=12345 + rank, P=hidden)
    handle = _lib.Handle(local)
    handle.set_option(_lib.OPT_STAGING, 0)
    model = GCN_B200(hidden, handle).to(dev)
    use_graph = not os.environ.get("SGRACE_NO_GRAPH")
    opt = torch.optim.Adam(model.parameters(), lr=0.01, capturable=use_graph)
    crit = torch.nn.CrossEntropyLoss(reduction="sum")
    x = torch.zeros(prob.N, prob.M, device=dev)
    x[torch.arange(prob.N, device=dev), torch.from_numpy(prob.fea_col.astype(np.int64)).to(dev)] = 1.0
    adj = tuple(torch.from_numpy(a).to(dev) for a in (prob.adj_rowptr, prob.adj_col, prob.adj_val))
    batch = torch.from_numpy(batch_np).to(dev)
    y = torch.from_numpy(y_np.astype(np.int64)).to(dev)
    total_graphs = graphs_per_rank * world
    loss_buf = torch.zeros((), device=dev)
    pool = pooling_csr(batch, graphs_per_rank)

    def step():
        opt.zero_grad(set_to_none=False)
        out = model(x, adj, batch, graphs_per_rank, pool)
        loss = crit(out, y) / total_graphs
        loss.backward()
        sdist.flat_allreduce_grads(model.parameters())
        opt.step()
        loss_buf.copy_(loss.detach())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the whole training step is launch-bound at this size: capture it once in a CUDA graph and replay it
    side = torch.cuda.Stream()
    model.train()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        l0 = handle.launch_count()
        for _ in range(max(warmup, 3)):
            step()
        launches_per_step = (handle.launch_count() - l0) // max(warmup, 3)
    barrier()
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
            step()
        run = graph.replay
    else:
        run = step
    barrier()
    with torch.cuda.stream(side):
        for _ in range(2):
            run()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    torch.cuda.current_stream().wait_stream(side)
    if rank != 0:
        return None
    return {
        "metric": "molecule_gcn_graphs_per_s", "value": total_graphs / (ms * 1e-3), "unit": "graphs/s",
        "n_gpus": world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "molecule", "graphs_per_step_per_gpu": graphs_per_rank, "nodes_per_gpu": prob.N,
                   "nnz_adj_per_gpu": prob.nnz_adj, "hidden": hidden,
                   "mode": "2-layer GCN forward on the accelerator + saved-tensor backward, Adam, DP grad all-reduce",
                   "cuda_graph": bool(use_graph)},
        "gpu_launches": int(launches_per_step * steps), "loss": float(loss_buf.item()),
    }


def bench_molecule(args):
    import json

    import torch.distributed as dist

    from . import dist as sdist

    rank, world, local = sdist.dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rec = molecule_record(args.steps, args.warmup, rank, world, local, int(getattr(args, "graphs", 0) or 188 * 64),
                          int(getattr(args, "hidden", 0) or 64))
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.destroy_process_group()
