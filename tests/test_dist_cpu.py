"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: the row partition + in-place
all-gather of XW and the data-parallel gradient all-reduce.  The accelerator calls are replaced by
torch stand-ins here (the CUDA path is covered by the -m gpu tests); what is checked is that the
sharded result equals the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sgracex1_b200 import dist as sdist
from sgracex1_b200 import graphs as G
from sgracex1_b200 import molecule_gcn as MG
from tests import util as U


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def torch_fea(x_local, W, out_slot):
    out_slot.copy_(x_local @ W)


def torch_adj(adj_local, xw_full, relu):
    rp, ci, va = adj_local
    A = torch.sparse_csr_tensor(rp.long(), ci.long(), va, size=(rp.numel() - 1, xw_full.shape[0]))
    out = A @ xw_full
    return out.relu() if relu else out


def torch_layer(handle, adj_csr, x, weight, relu):
    xw = x if weight is None else x @ weight
    return torch_adj(adj_csr, xw, relu)


def _row_partition_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pr = U.random_problem(5, n=203, m=24, p=16)
    rp, ci, va = pr["adj"]
    lo, hi = sdist.row_range(203, rank, world)
    loc = sdist.csr_row_slice(rp, ci, va, lo, hi)
    adj_local = tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in loc)
    layer = sdist.RowPartitionedLayer(203, 16, rank, world, "cpu", torch_fea, torch_adj)
    D = layer.forward(torch.from_numpy(pr["x"][lo:hi]), torch.from_numpy(pr["W"]), adj_local, 1)
    q.put((rank, lo, hi, D.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_row_partition_equals_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_row_partition_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pr = U.random_problem(5, n=203, m=24, p=16)
    A = np.zeros((203, 203), np.float32)
    rp, ci, va = pr["adj"]
    for r in range(203):
        A[r, ci[rp[r]:rp[r + 1]]] = va[rp[r]:rp[r + 1]]
    want = np.maximum(A @ (pr["x"] @ pr["W"]), 0)
    full = np.zeros_like(want)
    for rank, lo, hi, D in got:
        assert D.shape == (hi - lo, 16)
        full[lo:hi] = D
    np.testing.assert_allclose(full, want, rtol=1e-5, atol=1e-6)


def _molecule_setup(n_graphs):
    prob, batch, y = G.molecule_batch(n_graphs=n_graphs, seed=7, P=16)
    x = np.zeros((prob.N, prob.M), np.float32)
    x[np.arange(prob.N), prob.fea_col] = 1.0
    return prob, batch, y, x


def _shard(prob, batch, y, x, g0, g1):
    nodes = np.nonzero((batch >= g0) & (batch < g1))[0]
    n0, n1 = int(nodes[0]), int(nodes[-1]) + 1
    rp, ci, va = sdist.csr_row_slice(prob.adj_rowptr, prob.adj_col, prob.adj_val, n0, n1)
    return (torch.from_numpy(x[n0:n1]), (torch.from_numpy(rp), torch.from_numpy(ci - n0), torch.from_numpy(va)),
            torch.from_numpy(batch[n0:n1] - g0), torch.from_numpy(y[g0:g1].astype(np.int64)))


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_graphs = 12
    prob, batch, y, x = _molecule_setup(n_graphs)
    g0, g1 = sdist.shard_graphs(n_graphs, rank, world)
    xs, adj, bs, ys = _shard(prob, batch, y, x, g0, g1)
    model = MG.GCN_B200(16, None, layer_fn=torch_layer)
    model.eval()                                   # no dropout: results must match exactly
    out = model(xs, adj, bs, g1 - g0)
    loss = torch.nn.CrossEntropyLoss(reduction="sum")(out, ys) / n_graphs
    loss.backward()
    sdist.flat_allreduce_grads(model.parameters())
    q.put((rank, [p.grad.numpy().copy() for p in model.parameters() if p.grad is not None]))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradients_equal_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_graphs = 12
    prob, batch, y, x = _molecule_setup(n_graphs)
    xs, adj, bs, ys = _shard(prob, batch, y, x, 0, n_graphs)
    model = MG.GCN_B200(16, None, layer_fn=torch_layer)
    model.eval()
    loss = torch.nn.CrossEntropyLoss(reduction="sum")(model(xs, adj, bs, n_graphs), ys) / n_graphs
    loss.backward()
    want = [p.grad.numpy() for p in model.parameters() if p.grad is not None]
    for r in range(world):
        assert len(got[r]) == len(want)
        for a, b in zip(got[r], want):
            np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-6)


def test_partition_helpers():
    for n, w in ((10, 3), (2449029, 8), (7, 8), (16, 4)):
        blocks = [sdist.row_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        assert max(hi - lo for lo, hi in blocks) <= sdist.row_block(n, w)
        shards = [sdist.shard_graphs(n, r, w) for r in range(w)]
        assert shards[0][0] == 0 and shards[-1][1] == n
        assert max(b - a for a, b in shards) - min(b - a for a, b in shards) <= 1


def test_split_by_ownership_reassembles_the_adjacency():
    pr = U.random_problem(8, n=120, m=8, p=4, avg_deg=5)
    rp, ci, va = pr["adj"]
    world = 3
    block = sdist.row_block(120, world)
    x = np.random.default_rng(0).standard_normal((120, 8)).astype(np.float32)
    A = np.zeros((120, 120), np.float32)
    for r in range(120):
        A[r, ci[rp[r]:rp[r + 1]]] = va[rp[r]:rp[r + 1]]
    want = A @ x
    for rank in range(world):
        lo, hi = sdist.row_range(120, rank, world)
        loc = sdist.csr_row_slice(rp, ci, va, lo, hi)
        a_loc, a_rem, halo_rows = sdist.split_by_ownership(*loc, lo, hi, block)
        assert np.all((halo_rows < lo) | (halo_rows >= hi)) and np.all(np.diff(halo_rows) > 0)
        buf = np.zeros((block + len(halo_rows), 8), np.float32)
        buf[:hi - lo] = x[lo:hi]
        buf[block:] = x[halo_rows]
        got = np.zeros((hi - lo, 8), np.float32)
        for (r2, c2, v2) in (a_loc, a_rem):
            assert len(r2) == hi - lo + 1
            for r in range(hi - lo):
                got[r] += v2[r2[r]:r2[r + 1]] @ buf[c2[r2[r]:r2[r + 1]]]
        np.testing.assert_allclose(got, want[lo:hi], rtol=1e-5, atol=1e-6)


def test_halo_push_lists_fill_every_halo_slot_in_order():
    """The exchange bookkeeping of HaloLayer without a GPU: what every rank asks for (`_halo_wants`), turned into
    the owners' push lists (`_exchange_push_lists`), must deliver exactly the rows each rank's halo slots name,
    in slot order -- simulated with numpy copies in place of the copy engines."""
    world, n, width = 3, 1000, 4
    rng = np.random.default_rng(9)
    deg = rng.integers(1, 9, size=n)
    rp = np.zeros(n + 1, np.int32)
    np.cumsum(deg, out=rp[1:])
    ci = rng.integers(0, n, size=int(rp[-1])).astype(np.int32)
    va = np.ones(len(ci), np.float32)
    x = rng.standard_normal((n, width)).astype(np.float32)
    block = sdist.row_block(n, world)
    layers = []
    for r in range(world):
        lo, hi = sdist.row_range(n, r, world)
        lay = object.__new__(sdist.HaloLayer)            # the bookkeeping only: no device buffers
        lay.rank, lay.world, lay.block, lay.lo, lay.hi, lay.width = r, world, block, lo, hi, width
        loc = sdist.csr_row_slice(rp, ci, va, lo, hi)
        _, _, lay.halo_rows_np = sdist.split_by_ownership(*loc, lo, hi, block)
        lay.n_halo = len(lay.halo_rows_np)
        lay.bases = [10_000_000 * (q + 1) for q in range(world)]      # fake base addresses, one region per rank
        layers.append(lay)
    wants = [lay._halo_wants(lay.halo_rows_np)[1] for lay in layers]
    halos = [np.full((lay.n_halo, width), np.nan, np.float32) for lay in layers]
    for lay in layers:
        lay._exchange_push_lists(lay.halo_rows_np, "cpu", gather=lambda w: wants)
        rows_t, counts, dsts = lay.push
        assert sum(lay.a2a["in_splits"]) == sum(counts)
        for dest, rows, cnt, dst in zip(lay.push_ranks, rows_t, counts, dsts):
            slot0 = (dst - layers[dest].bases[dest]) // (width * 4) - block      # first halo slot this owner fills
            assert (dst - layers[dest].bases[dest]) % (width * 4) == 0 and 0 <= slot0 <= layers[dest].n_halo - cnt
            halos[dest][slot0:slot0 + cnt] = x[lay.lo + rows.numpy()]
    for lay, halo in zip(layers, halos):
        assert not np.isnan(halo).any()
        assert np.array_equal(halo, x[lay.halo_rows_np])
        assert lay.a2a["out_splits"][lay.rank] == 0 and sum(lay.a2a["out_splits"]) == lay.n_halo


@pytest.mark.parametrize("n_chunks", [1, 2, 4])
def test_chunked_halo_plan_reassembles_the_layer(n_chunks):
    """plan_halo_chunks / chunk_wants / chunk_push_lists (the pipelined exchange, host side): with the pushes of
    chunks 0..c delivered, the owned pass plus the remote pass of chunk c must give the rows of chunk c of A.X --
    i.e. a chunk never reads a halo slot that arrives later."""
    world, n, width = 3, 900, 5
    rng = np.random.default_rng(21)
    deg = rng.integers(0, 10, size=n)
    rp = np.zeros(n + 1, np.int32)
    np.cumsum(deg, out=rp[1:])
    ci = rng.integers(0, n, size=int(rp[-1])).astype(np.int32)
    va = rng.uniform(-1, 1, size=len(ci)).astype(np.float32)
    x = rng.standard_normal((n, width)).astype(np.float64)
    A = np.zeros((n, n))
    np.add.at(A, (np.repeat(np.arange(n), deg), ci), va.astype(np.float64))
    want = A @ x
    block = sdist.row_block(n, world)
    plans, wants = [], []
    for r in range(world):
        lo, hi = sdist.row_range(n, r, world)
        plans.append(sdist.plan_halo_chunks(*sdist.csr_row_slice(rp, ci, va, lo, hi), lo, hi, block, n_chunks))
        wants.append(sdist.chunk_wants(plans[-1], r, world, block))
    pushes = [sdist.chunk_push_lists(wants, r, sdist.row_range(n, r, world)[0], n_chunks) for r in range(world)]

    def dense(csr, ncols):
        crp, cci, cva = csr
        m = np.zeros((len(crp) - 1, ncols))
        np.add.at(m, (np.repeat(np.arange(len(crp) - 1), np.diff(crp)), cci), cva.astype(np.float64))
        return m

    for r in range(world):
        lo, hi = sdist.row_range(n, r, world)
        plan = plans[r]
        n_halo = len(plan["halo_rows"])
        assert plan["chunk_slots"][0] == 0 and plan["chunk_slots"][-1] == n_halo
        buf = np.full((block + n_halo, width), np.nan)
        buf[:hi - lo] = x[lo:hi]
        buf[hi - lo:block] = 0.0
        t = dense(plan["a_loc"], block + n_halo)[:, :block] @ np.nan_to_num(buf[:block])
        for c in range(n_chunks):
            for s in range(world):                          # deliver what every owner sends for chunk c
                if s == r:
                    continue
                slo = sdist.row_range(n, s, world)[0]
                for dest, slot0, rows in pushes[s][c]:
                    if dest == r:
                        assert plan["chunk_slots"][c] <= slot0 and slot0 + len(rows) <= plan["chunk_slots"][c + 1]
                        buf[block + slot0:block + slot0 + len(rows)] = x[slo + rows]
            r0, r1 = int(plan["cuts"][c]), int(plan["cuts"][c + 1])
            rem = dense(plan["a_rem"][c], block + n_halo)
            used = np.abs(rem).sum(axis=0) > 0
            assert not np.isnan(buf[used]).any(), f"rank {r} chunk {c} reads a halo slot that has not arrived"
            got = t[r0:r1] + rem @ np.nan_to_num(buf)
            np.testing.assert_allclose(got, want[lo + r0:lo + r1], rtol=1e-9, atol=1e-9)
        assert not np.isnan(buf[block:]).any() and np.array_equal(buf[block:], x[plan["halo_rows"]])
