"""Seeded synthetic graphs of the benchmark shapes.

Everything is numpy on the host; nothing here reads the reference checkout.  Shapes:
  cora_shape      N=2708,  M=1433, nnz_adj~13264,  nnz_fea=49216, P=16
  citeseer_shape  N=3327,  M=3703, nnz_adj~12431,  nnz_fea~105165
  pubmed_shape    N=19717, M=500,  nnz_adj~108365, nnz_fea~988031
  products_shape  N=2449029, nnz_adj~61.86M, M=100 dense, P=256 (row-partitionable)
  molecule batch  block-diagonal batch of small molecular graphs (MUTAG-like: 17.9 nodes,
                  2.2 average degree, 7 one-hot node labels, 0/1 adjacency, no self-loops)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# degree histogram of the Cora adjacency with self-loops (count of rows per degree 0..45; the
# tail up to 169 is re-created by HUBS below).  Statistics only: min 2, mean 4.9, max 169.
_CORA_DEG_HIST = [0, 0, 485, 583, 553, 389, 281, 131, 82, 57, 25, 26, 14, 18, 5, 6, 6, 7, 8, 3, 5, 0,
                  3, 1, 3, 0, 0, 1, 0, 0, 1, 2, 1, 2, 1, 1, 0, 1, 0, 0, 0, 1, 0, 1, 0, 1]
_CORA_HUBS = [66, 75, 79, 169]


@dataclass
class LayerProblem:
    """One layer's inputs in the accelerator's buffer formats (float32 values)."""
    N: int
    M: int
    P: int
    adj_rowptr: np.ndarray
    adj_col: np.ndarray
    adj_val: np.ndarray
    fea_rowptr: np.ndarray | None
    fea_col: np.ndarray | None
    fea_val: np.ndarray | None       # CSR values, or dense N*M row-major when x_dense
    x_dense: bool
    W: np.ndarray                    # M x P
    n_graphs: int = 1

    @property
    def nnz_adj(self):
        return int(self.adj_rowptr[-1])

    @property
    def nnz_fea(self):
        return int(self.fea_rowptr[-1]) if self.fea_rowptr is not None else self.N * self.M

    @property
    def B(self):
        """B buffer contents: W transposed, P x M row-major, flattened."""
        return np.ascontiguousarray(self.W.T).reshape(-1)

    def algorithmic_bytes(self, elt=4):
        """BASELINE.md section 3 definitions."""
        N, M, P = self.N, self.M, self.P
        if self.x_dense:
            fea = N * M * elt + M * P * elt + N * P * elt
        else:
            fea = (N + 1) * 4 + self.nnz_fea * (4 + elt) + M * P * elt + N * P * elt
        adj = (N + 1) * 4 + self.nnz_adj * (4 + elt) + 2 * N * P * elt
        return {"fea": fea, "adj": adj, "layer": fea + adj}


def _sym_norm_csr(n, rows, cols):
    """D^-1/2 (A) D^-1/2 for a symmetric 0/1 pattern that already holds the self-loops."""
    key = np.unique(rows.astype(np.int64) * n + cols.astype(np.int64))
    r = (key // n).astype(np.int32)
    c = (key % n).astype(np.int32)
    deg = np.bincount(r, minlength=n).astype(np.float64)
    dinv = np.where(deg > 0, deg ** -0.5, 0.0)
    val = (dinv[r] * dinv[c]).astype(np.float32)
    rowptr = np.zeros(n + 1, np.int32)
    np.cumsum(np.bincount(r, minlength=n), out=rowptr[1:])
    return rowptr, c, val


def _degree_sequence(n, hist, hubs, rng):
    degs = np.repeat(np.arange(len(hist)), hist)
    seq = rng.choice(degs, size=n, replace=True)
    if n >= 8 * len(hubs):
        idx = rng.choice(n, size=len(hubs), replace=False)
        seq[idx] = hubs
    return seq


def random_sym_graph(n, target_nnz, rng, hist=_CORA_DEG_HIST, hubs=_CORA_HUBS, locality=None):
    """Symmetric graph with self-loops whose degree sequence follows `hist` (+ `hubs`)."""
    seq = _degree_sequence(n, hist, [h for h in hubs if h < n], rng).astype(np.int64) - 1  # minus self-loop
    seq = np.maximum(seq, 0)
    # each undirected edge is drawn once from half of the stubs; scale to hit the target count
    want_edges = max(0, (target_nnz - n) // 2)
    p = seq / max(1, seq.sum())
    src = rng.choice(n, size=want_edges, p=p)
    if locality:
        dst = (src + rng.integers(-locality, locality + 1, size=want_edges)) % n
    else:
        dst = rng.choice(n, size=want_edges, p=p)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    loops = np.arange(n)
    rows = np.concatenate([src, dst, loops])
    cols = np.concatenate([dst, src, loops])
    return _sym_norm_csr(n, rows, cols)


def random_sparse_features(n, m, nnz, rng, lo=1, hi=30, binary=True):
    mean = nnz / n
    cnt = np.clip(rng.binomial(max(hi, int(2 * mean)), min(1.0, mean / max(hi, int(2 * mean))), size=n), lo, min(hi, m))
    # nudge to the requested total
    diff = int(nnz - cnt.sum())
    while diff != 0:
        idx = rng.integers(0, n, size=abs(diff))
        step = 1 if diff > 0 else -1
        for i in idx:
            if lo <= cnt[i] + step <= min(hi, m):
                cnt[i] += step
                diff -= step
                if diff == 0:
                    break
    rowptr = np.zeros(n + 1, np.int32)
    np.cumsum(cnt, out=rowptr[1:])
    total = int(rowptr[-1])
    # distinct sorted columns per row: draw, then de-duplicate by re-drawing collisions
    col = rng.integers(0, m, size=total).astype(np.int32)
    rows = np.repeat(np.arange(n), cnt)
    for _ in range(8):
        order = np.lexsort((col, rows))
        col, rows = col[order], rows[order]
        dup = np.zeros(total, bool)
        dup[1:] = (rows[1:] == rows[:-1]) & (col[1:] == col[:-1])
        if not dup.any():
            break
        col[dup] = rng.integers(0, m, size=int(dup.sum()))
    order = np.lexsort((col, rows))
    col = col[order]
    val = np.ones(total, np.float32) if binary else rng.random(total, dtype=np.float32)
    return rowptr, col, val


def cora_shape(seed=0, P=16, n=2708, m=1433, nnz_adj=13264, nnz_fea=49216) -> LayerProblem:
    rng = np.random.default_rng(seed)
    arp, aci, ava = random_sym_graph(n, nnz_adj, rng)
    frp, fci, fva = random_sparse_features(n, m, nnz_fea, rng)
    W = rng.uniform(-0.25, 0.25, size=(m, P)).astype(np.float32)
    return LayerProblem(n, m, P, arp, aci, ava, frp, fci, fva, False, W)


def citeseer_shape(seed=1, P=16) -> LayerProblem:
    rng = np.random.default_rng(seed)
    n, m = 3327, 3703
    arp, aci, ava = random_sym_graph(n, 12431, rng, hubs=[100])
    frp, fci, fva = random_sparse_features(n, m, 105165, rng, lo=1, hi=54)
    W = np.clip(rng.uniform(-0.25, 0.25, size=(m, P)), -1, 1).astype(np.float32)
    return LayerProblem(n, m, P, arp, aci, ava, frp, fci, fva, False, W)


def pubmed_shape(seed=1, P=16) -> LayerProblem:
    rng = np.random.default_rng(seed)
    n, m = 19717, 500
    arp, aci, ava = random_sym_graph(n, 108365, rng, hubs=[172, 150, 120])
    frp, fci, fva = random_sparse_features(n, m, 988031, rng, lo=10, hi=120, binary=False)
    W = np.clip(rng.uniform(-0.25, 0.25, size=(m, P)), -1, 1).astype(np.float32)
    return LayerProblem(n, m, P, arp, aci, ava, frp, fci, fva, False, W)


def block_diagonal(problems, copies: int) -> LayerProblem:
    """Batch `copies` graphs (cycling through `problems`) as one block-diagonal layer -- how
    the reference batches molecule graphs (Graph_Classification.ipynb cell 10 output) and the
    bandwidth variant of the Cora-shape benchmark."""
    p0 = problems[0]
    arp, aci, ava, frp, fci, fva = [np.zeros(1, np.int64)], [], [], [np.zeros(1, np.int64)], [], []
    n_off = a_off = f_off = 0
    dense = p0.x_dense
    for k in range(copies):
        p = problems[k % len(problems)]
        arp.append(p.adj_rowptr[1:].astype(np.int64) + a_off)
        aci.append(p.adj_col.astype(np.int64) + n_off)
        ava.append(p.adj_val)
        if not dense:
            frp.append(p.fea_rowptr[1:].astype(np.int64) + f_off)
            fci.append(p.fea_col)
            f_off += p.nnz_fea
        fva.append(p.fea_val)
        n_off += p.N
        a_off += p.nnz_adj
    if max(n_off, a_off, f_off) >= 2 ** 31:
        raise ValueError("batch exceeds int32 indexing")
    return LayerProblem(
        n_off, p0.M, p0.P,
        np.concatenate(arp).astype(np.int32), np.concatenate(aci).astype(np.int32), np.concatenate(ava),
        None if dense else np.concatenate(frp).astype(np.int32),
        None if dense else np.concatenate(fci).astype(np.int32),
        np.concatenate(fva), dense, p0.W, n_graphs=copies)


def molecule_graph(rng, n_lo=10, n_hi=28, n_labels=7):
    """One MUTAG-like molecule: a random tree plus a few ring-closing bonds, degree <= 4,
    0/1 adjacency without self-loops, one-hot node labels."""
    n = int(rng.integers(n_lo, n_hi + 1))
    deg = np.zeros(n, np.int64)
    edges = set()
    for v in range(1, n):
        for _ in range(16):
            u = int(rng.integers(max(0, v - 6), v))
            if deg[u] < 3:
                break
        edges.add((u, v))
        deg[u] += 1
        deg[v] += 1
    for _ in range(max(1, n // 6)):
        u, v = sorted(int(x) for x in rng.integers(0, n, size=2))
        if u != v and (u, v) not in edges and deg[u] < 4 and deg[v] < 4 and v - u >= 3:
            edges.add((u, v))
            deg[u] += 1
            deg[v] += 1
    e = np.array(sorted(edges), np.int64).reshape(-1, 2)
    rows = np.concatenate([e[:, 0], e[:, 1]])
    cols = np.concatenate([e[:, 1], e[:, 0]])
    labels = rng.choice(n_labels, size=n, p=np.array([0.66, 0.1, 0.18, 0.02, 0.02, 0.01, 0.01]))
    return n, rows, cols, labels


def molecule_batch(n_graphs=188, seed=12345, P=64, n_labels=7, dense_features=False):
    """Block-diagonal batch of molecule graphs + one-hot features (sparse CSR, one nnz per row)."""
    rng = np.random.default_rng(seed)
    rows, cols, labels, batch = [], [], [], []
    off = 0
    for g in range(n_graphs):
        n, r, c, lab = molecule_graph(rng, n_labels=n_labels)
        rows.append(r + off)
        cols.append(c + off)
        labels.append(lab)
        batch.append(np.full(n, g, np.int64))
        off += n
    N = off
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    arp = np.zeros(N + 1, np.int32)
    np.cumsum(np.bincount(rows, minlength=N), out=arp[1:])
    labels = np.concatenate(labels)
    frp = np.arange(N + 1, dtype=np.int32)
    stdv = 1.0 / np.sqrt(P)
    W = rng.uniform(-stdv, stdv, size=(n_labels, P)).astype(np.float32)
    prob = LayerProblem(N, n_labels, P, arp, cols.astype(np.int32), np.ones(len(cols), np.float32),
                        frp, labels.astype(np.int32), np.ones(N, np.float32), False, W, n_graphs=n_graphs)
    y = rng.integers(0, 2, size=n_graphs)
    return prob, np.concatenate(batch), y


def products_shape_rows(row_begin, row_end, n_total=2_449_029, mean_deg=25.26, seed=2, block=4096,
                        far_frac=0.10, max_deg=17000):
    """Rows [row_begin, row_end) of the ogbn-products-shape adjacency (global column indices).

    Row-wise generator so that every rank can build its own partition without materialising
    the whole graph: degrees are Zipf-like (mean ~25.3, max ~17k); 90% of a row's neighbours
    fall in the row's block of `block` ids, 10% anywhere.  Values 1/deg (row-normalised).
    The pattern is not symmetrised (the ADJ kernel does not care)."""
    n = row_end - row_begin
    rng = np.random.default_rng([seed, row_begin])
    # Pareto-tailed degrees, scaled to the requested mean
    raw = (rng.pareto(1.6, size=n) + 1.0)
    deg = np.minimum(np.maximum(1, np.round(raw * mean_deg / 2.6)), max_deg).astype(np.int64)
    rowptr = np.zeros(n + 1, np.int64)
    np.cumsum(deg, out=rowptr[1:])
    total = int(rowptr[-1])
    rows = np.repeat(np.arange(row_begin, row_end, dtype=np.int64), deg)
    local = rng.random(total) >= far_frac
    blk0 = (rows // block) * block
    col = np.where(local, blk0 + rng.integers(0, block, size=total), rng.integers(0, n_total, size=total))
    col = np.minimum(col, n_total - 1)
    order = np.lexsort((col, rows))
    col = col[order].astype(np.int32)
    val = np.repeat((1.0 / deg).astype(np.float32), deg)
    if total >= 2 ** 31:
        raise ValueError("partition exceeds int32 indexing")
    return rowptr.astype(np.int32), col, val
