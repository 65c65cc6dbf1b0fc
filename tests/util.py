"""Shared helpers for the tests (CPU and GPU)."""
import os

import numpy as np

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def coo_to_csr(ei, val, n):
    r, c = np.asarray(ei[0]), np.asarray(ei[1])
    assert np.all(np.diff(r.astype(np.int64) * n + c) > 0), "COO input must be sorted and unique"
    rp = np.zeros(n + 1, np.int32)
    np.cumsum(np.bincount(r, minlength=n), out=rp[1:])
    return rp, c.astype(np.int32), np.asarray(val, np.float32)


def dense_to_csr(x):
    """x._to_sparse_csr() / x.to_sparse(): drops exact zeros, row-major order."""
    n, _ = x.shape
    mask = x != 0
    rp = np.zeros(n + 1, np.int32)
    np.cumsum(mask.sum(1), out=rp[1:])
    return rp, np.nonzero(mask)[1].astype(np.int32), x[mask].astype(np.float32)


def citeseer_half():
    g = np.load(os.path.join(GOLDEN, "citeseer_half.npz"))
    n = len(g["adj_rowptr"]) - 1
    adj = (g["adj_rowptr"], g["adj_col"].astype(np.int32), g["adj_val_f16"].view(np.uint16))
    nnzf = int(g["fea_rowptr"][-1])
    fea = (g["fea_rowptr"], g["fea_col"].astype(np.int32), np.full(nnzf, 0x3C00, np.uint16))   # 1.0
    w16 = g["w_f16"]
    return g, n, adj, fea, w16


def random_problem(seed, n=200, m=64, p=16, dens_x=0.15, avg_deg=4, val_scale=0.5, empty_rows=True):
    """Small seeded layer problem with ragged rows, empty rows and duplicate-free CSR."""
    rng = np.random.default_rng(seed)
    deg = rng.poisson(avg_deg, size=n)
    deg[rng.integers(0, n, size=max(1, n // 50))] = min(n, 40)          # a few long rows
    if empty_rows:
        deg[rng.integers(0, n, size=max(1, n // 20))] = 0
    deg = np.minimum(deg, n)
    rp = np.zeros(n + 1, np.int32)
    np.cumsum(deg, out=rp[1:])
    ci = np.concatenate([np.sort(rng.choice(n, size=d, replace=False)) for d in deg] + [np.zeros(0, np.int64)])
    av = rng.uniform(-val_scale, val_scale, size=len(ci)).astype(np.float32)
    x = ((rng.random((n, m)) < dens_x) * rng.uniform(-1, 1, size=(n, m))).astype(np.float32)
    if empty_rows:
        x[rng.integers(0, n, size=3)] = 0
    w = rng.uniform(-0.5, 0.5, size=(m, p)).astype(np.float32)
    return dict(N=n, M=m, P=p, adj=(rp, ci.astype(np.int32), av), x=x, fea=dense_to_csr(x), W=w)


def to_storage_problem(pr, dtype):
    adj = (pr["adj"][0], pr["adj"][1], O.to_storage(pr["adj"][2], dtype))
    fea = (pr["fea"][0], pr["fea"][1], O.to_storage(pr["fea"][2], dtype))
    B = O.to_storage(O.weights_to_B(pr["W"]), dtype)
    xd = O.to_storage(pr["x"], dtype)
    return adj, fea, B, xd


ELEMENTWISE_LOG = []      # (what, worst plain relative error over the significant elements, their count): printed at session end


def assert_close_f32(got, want, rtol=1e-5, what="", sig=1e-3, elem_rtol=2e-3):
    """Float tolerance of the north star: 1e-5 relative.  Elements that suffer cancellation are judged against the
    magnitude of their row (atol = rtol * max|row|).  Beside that row-normalised bar, the PLAIN element-wise relative
    error |got - want| / |want| is computed over every element of at least `sig` of its row maximum, recorded in
    ELEMENTWISE_LOG (shown in the pytest summary) and held to `elem_rtol`.  An absolute error of 1e-5 of the row maximum
    is a relative error of up to 1e-2 on an element that is 1e-3 of it, so the element-wise bar is necessarily looser
    than the row-normalised one; measured worst case over the GPU suite: 3.7e-4 (a 256-wide row with cancellation)."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = np.maximum(np.abs(want).max(axis=-1, keepdims=True), 1e-30) if want.ndim > 1 else np.abs(want).max() + 1e-30
    err = np.abs(got - want)
    bad = err > rtol * np.maximum(np.abs(want), scale)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} elements beyond {rtol} rel; max err {err.max():.3e}"
    significant = np.abs(want) >= sig * scale
    if significant.any():
        rel = err[significant] / np.abs(want[significant])
        ELEMENTWISE_LOG.append((what, float(rel.max()), int(significant.sum())))
        assert rel.max() <= elem_rtol, f"{what}: plain element-wise relative error {rel.max():.3e} > {elem_rtol} on elements >= {sig} of the row max"
