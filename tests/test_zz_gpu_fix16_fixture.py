"""FIX16 C-simulation mode on the GPU against outputs of the reference's own source compiled in its EIGHTBIT
configuration (tests/golden/ref_hls_fix16.npz, made by tests/golden/make_golden.py:ref_hls_fix16 from
oracle/_ref/libsgrace_hlsref_fix16.so): int16 Q2.14 codes, bit for bit, including the set whose sums wrap around.
Reference: mmult_top, gnn-rfsoc-mt-all-2022/src/kernelMatrixmult_all.cpp:3762-3967 with matrix_mult.h:105-109 types."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from sgracex1_b200 import _lib
from tests import util as U
from tests.test_gpu_parity import run_host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ip():
    from sgracex1_b200.pynq_compat import MmultTop
    return MmultTop(0)


@pytest.mark.parametrize("tag", ["small", "wrap"])
def test_gpu_fix16_matches_compiled_reference_fixtures(ip, tag):
    g = np.load(os.path.join(U.GOLDEN, "ref_hls_fix16.npz"))
    N, M = int(g[f"{tag}_N"]), int(g[f"{tag}_M"])
    adj = (g[f"{tag}_adj_rowptr"], g[f"{tag}_adj_col"], g[f"{tag}_adj_val"])
    fea = (g[f"{tag}_fea_rowptr"], g[f"{tag}_fea_col"], g[f"{tag}_fea_val"])
    for P in (16, 7):
        for relu in (0, 1):
            D = run_host(ip, _lib.MODE_FIX16_CSIM, N=N, M=M, P=P, adj=adj, fea=fea, B=g[f"{tag}_B_P{P}"], relu=relu, lat_fea=1, lat_adj=1)
            assert np.array_equal(D.view(np.uint8), g[f"{tag}_sparse_P{P}_relu{relu}"].view(np.uint8)), (tag, P, relu)
    D = run_host(ip, _lib.MODE_FIX16_CSIM, N=N, M=24, P=10, adj=adj, x_dense=g[f"{tag}_x_dense"], B=g[f"{tag}_B_dense"], relu=1,
                 lat_fea=1, lat_adj=1)
    assert np.array_equal(D.view(np.uint8), g[f"{tag}_dense_P10_relu1"].view(np.uint8)), tag
    # and the oracle gives the same codes (the CPU suite checks this too)
    ref = O.layer(dtype=O.FIX16, N=N, M_fea=24, P=10, adj=adj, x_dense=g[f"{tag}_x_dense"], B=g[f"{tag}_B_dense"], relu=1,
                  spmm_block=1, lat_fea=1, lat_adj=1)
    assert np.array_equal(D.view(np.uint8), ref.view(np.uint8))
