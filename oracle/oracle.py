"""ctypes binding of the CPU oracle (oracle/sgrace_oracle.c).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from sgracex1_b200/.

Also holds the loaders for the reference's fixture formats:
  * 3-line CSR text  (rowPtr / columnIndex / values, comma separated) --
    reference gnn-rfsoc-mt-all-2022/src/main_float.cpp:415-536 (loadcsr_adj) and
    :538-659 (loadcsr_fea), jupyter/test/mmult-master.ipynb cells 25-26;
  * dense weight text, one matrix row per line, stored TRANSPOSED into the B
    buffer -- main_float.cpp:138-199 (`A[i + j*N]`), mmult-master.ipynb cell 18.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

F32, F16, FIX16 = 0, 1, 2
_NP = {F32: np.float32, F16: np.uint16, FIX16: np.int16}


class _Layer(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "dtype", "gemm_mode", "relu", "N_adj", "M_adj", "M_fea", "P_w", "spmm_block", "lat_fea",
        "lat_adj", "fea_threads", "adj_threads", "b_width_block", "use_sblocks")] + [
        (n, C.c_void_p) for n in (
            "rowPtr_fea", "columnIndex_fea", "values_fea", "rowPtr_adj", "columnIndex_adj",
            "values_adj", "B", "D", "XW")]


class _QLayer(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "gemm_mode", "gat_mode", "relu", "qbits", "N_adj", "M_fea", "P_w", "scale_fea",
        "internal_quantization")] + [
        ("qscale_fea", C.c_float), ("qscale_w", C.c_float), ("qscale_adj", C.c_float),
        ("f_z", C.c_int), ("w_z", C.c_int), ("a_z", C.c_int),
        ("deq_o", C.c_float), ("alpha", C.c_float)] + [
        (n, C.c_void_p) for n in (
            "rowPtr_fea", "columnIndex_fea", "values_fea", "rowPtr_adj", "columnIndex_adj",
            "values_adj", "B", "attention", "D", "E", "S", "Wh", "max_fea")]


def build(force: bool = False) -> str:
    """Compile oracle/libsgrace_oracle.so (and oracle/_ref when the reference is present)."""
    so = os.path.join(_HERE, "libsgrace_oracle.so")
    src = os.path.join(_HERE, "sgrace_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "libsgrace_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.sgrace_oracle_layer.argtypes = [C.POINTER(_Layer)]
        _LIB.sgrace_oracle_qlayer.argtypes = [C.POINTER(_QLayer)]
        _LIB.sgrace_oracle_layer_batch.argtypes = [C.POINTER(_Layer), C.c_int, C.c_int]
        _LIB.sgo_f32_to_f16.argtypes = [C.c_float]
        _LIB.sgo_f32_to_f16.restype = C.c_uint16
        _LIB.sgo_f16_to_f32.argtypes = [C.c_uint16]
        _LIB.sgo_f16_to_f32.restype = C.c_float
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


# ----------------------------------------------------------------------------
# storage-type conversion (what `(BTYPE)val` does in main_float.cpp:178,486,614)
# ----------------------------------------------------------------------------
def to_storage(x, dtype):
    x = np.asarray(x)
    if dtype == F32:
        return x.astype(np.float32)
    if dtype == F16:
        return x.astype(np.float32).astype(np.float16).view(np.uint16)
    if dtype == FIX16:
        # ap_fixed<16,2>(double): AP_TRN drops the bits below 2^-14 (toward -inf), AP_WRAP
        v = np.floor(x.astype(np.float64) * 16384.0).astype(np.int64)
        return (v & 0xFFFF).astype(np.uint16).view(np.int16)
    raise ValueError(dtype)


def from_storage(a, dtype):
    if dtype == F32:
        return np.asarray(a, np.float32)
    if dtype == F16:
        return np.asarray(a).view(np.float16).astype(np.float32)
    if dtype == FIX16:
        return np.asarray(a, np.int16).astype(np.float32) / 16384.0
    raise ValueError(dtype)


# ----------------------------------------------------------------------------
# fixture loaders
# ----------------------------------------------------------------------------
@dataclass
class Csr:
    rowptr: np.ndarray
    col: np.ndarray
    val: np.ndarray  # float32 as read

    @property
    def n(self):
        return len(self.rowptr) - 1

    @property
    def nnz(self):
        return int(self.rowptr[-1])


def load_csr_txt(path) -> Csr:
    with open(path) as f:
        lines = [ln.strip().rstrip(",") for ln in f.readlines() if ln.strip()]
    rp = np.array(lines[0].split(","), dtype=np.int64).astype(np.int32)
    ci = np.array(lines[1].split(","), dtype=np.int64).astype(np.int32)
    va = np.array(lines[2].split(","), dtype=np.float64).astype(np.float32)
    return Csr(rp, ci, va)


def load_dense_txt(path) -> np.ndarray:
    return np.loadtxt(path, delimiter=",", dtype=np.float32, ndmin=2)


def weights_to_B(w: np.ndarray) -> np.ndarray:
    """W (M x P) -> B buffer contents (P x M row-major, i.e. W transposed, flattened)."""
    return np.ascontiguousarray(np.asarray(w).T).reshape(-1)


# ----------------------------------------------------------------------------
# the two oracles
# ----------------------------------------------------------------------------
def layer(*, dtype, N, M_fea, P, adj, B, fea=None, x_dense=None, relu=0, spmm_block=1,
          lat_fea=None, lat_adj=None, fea_threads=1, adj_threads=1, b_width_block=2,
          use_sblocks=0, return_xw=False):
    """Run sgrace_oracle_layer.  adj/fea are (rowptr, col, values) triples ALREADY in
    storage type; x_dense is an N x M array in storage type (gemm_mode=1); B is the
    transposed weight buffer in storage type."""
    npdt = _NP[dtype]
    default_lat = {F32: 6, F16: 4, FIX16: 1}[dtype]       # matrix_mult.h:117-150
    d = _Layer()
    d.dtype, d.relu, d.N_adj, d.M_adj, d.M_fea, d.P_w = dtype, int(relu), N, N, M_fea, P
    d.spmm_block, d.fea_threads, d.adj_threads = spmm_block, fea_threads, adj_threads
    d.lat_fea = default_lat if lat_fea is None else lat_fea
    d.lat_adj = default_lat if lat_adj is None else lat_adj
    d.b_width_block, d.use_sblocks = b_width_block, use_sblocks
    keep = []
    if x_dense is not None:
        d.gemm_mode = 1
        xv = _c(x_dense, npdt).reshape(-1)
        keep.append(xv)
        d.values_fea = _p(xv)
    else:
        d.gemm_mode = 0
        rp, ci, va = _c(fea[0], np.int32), _c(fea[1], np.int32), _c(fea[2], npdt)
        keep += [rp, ci, va]
        d.rowPtr_fea, d.columnIndex_fea, d.values_fea = _p(rp), _p(ci), _p(va)
    rp, ci, va = _c(adj[0], np.int32), _c(adj[1], np.int32), _c(adj[2], npdt)
    Bc = _c(B, npdt)
    keep += [rp, ci, va, Bc]
    d.rowPtr_adj, d.columnIndex_adj, d.values_adj, d.B = _p(rp), _p(ci), _p(va), _p(Bc)
    D = np.zeros((N, P), dtype=npdt)
    XW = np.zeros((N, P), dtype=npdt)
    d.D, d.XW = _p(D), _p(XW)
    rc = lib().sgrace_oracle_layer(C.byref(d))
    if rc != 0:
        raise RuntimeError(f"sgrace_oracle_layer failed: {rc}")
    return (D, XW) if return_xw else D


def layer_batch(descs, threads):
    arr = (_Layer * len(descs))(*descs)
    rc = lib().sgrace_oracle_layer_batch(arr, len(descs), int(threads))
    if rc != 0:
        raise RuntimeError("sgrace_oracle_layer_batch failed")


def max_threads() -> int:
    return int(lib().sgrace_oracle_max_threads())


def qlayer(*, N, M_fea, P, adj, B, fea=None, x_dense=None, attention=None, relu=0, gat=0,
           qbits=8, consts=None, alpha=0.2, return_all=False):
    """Run sgrace_oracle_qlayer.  `consts` = dict(f_s,f_z,w_s,w_z,a_s,a_z,scale_fea,
    internal_quantization,deq_o) -- see sgracex1_b200.quant.layer_constants()."""
    d = _QLayer()
    d.gat_mode, d.relu, d.qbits, d.N_adj, d.M_fea, d.P_w = int(gat), int(relu), qbits, N, M_fea, P
    if qbits:
        d.scale_fea = consts["scale_fea"]
        d.internal_quantization = consts["internal_quantization"]
        d.qscale_fea = float(np.float32(1.0 / consts["f_s"]))
        d.qscale_w = float(np.float32(1.0 / consts["w_s"]))
        d.qscale_adj = float(np.float32(1.0 / consts["a_s"]))
        d.f_z, d.w_z, d.a_z = consts["f_z"], consts["w_z"], consts["a_z"]
        d.deq_o = consts["deq_o"]
    d.alpha = alpha
    keep = []
    if x_dense is not None:
        d.gemm_mode = 1
        xv = _c(x_dense, np.float32).reshape(-1)
        keep.append(xv)
        d.values_fea = _p(xv)
    else:
        rp, ci, va = _c(fea[0], np.int32), _c(fea[1], np.int32), _c(fea[2], np.float32)
        keep += [rp, ci, va]
        d.rowPtr_fea, d.columnIndex_fea, d.values_fea = _p(rp), _p(ci), _p(va)
    rp, ci, va = _c(adj[0], np.int32), _c(adj[1], np.int32), _c(adj[2], np.float32)
    Bc = _c(B, np.float32)
    keep += [rp, ci, va, Bc]
    d.rowPtr_adj, d.columnIndex_adj, d.values_adj, d.B = _p(rp), _p(ci), _p(va), _p(Bc)
    if gat:
        at = _c(np.asarray(attention).reshape(-1), np.float32)
        keep.append(at)
        d.attention = _p(at)
    nnz = int(rp[-1])
    D = np.zeros((N, P), np.float32)
    E = np.zeros(max(nnz, 1), np.float32)
    S = np.zeros(max(nnz, 1), np.float32)
    Wh = np.zeros((N, P), np.float32)
    mf = np.zeros(1, np.int32)
    d.D, d.E, d.S, d.Wh, d.max_fea = _p(D), _p(E), _p(S), _p(Wh), _p(mf)
    rc = lib().sgrace_oracle_qlayer(C.byref(d))
    if rc != 0:
        raise RuntimeError(f"sgrace_oracle_qlayer failed: {rc}")
    if return_all:
        return dict(D=D, E=E[:nnz], S=S[:nnz], Wh=Wh, max_fea=int(mf[0]))
    return D


# ----------------------------------------------------------------------------
# oracle/_ref: the reference's own kernelMatrixmult_all.cpp compiled natively against
# oracle/hls_shim/ (see oracle/Makefile).  Fixed build-time configuration: HALF (or float
# via the shim), SPMM_BLOCK=1, FADD latency 4, 1 FEA / 1 ADJ thread, N,M <= 6144.
# ----------------------------------------------------------------------------
_REF = {}


def ref_available(kind="half") -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", f"libsgrace_hlsref_{kind}.so"))


def ref_lib(kind="half"):
    if kind not in _REF:
        L = C.CDLL(os.path.join(_HERE, "_ref", f"libsgrace_hlsref_{kind}.so"))
        L.sgrace_ref_mmult_top.argtypes = [C.c_int] * 6 + [C.c_void_p] * 8
        _REF[kind] = L
    return _REF[kind]


def ref_layer(*, kind, N, M_fea, P, adj, B, fea=None, x_dense=None, relu=0):
    """Call the reference's mmult_top.  kind='half' -> uint16 storage, 'float' -> float32, 'fix16' -> int16 Q2.14 codes
    (the EIGHTBIT configuration of the source with the ap_fixed<16,2> stand-in of oracle/hls_shim/ap_int.h)."""
    npdt = {"half": np.uint16, "fix16": np.int16}.get(kind, np.float32)
    L = ref_lib(kind)
    rp, ci, va = _c(adj[0], np.int32), _c(adj[1], np.int32), _c(adj[2], npdt)
    Bc = _c(B, npdt)
    D = np.zeros((N, P), dtype=npdt)
    if x_dense is not None:
        xv = _c(x_dense, npdt).reshape(-1)
        frp = np.zeros(1, np.int32)
        fci = np.zeros(1, np.int32)
        gm = 1
    else:
        frp, fci, xv = _c(fea[0], np.int32), _c(fea[1], np.int32), _c(fea[2], npdt)
        gm = 0
    rc = L.sgrace_ref_mmult_top(gm, int(relu), N, N, M_fea, P, _p(Bc), _p(D), _p(frp), _p(fci),
                                _p(xv), _p(rp), _p(ci), _p(va))
    if rc != 0:
        raise RuntimeError(f"sgrace_ref_mmult_top failed: {rc}")
    return D
