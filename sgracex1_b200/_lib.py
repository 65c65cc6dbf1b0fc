"""ctypes binding of libsgrace_b200.so (include/sgrace_b200.h).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsgrace_b200.so")

# enums of include/sgrace_b200.h
MODE_F32_FAST, MODE_F32_CSIM, MODE_F16_CSIM, MODE_FIX16_CSIM, MODE_FULL = 0, 1, 2, 3, 4
(OPT_MODE, OPT_SPMM_BLOCK, OPT_LAT_FEA, OPT_LAT_ADJ, OPT_FEA_THREADS, OPT_ADJ_THREADS,
 OPT_USE_SBLOCKS, OPT_INDEX_FORMAT, OPT_QBITS, OPT_STAGING, OPT_LONG_ROW, OPT_LEAKY_ALPHA_BITS,
 OPT_VALIDATE, OPT_DENSE_TC, OPT_STREAM_KERNEL, OPT_AGG_FIRST, OPT_ACCUMULATE, OPT_FUSED_SMALL, OPT_ROW_OFFSET,
 OPT_ADJ_PLAN, OPT_PANEL_LAUNCHES, OPT_PLAN_BUILDS, OPT_OVERLAP, OPT_OVERLAPPED_STARTS, OPT_PIPELINED_STARTS, OPT_PUSH_CTAS) = range(1, 27)
REG_CTRL, REG_MAX_FEA = 0x00, 0x70

EXPORTS = (
    "sgrace_create", "sgrace_destroy", "sgrace_last_error", "sgrace_version", "sgrace_alloc",
    "sgrace_free", "sgrace_sync_to_device", "sgrace_sync_from_device", "sgrace_write_reg",
    "sgrace_read_reg", "sgrace_write_reg64", "sgrace_reg_offset", "sgrace_set_option",
    "sgrace_get_option", "sgrace_set_stream", "sgrace_start", "sgrace_done", "sgrace_wait",
    "sgrace_stage_times", "sgrace_layer_run", "sgrace_fea_run", "sgrace_adj_run",
    "sgrace_launch_count", "sgrace_dense_run", "sgrace_peer_alloc", "sgrace_peer_open", "sgrace_peer_close", "sgrace_peer_release", "sgrace_prune_adjacency",
    "sgrace_adj_run_peer", "sgrace_halo_gather", "sgrace_halo_push", "sgrace_xty_run",
    "sgrace_peer_copy", "sgrace_peer_signal", "sgrace_wait_flag", "sgrace_sym_norm", "sgrace_dense_to_csr",
)


class LayerDesc(C.Structure):
    """sgrace_layer_desc"""
    _fields_ = [
        ("gemm_mode", C.c_int32), ("relu", C.c_int32), ("gat_mode", C.c_int32),
        ("N_adj", C.c_int32), ("M_adj", C.c_int32), ("M_fea", C.c_int32), ("P_w", C.c_int32),
        ("nnz_fea", C.c_int32), ("nnz_adj", C.c_int32),
        ("scale_fea", C.c_int32), ("internal_quantization", C.c_int32),
        ("qscale_fea", C.c_float), ("qscale_w", C.c_float), ("qscale_adj", C.c_float),
        ("deq_factor", C.c_float),
        ("rowPtr_fea", C.c_void_p), ("columnIndex_fea", C.c_void_p), ("values_fea", C.c_void_p),
        ("rowPtr_adj", C.c_void_p), ("columnIndex_adj", C.c_void_p), ("values_adj", C.c_void_p),
        ("B", C.c_void_p), ("attention", C.c_void_p), ("D", C.c_void_p),
        ("E", C.c_void_p), ("S", C.c_void_p), ("XW", C.c_void_p),
    ]


class SgraceError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsgrace_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libsgrace_b200.so; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the sm_100a CUDA library is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or sgracex1_b200/csrc/build.sh). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.sgrace_create.argtypes = [C.c_int, C.POINTER(H)]
    lib.sgrace_destroy.argtypes = [H]
    lib.sgrace_last_error.argtypes = [H]
    lib.sgrace_last_error.restype = C.c_char_p
    lib.sgrace_version.restype = C.c_char_p
    lib.sgrace_alloc.argtypes = [H, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.sgrace_free.argtypes = [H, C.c_uint64]
    lib.sgrace_sync_to_device.argtypes = [H, C.c_uint64, C.c_size_t]
    lib.sgrace_sync_from_device.argtypes = [H, C.c_uint64, C.c_size_t]
    lib.sgrace_write_reg.argtypes = [H, C.c_uint32, C.c_uint32]
    lib.sgrace_read_reg.argtypes = [H, C.c_uint32, C.POINTER(C.c_uint32)]
    lib.sgrace_write_reg64.argtypes = [H, C.c_uint32, C.c_uint64]
    lib.sgrace_reg_offset.argtypes = [C.c_char_p, C.POINTER(C.c_uint32)]
    lib.sgrace_set_option.argtypes = [H, C.c_int, C.c_int64]
    lib.sgrace_get_option.argtypes = [H, C.c_int, C.POINTER(C.c_int64)]
    lib.sgrace_set_stream.argtypes = [H, C.c_void_p]
    lib.sgrace_start.argtypes = [H]
    lib.sgrace_done.argtypes = [H, C.POINTER(C.c_int)]
    lib.sgrace_wait.argtypes = [H]
    lib.sgrace_stage_times.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.sgrace_layer_run.argtypes = [H, C.POINTER(LayerDesc)]
    lib.sgrace_fea_run.argtypes = [H, C.POINTER(LayerDesc), C.c_void_p]
    lib.sgrace_adj_run.argtypes = [H, C.POINTER(LayerDesc), C.c_void_p, C.c_int32]
    lib.sgrace_launch_count.argtypes = [H, C.POINTER(C.c_uint64)]
    lib.sgrace_peer_alloc.argtypes = [H, C.c_size_t, C.POINTER(C.c_uint64), C.c_char_p]
    lib.sgrace_peer_open.argtypes = [H, C.c_char_p, C.POINTER(C.c_uint64)]
    lib.sgrace_peer_release.argtypes = [H]
    lib.sgrace_peer_close.argtypes = [H]
    lib.sgrace_adj_run_peer.argtypes = [H, C.POINTER(LayerDesc), C.POINTER(C.c_uint64), C.c_int32, C.c_int32]
    lib.sgrace_halo_gather.argtypes = [H, C.POINTER(C.c_uint64), C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    lib.sgrace_halo_push.argtypes = [H, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]
    lib.sgrace_peer_copy.argtypes = [H, C.c_uint64, C.c_uint64, C.c_size_t]
    lib.sgrace_peer_signal.argtypes = [H, C.c_uint64, C.c_uint32]
    lib.sgrace_wait_flag.argtypes = [H, C.c_uint64, C.c_uint32]
    lib.sgrace_sym_norm.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    lib.sgrace_dense_to_csr.argtypes = [H, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_int64)]
    lib.sgrace_xty_run.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.sgrace_prune_adjacency.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    lib.sgrace_dense_run.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    for name in EXPORTS:
        if name not in ("sgrace_last_error", "sgrace_version"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def reg_offset(name: str):
    """AXI-Lite offset of a register name, or None when the map does not know it."""
    off = C.c_uint32()
    rc = load().sgrace_reg_offset(name.encode(), C.byref(off))
    return int(off.value) if rc == 0 else None


class Handle:
    """One accelerator instance (one CUDA stream), the stand-in for `ol.mmult_top_0`."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.h = C.c_void_p()
        rc = self.lib.sgrace_create(int(device), C.byref(self.h))
        if rc != 0:
            raise SgraceError(rc, f"sgrace_create(device={device}) failed -- is a CUDA device visible?")
        self.device = int(device)

    def _ck(self, rc):
        if rc != 0:
            raise SgraceError(rc, self.lib.sgrace_last_error(self.h).decode(errors="replace"))

    def close(self):
        if self.h:
            self.lib.sgrace_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # options ---------------------------------------------------------------
    def set_option(self, key, value):
        self._ck(self.lib.sgrace_set_option(self.h, int(key), int(value)))

    def get_option(self, key):
        v = C.c_int64()
        self._ck(self.lib.sgrace_get_option(self.h, int(key), C.byref(v)))
        return int(v.value)

    def set_stream(self, cuda_stream_ptr):
        """Bind the handle to a CUDA stream (0 / None: the handle's own).  Re-binding synchronises the
        old stream, so repeated calls with the same stream are skipped."""
        if getattr(self, "_stream", -1) == (cuda_stream_ptr or 0):
            return
        self._ck(self.lib.sgrace_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))
        self._stream = cuda_stream_ptr or 0

    # buffers ---------------------------------------------------------------
    def alloc(self, nbytes):
        host, dev = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.sgrace_alloc(self.h, int(nbytes), C.byref(host), C.byref(dev)))
        return int(host.value), int(dev.value)

    def free(self, dev_addr):
        self._ck(self.lib.sgrace_free(self.h, int(dev_addr)))

    def sync_to_device(self, dev_addr, nbytes):
        self._ck(self.lib.sgrace_sync_to_device(self.h, int(dev_addr), int(nbytes)))

    def sync_from_device(self, dev_addr, nbytes):
        self._ck(self.lib.sgrace_sync_from_device(self.h, int(dev_addr), int(nbytes)))

    # registers -------------------------------------------------------------
    def write_reg(self, off, value):
        self._ck(self.lib.sgrace_write_reg(self.h, int(off), int(value) & 0xFFFFFFFF))

    def write_reg64(self, off, value):
        self._ck(self.lib.sgrace_write_reg64(self.h, int(off), int(value) & 0xFFFFFFFFFFFFFFFF))

    def read_reg(self, off):
        v = C.c_uint32()
        self._ck(self.lib.sgrace_read_reg(self.h, int(off), C.byref(v)))
        return int(v.value)

    # run -------------------------------------------------------------------
    def start(self):
        self._ck(self.lib.sgrace_start(self.h))

    def done(self):
        d = C.c_int()
        self._ck(self.lib.sgrace_done(self.h, C.byref(d)))
        return bool(d.value)

    def wait(self):
        self._ck(self.lib.sgrace_wait(self.h))

    def stage_times(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._ck(self.lib.sgrace_stage_times(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return float(a.value), float(b.value), float(c.value)

    def layer_run(self, desc: LayerDesc):
        self._ck(self.lib.sgrace_layer_run(self.h, C.byref(desc)))

    def fea_run(self, desc: LayerDesc, xw_out_ptr):
        self._ck(self.lib.sgrace_fea_run(self.h, C.byref(desc), C.c_void_p(xw_out_ptr)))

    def adj_run(self, desc: LayerDesc, xw_in_ptr, xw_rows):
        self._ck(self.lib.sgrace_adj_run(self.h, C.byref(desc), C.c_void_p(xw_in_ptr), int(xw_rows)))

    def dense_run(self, x_ptr, b_ptr, out_ptr, N, M, P, relu=0):
        self._ck(self.lib.sgrace_dense_run(self.h, C.c_void_p(x_ptr), C.c_void_p(b_ptr), C.c_void_p(out_ptr), int(N), int(M),
                                           int(P), int(relu)))

    # peer memory (multi-GPU gathers over NVLink) --------------------------------
    def peer_alloc(self, nbytes):
        """Device buffer other ranks can map: returns (device address, 64-byte IPC handle)."""
        addr = C.c_uint64()
        buf = C.create_string_buffer(64)
        self._ck(self.lib.sgrace_peer_alloc(self.h, int(nbytes), C.byref(addr), buf))
        return int(addr.value), buf.raw

    def peer_open(self, handle_bytes):
        addr = C.c_uint64()
        self._ck(self.lib.sgrace_peer_open(self.h, C.create_string_buffer(bytes(handle_bytes), 64), C.byref(addr)))
        return int(addr.value)

    def peer_close(self):
        self._ck(self.lib.sgrace_peer_close(self.h))

    def peer_release(self):
        self._ck(self.lib.sgrace_peer_release(self.h))

    def adj_run_peer(self, desc: LayerDesc, bases, block_rows):
        arr = (C.c_uint64 * len(bases))(*[int(b) for b in bases])
        self._ck(self.lib.sgrace_adj_run_peer(self.h, C.byref(desc), arr, len(bases), int(block_rows)))

    def halo_gather(self, bases, block_rows, rows_ptr, n_rows, width, dst_ptr):
        arr = (C.c_uint64 * len(bases))(*[int(b) for b in bases])
        self._ck(self.lib.sgrace_halo_gather(self.h, arr, len(bases), int(block_rows), C.c_void_p(rows_ptr), int(n_rows),
                                             int(width), C.c_void_p(dst_ptr)))

    def halo_push(self, local_ptr, width, rows_ptrs, counts, dst_ptrs):
        n = len(rows_ptrs)
        self._ck(self.lib.sgrace_halo_push(self.h, C.c_void_p(local_ptr), int(width), n, (C.c_uint64 * n)(*[int(p) for p in rows_ptrs]),
                                           (C.c_int64 * n)(*[int(c) for c in counts]), (C.c_uint64 * n)(*[int(p) for p in dst_ptrs])))

    def peer_copy(self, dst_addr, src_addr, nbytes):
        self._ck(self.lib.sgrace_peer_copy(self.h, int(dst_addr), int(src_addr), int(nbytes)))

    def peer_signal(self, flag_addr, value):
        self._ck(self.lib.sgrace_peer_signal(self.h, int(flag_addr), int(value) & 0xffffffff))

    def wait_flag(self, flag_addr, value):
        self._ck(self.lib.sgrace_wait_flag(self.h, int(flag_addr), int(value) & 0xffffffff))

    def sym_norm(self, row_ptr, col_ptr, weight_ptr, nnz, n_nodes, fill, capacity, out_row, out_col, out_val, out_rowptr=None):
        """sgrace_sym_norm on device pointers; returns the length of the result."""
        n = C.c_int64()
        self._ck(self.lib.sgrace_sym_norm(self.h, C.c_void_p(row_ptr), C.c_void_p(col_ptr), C.c_void_p(weight_ptr or None), int(nnz),
                                          int(n_nodes), float(fill), int(capacity), C.c_void_p(out_row), C.c_void_p(out_col),
                                          C.c_void_p(out_val), C.c_void_p(out_rowptr or None), C.byref(n)))
        return int(n.value)

    def dense_to_csr(self, x_ptr, n, m, capacity, rowptr, col, val):
        """sgrace_dense_to_csr; returns (status, nnz): status is SGRACE_EBOUNDS when nnz exceeds the capacity."""
        k = C.c_int64()
        rc = self.lib.sgrace_dense_to_csr(self.h, C.c_void_p(x_ptr), int(n), int(m), int(capacity), C.c_void_p(rowptr),
                                          C.c_void_p(col or None), C.c_void_p(val or None), C.byref(k))
        if rc not in (0, -5):
            self._ck(rc)
        return rc, int(k.value)

    def prune_adjacency(self, rowptr, col, val, n, nnz, qscale_adj, qbits, out_rowptr, out_col, out_val, kept=0):
        """sgrace_prune_adjacency (device pointers); returns the number of surviving non-zeros."""
        k = C.c_int64()
        self._ck(self.lib.sgrace_prune_adjacency(self.h, C.c_void_p(rowptr), C.c_void_p(col or None), C.c_void_p(val or None), int(n),
                                                 int(nnz), C.c_float(qscale_adj), int(qbits), C.c_void_p(out_rowptr),
                                                 C.c_void_p(out_col or None), C.c_void_p(out_val or None), C.c_void_p(kept or None),
                                                 C.byref(k)))
        return int(k.value)

    def xty_run(self, x_ptr, y_ptr, out_ptr, N, M, P):
        self._ck(self.lib.sgrace_xty_run(self.h, C.c_void_p(x_ptr), C.c_void_p(y_ptr), C.c_void_p(out_ptr), int(N), int(M), int(P)))

    def launch_count(self):
        v = C.c_uint64()
        self._ck(self.lib.sgrace_launch_count(self.h, C.byref(v)))
        return int(v.value)
