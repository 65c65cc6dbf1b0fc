"""Register-level drivers for one layer, the way the reference's notebooks drive the IP.

`HostLayer` follows jupyter/test/mmult-master.ipynb cells 16-36: allocate buffers, fill them from
numpy, program the pointer / size registers, pulse AP_START, poll AP_DONE, read D back.  Host
buffers in, host buffers out; the library stages them over PCIe around the kernels.

`DeviceLayer` keeps every buffer resident in HBM (torch CUDA tensors) and calls the C ABI's
direct entry point (`sgrace_layer_run`, the argument list of `mmult_top`,
kernelMatrixmult_all.cpp:3762-3774) -- what a long-running job does between epochs.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .pynq_compat import MmultTop, allocate

_STORAGE = {
    _lib.MODE_F32_FAST: np.float32, _lib.MODE_F32_CSIM: np.float32, _lib.MODE_F16_CSIM: np.uint16,
    _lib.MODE_FIX16_CSIM: np.int16, _lib.MODE_FULL: np.float32,
}


def storage_dtype(mode):
    return _STORAGE[mode]


class HostLayer:
    """One accelerator + one set of buffers sized for `capacity` (N, M, P, nnz_adj, nnz_fea)."""

    def __init__(self, ip: MmultTop, mode=_lib.MODE_F32_FAST, *, N, M, P, nnz_adj, nnz_fea, dense=False,
                 gat=False, coo=False):
        self.ip, self.mode, self.coo, self.gat = ip, mode, coo, gat
        ip.configure(mode=mode, index_format=1 if coo else 0)
        vt = storage_dtype(mode)
        self.vt = vt
        al = lambda n, dt: allocate(max(int(n), 1), dtype=dt, target=ip)
        self.rowPtr_adj = al(nnz_adj if coo else N + 1, np.int32)
        self.columnIndex_adj = al(nnz_adj, np.int32)
        self.values_adj = al(nnz_adj, vt)
        self.rowPtr_fea = al((nnz_fea if coo else N + 1) if not dense else 1, np.int32)
        self.columnIndex_fea = al(nnz_fea if not dense else 1, np.int32)
        self.values_fea = al(N * M if dense else nnz_fea, vt)
        self.B = al(M * P, vt)
        self.D = al(N * P, vt)
        self.profiling = al(16, np.int64)
        if gat:
            self.E = al(nnz_adj, np.float32)
            self.S = al(nnz_adj, np.float32)
            self.attention = al(2 * P, np.float32)
        self._program_pointers()

    def _program_pointers(self):
        rm = self.ip.register_map
        for i in "1234":       # the notebooks point all four port replicas at the same buffer
            setattr(rm, f"rowPtr_fea{i}_offset_1", self.rowPtr_fea.physical_address)
            setattr(rm, f"columnIndex_fea{i}_offset_1", self.columnIndex_fea.physical_address)
            setattr(rm, f"values_fea{i}_offset_1", self.values_fea.physical_address)
            setattr(rm, f"rowPtr_adj{i}_offset_1", self.rowPtr_adj.physical_address)
            setattr(rm, f"columnIndex_adj{i}_offset_1", self.columnIndex_adj.physical_address)
            setattr(rm, f"values_adj{i}_offset_1", self.values_adj.physical_address)
            setattr(rm, f"D{i}_offset_1", self.D.physical_address)
        rm.B_offset_1 = self.B.physical_address
        rm.profiling_offset_1 = self.profiling.physical_address
        if self.gat:
            rm.E1_offset_1 = self.E.physical_address
            rm.S1_offset_1 = self.S.physical_address
            rm.ate_m_offset_1 = self.attention.physical_address

    def load(self, *, N, M, P, adj, B, fea=None, x_dense=None, relu=0, attention=None, gat_mode=0):
        """Fill the host buffers (values already in storage type) and the size registers."""
        rm = self.ip.register_map
        rp, ci, va = adj
        nnz_adj = len(ci)
        if self.coo:
            rows = np.repeat(np.arange(N, dtype=np.int32), np.diff(rp))
            self.rowPtr_adj[:nnz_adj] = rows
            rm.nnz_adj1 = nnz_adj
        else:
            self.rowPtr_adj[:N + 1] = rp
        self.columnIndex_adj[:nnz_adj] = ci
        self.values_adj[:nnz_adj] = va
        if x_dense is not None:
            self.values_fea[:N * M] = np.asarray(x_dense).reshape(-1)
            rm.gemm_mode = 1
        else:
            frp, fci, fva = fea
            nnz_fea = len(fci)
            if self.coo:
                self.rowPtr_fea[:nnz_fea] = np.repeat(np.arange(N, dtype=np.int32), np.diff(frp))
                rm.nnz_fea1 = nnz_fea
            else:
                self.rowPtr_fea[:N + 1] = frp
            self.columnIndex_fea[:nnz_fea] = fci
            self.values_fea[:nnz_fea] = fva
            rm.gemm_mode = 0
        self.B[:M * P] = B
        if attention is not None:
            self.attention[:2 * P] = np.asarray(attention, np.float32).reshape(-1)
        rm.relu = int(relu)
        rm.gat_mode = int(gat_mode)
        rm.N_adj, rm.M_adj, rm.M_fea, rm.P_w = N, N, M, P
        self.N, self.P, self.nnz_adj = N, P, nnz_adj

    def run(self):
        """AP_START, spin on AP_DONE (mmult-master.ipynb cell 32), return D as (N, P)."""
        rm = self.ip.register_map
        rm.CTRL.AP_START = 1
        while rm.CTRL.AP_DONE == 0:
            pass
        self.ip.handle.wait()
        return np.array(self.D[:self.N * self.P]).reshape(self.N, self.P)

    def free(self):
        for name in ("rowPtr_adj", "columnIndex_adj", "values_adj", "rowPtr_fea", "columnIndex_fea",
                     "values_fea", "B", "D", "profiling", "E", "S", "attention"):
            b = getattr(self, name, None)
            if b is not None:
                b.freebuffer()


class DeviceLayer:
    """All buffers resident in HBM as torch tensors; one `sgrace_layer_run` per call."""

    def __init__(self, handle: _lib.Handle, mode=_lib.MODE_F32_FAST, device="cuda:0"):
        import torch
        self.torch, self.h, self.mode, self.device = torch, handle, mode, device
        handle.set_option(_lib.OPT_MODE, mode)
        handle.set_option(_lib.OPT_INDEX_FORMAT, 0)
        self.t = {}
        self.desc = _lib.LayerDesc()

    def load(self, *, N, M, P, adj, B, fea=None, x_dense=None, relu=0, attention=None, gat_mode=0,
             consts=None, want_es=False):
        torch = self.torch
        dev = self.device
        vt = storage_dtype(self.mode)

        def up(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32,
                   np.dtype(np.uint16): torch.int16, np.dtype(np.int16): torch.int16}[np.dtype(dt)]
            src = torch.from_numpy(a.view(np.int16) if dt == np.uint16 else a)
            return src.to(dev, dtype=tdt)

        t = self.t
        t["rp_a"], t["ci_a"], t["va_a"] = up(adj[0], np.int32), up(adj[1], np.int32), up(adj[2], vt)
        d = self.desc
        if x_dense is not None:
            t["va_f"] = up(np.asarray(x_dense).reshape(-1), vt)
            d.gemm_mode = 1
            d.rowPtr_fea = d.columnIndex_fea = None
        else:
            t["rp_f"], t["ci_f"], t["va_f"] = up(fea[0], np.int32), up(fea[1], np.int32), up(fea[2], vt)
            d.gemm_mode = 0
            d.rowPtr_fea, d.columnIndex_fea = t["rp_f"].data_ptr(), t["ci_f"].data_ptr()
            d.nnz_fea = int(len(fea[1]))
        t["B"] = up(B, vt)
        tdt = torch.float32 if vt == np.float32 else torch.int16
        t["D"] = torch.zeros(N * P, dtype=tdt, device=dev)
        t["XW"] = torch.zeros(N * P, dtype=tdt, device=dev)
        d.relu, d.gat_mode = int(relu), int(gat_mode)
        d.N_adj, d.M_adj, d.M_fea, d.P_w = N, N, M, P
        d.nnz_adj = int(len(adj[1]))
        d.values_fea = t["va_f"].data_ptr()
        d.rowPtr_adj, d.columnIndex_adj, d.values_adj = t["rp_a"].data_ptr(), t["ci_a"].data_ptr(), t["va_a"].data_ptr()
        d.B, d.D, d.XW = t["B"].data_ptr(), t["D"].data_ptr(), t["XW"].data_ptr()
        if attention is not None:
            t["att"] = up(np.asarray(attention).reshape(-1), np.float32)
            d.attention = t["att"].data_ptr()
        if want_es or gat_mode:
            t["E"] = torch.zeros(max(1, d.nnz_adj), dtype=torch.float32, device=dev)
            t["S"] = torch.zeros(max(1, d.nnz_adj), dtype=torch.float32, device=dev)
            d.E, d.S = t["E"].data_ptr(), t["S"].data_ptr()
        if consts is not None:
            d.scale_fea = int(consts["scale_fea"])
            d.internal_quantization = int(consts["internal_quantization"])
            d.qscale_fea = float(np.float32(1.0 / consts["f_s"]))
            d.qscale_w = float(np.float32(1.0 / consts["w_s"]))
            d.qscale_adj = float(np.float32(1.0 / consts["a_s"]))
            d.deq_factor = float(np.float32(consts["deq_o"]))
        self.N, self.P = N, P
        torch.cuda.synchronize()

    def run(self, sync=True):
        self.h.layer_run(self.desc)
        if sync:
            self.h.wait()

    def result(self, name="D"):
        self.h.wait()
        a = self.t[name].cpu().numpy()
        if name in ("D", "XW"):
            a = a.reshape(self.N, self.P)
            if self.mode == _lib.MODE_F16_CSIM:
                a = a.view(np.uint16)
        return a
