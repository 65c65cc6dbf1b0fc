"""PYNQ look-alikes over libsgrace_b200.so, so reference driver code runs unmodified.

What the reference uses from PYNQ and what stands in for it here:

  Overlay("gat_all_unsigned.bit").mmult_top_0     -> Overlay(...).mmult_top_0  (a `MmultTop`)
      demo/sgrace_lib/sgrace.py:1274-1278; Graph_Classification.ipynb cell 11:4-5
  my_ip.register_map.<name> = int                 -> RegisterMap.__setattr__  (sgrace_write_reg)
      sgrace.py:334-420, 1744-1891; mmult-master.ipynb cell 31
  my_ip.register_map.CTRL.AP_START = 1 / .AP_DONE -> RegisterMap.CTRL         (sgrace_start / sgrace_done)
      sgrace.py:488-491
  int(my_ip.register_map.max_fea)                 -> RegisterMap.__getattr__  (sgrace_read_reg)
      sgrace.py:506
  allocate(shape, dtype) / .physical_address      -> allocate()               (sgrace_alloc)
      sgrace.py:1552-1642; notebook cell 11:7-20
  buf.freebuffer()                                -> PynqBuffer.freebuffer    (sgrace_free)

Differences a maintainer should know about: `physical_address` is a 64-bit CUDA device
address, so writing it to `<name>_offset_1` also fills `<name>_offset_2`; register names the
hardware hand-off file does not list (E2..E4, S2..S4 -- sgrace.py:1855-1861) are accepted and
ignored, as on the board.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_default_ip = None


def default_ip(device: int = 0):
    """The accelerator instance `allocate()` binds buffers to (the most recent Overlay)."""
    global _default_ip
    if _default_ip is None:
        _default_ip = MmultTop(device)
    return _default_ip


class _Ctrl:
    """CTRL register fields (gat_all_unsigned.hwh:16156-16240)."""

    def __init__(self, ip):
        object.__setattr__(self, "_ip", ip)

    def __setattr__(self, name, value):
        if name == "AP_START":
            if int(value) & 1:
                self._ip.started = True
                self._ip.handle.start()
        elif name in ("AUTO_RESTART", "INTERRUPT"):
            pass
        else:
            raise AttributeError(f"CTRL.{name} is read-only")

    @property
    def AP_DONE(self):
        return 1 if self._ip.handle.done() else 0

    @property
    def AP_READY(self):
        return self.AP_DONE

    @property
    def AP_IDLE(self):
        return 0 if (self._ip.started and not self._ip.handle.done()) else 1

    def __repr__(self):
        return f"Register(AP_START=0, AP_DONE={self.AP_DONE}, AP_IDLE={self.AP_IDLE}, AP_READY={self.AP_READY})"


class RegisterMap:
    def __init__(self, ip):
        object.__setattr__(self, "_ip", ip)
        object.__setattr__(self, "_ctrl", _Ctrl(ip))
        object.__setattr__(self, "_ignored", {})
        object.__setattr__(self, "_offsets", {})

    def _offset(self, name):
        offs = self._offsets
        if name not in offs:
            offs[name] = _lib.reg_offset(name)
        return offs[name]

    @property
    def CTRL(self):
        return self._ctrl

    def __setattr__(self, name, value):
        off = self._offset(name)
        if off is None:
            # not in the hardware hand-off register list: accepted and ignored, like a write to an
            # unmapped AXI-Lite address
            self._ignored[name] = int(value)
            return
        value = int(value)
        if name.endswith("_offset_1"):
            # device pointers are 64-bit: always write both halves, so that a small value written after a
            # 64-bit address does not keep the old upper word in *_offset_2
            self._ip.handle.write_reg64(off, value)
        else:
            self._ip.handle.write_reg(off, value)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        off = self._offset(name)
        if off is None:
            if name in self._ignored:
                return self._ignored[name]
            raise AttributeError(f"no register named {name}")
        return self._ip.handle.read_reg(off)

    def __repr__(self):
        return "RegisterMap {" + ", ".join(sorted(k for k in self._offsets if self._offsets[k] is not None)) + "}"


class MmultTop:
    """Stand-in for the `mmult_top_0` IP of the overlay."""

    def __init__(self, device: int = 0):
        self.handle = _lib.Handle(device)
        self.register_map = RegisterMap(self)
        self.started = False

    # convenience for the options that were HLS #defines
    def configure(self, **kw):
        keys = dict(mode=_lib.OPT_MODE, spmm_block=_lib.OPT_SPMM_BLOCK, lat_fea=_lib.OPT_LAT_FEA,
                    lat_adj=_lib.OPT_LAT_ADJ, fea_threads=_lib.OPT_FEA_THREADS,
                    adj_threads=_lib.OPT_ADJ_THREADS, use_sblocks=_lib.OPT_USE_SBLOCKS,
                    index_format=_lib.OPT_INDEX_FORMAT, qbits=_lib.OPT_QBITS, staging=_lib.OPT_STAGING,
                    long_row=_lib.OPT_LONG_ROW, validate=_lib.OPT_VALIDATE, dense_tc=_lib.OPT_DENSE_TC,
                    stream_kernel=_lib.OPT_STREAM_KERNEL, agg_first=_lib.OPT_AGG_FIRST,
                    fused_small=_lib.OPT_FUSED_SMALL, row_offset=_lib.OPT_ROW_OFFSET, adj_plan=_lib.OPT_ADJ_PLAN, overlap=_lib.OPT_OVERLAP)
        for k, v in kw.items():
            if k == "leaky_alpha":
                bits = int(np.asarray(v, dtype=np.float32).view(np.uint32))
                self.handle.set_option(_lib.OPT_LEAKY_ALPHA_BITS, bits)
            else:
                self.handle.set_option(keys[k], v)
        return self

    def run_kernel(self):
        """AP_START then spin on AP_DONE (mmult-master.ipynb cell 32)."""
        self.register_map.CTRL.AP_START = 1
        self.handle.wait()


class Overlay:
    """`Overlay(bitfile)`; the bitfile name only selects defaults:
    gat_all_unsigned.bit -> full design (float32 buffers, COO row indices, quantise/GAT),
    gnn_all.bit / anything else -> open design (true CSR).  The arithmetic type of the open
    design is a build-time choice in the reference (matrix_mult.h:80); pick it with
    `ol.mmult_top_0.configure(mode=...)`; the notebooks' fp16 buffers need MODE_F16_CSIM."""

    def __init__(self, bitfile: str = "gnn_all.bit", device: int = 0, download: bool = True):
        global _default_ip
        self.bitfile_name = bitfile
        self.mmult_top_0 = MmultTop(device)
        if "gat_all" in str(bitfile):
            self.mmult_top_0.configure(mode=_lib.MODE_FULL, index_format=1)
        _default_ip = self.mmult_top_0
        self.ip_dict = {"mmult_top_0": {"type": "xilinx.com:hls:mmult_top:1.0", "phys_addr": 0}}


class PynqBuffer(np.ndarray):
    """numpy array over the pinned host mirror of a device buffer."""

    def __new__(cls, shape, dtype, ip):
        dtype = np.dtype(dtype)
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        shape = tuple(int(s) for s in shape)
        nbytes = max(1, int(np.prod(shape)) * dtype.itemsize)
        host, dev = ip.handle.alloc(nbytes)
        raw = (C.c_char * nbytes).from_address(host)
        obj = np.ndarray.__new__(cls, shape, dtype, buffer=raw)
        obj._ip = ip
        obj._raw = raw
        obj.physical_address = dev
        obj.device_address = dev
        obj._freed = False
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        # views share the parent's device buffer, at the same byte offset
        base = getattr(obj, "physical_address", None)
        self._ip = getattr(obj, "_ip", None)
        self._raw = getattr(obj, "_raw", None)
        self._freed = True          # only the owner frees
        if base is not None and isinstance(obj, np.ndarray) and self.size and obj.size:
            try:
                delta = self.__array_interface__["data"][0] - obj.__array_interface__["data"][0]
                self.physical_address = base + delta
                self.device_address = base + delta
            except Exception:
                self.physical_address = base
                self.device_address = base

    def freebuffer(self):
        if not getattr(self, "_freed", True):
            self._ip.handle.free(self.physical_address)
            self._freed = True

    close = freebuffer

    def flush(self):          # host -> device
        self._ip.handle.sync_to_device(self.physical_address, self.nbytes)

    sync_to_device = flush

    def invalidate(self):     # device -> host
        self._ip.handle.sync_from_device(self.physical_address, self.nbytes)

    sync_from_device = invalidate


def allocate(shape, dtype=np.uint32, target=None, **_ignored):
    """`pynq.allocate` look-alike.  `target` may be an Overlay, its IP, or None (most recent)."""
    ip = target.mmult_top_0 if isinstance(target, Overlay) else (target or default_ip())
    return PynqBuffer(shape, dtype, ip)
