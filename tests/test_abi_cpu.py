"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/sgrace_b200.h
declares, knows the reference's register names at the reference's AXI-Lite offsets, and fails
loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes
import os
import re

import pytest

from sgracex1_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "sgrace_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sgrace_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sgrace_b200.h but not exported"
    assert set(_lib.EXPORTS) == set(syms)


def test_option_numbers_match_the_python_binding():
    """Every SGRACE_OPT_* of the header has the same number as _lib.OPT_* (both lists are written by hand), and the
    option keys `configure()` accepts exist."""
    hdr = open(os.path.join(ROOT, "include", "sgrace_b200.h")).read()
    opts = dict(re.findall(r"SGRACE_OPT_([A-Z0-9_]+)\s*=\s*(\d+)", hdr))
    assert len(opts) >= 26 and len(set(opts.values())) == len(opts)          # no number used twice
    for name, value in opts.items():
        assert getattr(_lib, "OPT_" + name) == int(value), name
    modes = dict(re.findall(r"SGRACE_MODE_([A-Z0-9_]+)\s*=\s*(\d+)", hdr))
    for name, value in modes.items():
        assert getattr(_lib, "MODE_" + name) == int(value), name


def test_register_offsets_follow_the_hardware_handoff_file():
    # spot values from demo/zcu104/gat_all_unsigned.hwh:16153-18563
    want = {"CTRL": 0x0, "gemm_mode": 0x58, "relu": 0x60, "gat_mode": 0x50, "scale_fea": 0x68, "max_fea": 0x70,
            "deq_factor": 0x40, "quantization_scale_adj": 0x28, "quantization_scale_fea": 0x30,
            "quantization_scale_w": 0x38, "quantized_multiplier": 0x88, "N_adj": 0xe4, "M_adj": 0xec,
            "M_fea": 0xf4, "P_w": 0xfc, "B_offset_1": 0x104, "B_offset_2": 0x108, "D1_offset_1": 0x110,
            "D4_offset_1": 0x134, "E1_offset_1": 0x140, "S1_offset_1": 0x14c, "ate_m_offset_1": 0x158,
            "nnz_fea1": 0x16c, "rowPtr_fea1_offset_1": 0x18c, "columnIndex_fea3_offset_1": 0x1d4,
            "values_fea4_offset_2": 0x214, "nnz_adj1": 0x21c, "rowPtr_adj1_offset_1": 0x23c,
            "columnIndex_adj2_offset_1": 0x278, "values_adj4_offset_1": 0x2c0, "profiling_offset_1": 0xb0,
            "bias_count": 0xa8, "load_weights": 0x10, "beta_qu": 0x18, "f_align": 0x20, "layer_count": 0x80}
    for name, off in want.items():
        assert _lib.reg_offset(name) == off, name
    # names the driver writes but the hand-off file does not list (sgrace.py:1855-1861)
    for name in ("E2_offset_1", "S4_offset_1", "nonexistent"):
        assert _lib.reg_offset(name) is None


def test_layer_desc_layout_matches_header():
    # 15 four-byte scalars, padding to 8, then 12 pointers
    assert ctypes.sizeof(_lib.LayerDesc) == 64 + 12 * 8
    assert _lib.LayerDesc.rowPtr_fea.offset == 64


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.SgraceError):
        _lib.Handle(0)
    from sgracex1_b200.pynq_compat import Overlay
    with pytest.raises(_lib.SgraceError):
        Overlay("gnn_all.bit")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sgracex1_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"libsgrace_oracle", r"oracle[/.]", r"_ref/"):
                    assert not re.search(pat, src, flags=re.M), f"{f} reaches into oracle/ ({pat})"
