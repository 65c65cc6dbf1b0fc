"""sgracex1_b200 -- B200 (sm_100a) replacement for the SGRACE graph-layer accelerator.

Only the hot path lives here: the fused layer D = act(A . (X . W)) behind the reference's own
PYNQ register-map / buffer interface.

    csrc/            hand-written CUDA kernels + the C ABI (include/sgrace_b200.h)
    _lib.py          ctypes binding of libsgrace_b200.so (no fallback: raises if not built)
    pynq_compat.py   Overlay / allocate / register_map look-alikes
    config.py        the reference's flag module (demo/*/config.py)
    sgrace.py        host mirror of demo/sgrace_lib/sgrace.py (GATConv_SGRACE, init_SGRACE ...)
    molecule_gcn.py  host mirror of the molecule-GCN notebook layer (FPYNQ, GCN_PYNQ ...)
    graphs.py        seeded synthetic graphs of the benchmark shapes
    dist.py          row-partitioned / data-parallel multi-GPU drivers (torch.distributed)
"""
__version__ = "0.1.0"
