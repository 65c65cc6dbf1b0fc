/*
 * sgrace_b200.h -- C ABI of libsgrace_b200.so, the B200 (sm_100a) replacement for the SGRACE
 * FPGA layer accelerator `mmult_top`.
 *
 * The reference reaches its accelerator through PYNQ: an AXI-Lite register file
 * (`my_ip.register_map.<name> = v`), physically contiguous DMA buffers (`allocate()` ->
 * `.physical_address`) and the AP_START / AP_DONE handshake.  This header is that boundary:
 *
 *   reference interface                                     entry point here
 *   ------------------------------------------------------  ---------------------------
 *   Overlay("gat_all_unsigned.bit"); ol.mmult_top_0         sgrace_create / sgrace_destroy
 *     (demo/sgrace_lib/sgrace.py:1274-1278,
 *      jupyter/molecule_gcn/Graph_Classification.ipynb cell 11:4-5)
 *   pynq.allocate(n, dtype) / buf.physical_address          sgrace_alloc
 *     (sgrace.py:1552-1642, notebook cell 11:7-20)
 *   buf.freebuffer()  (jupyter/test/mmult-master.ipynb 50)  sgrace_free
 *   register_map.<reg> = value                              sgrace_write_reg
 *     (sgrace.py:334-420, 1744-1891; kernelMatrixmult_all.cpp:3777-3861;
 *      offsets: demo/zcu104/gat_all_unsigned.hwh:16153-18563)
 *   int(register_map.max_fea)  (sgrace.py:506)              sgrace_read_reg
 *   register_map.CTRL.AP_START = 1  (sgrace.py:488)         sgrace_start
 *   register_map.CTRL.AP_DONE poll  (sgrace.py:489-491)     sgrace_done / sgrace_wait
 *   HLS build-time #defines (src/matrix_mult.h:80-195)      sgrace_set_option
 *   mmult_top(...) argument list                            sgrace_layer_run (device pointers,
 *     (kernelMatrixmult_all.cpp:3762-3774)                   no register file, no staging)
 *
 * All functions return 0 on success, a negative SGRACE_E* code otherwise, and never throw.
 * One handle owns one CUDA stream; calls on a handle must be serialised by the caller;
 * different handles may be used from different threads / ranks.
 */
#ifndef SGRACE_B200_H
#define SGRACE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgrace_handle sgrace_handle;

enum {
    SGRACE_OK = 0,
    SGRACE_EINVAL = -1,     /* bad argument / inconsistent registers               */
    SGRACE_ENOMEM = -2,
    SGRACE_ECUDA = -3,      /* CUDA error, text in sgrace_last_error               */
    SGRACE_EUNSUPPORTED = -4,
    SGRACE_EBOUNDS = -5     /* rowPtr not monotonic / column index out of range    */
};

/* ---- arithmetic modes (what matrix_mult.h:69-151 selects at build time) ---- */
enum {
    SGRACE_MODE_F32_FAST = 0,   /* float32, FMA, row order; within 1e-5 of FLOAT C-sim     */
    SGRACE_MODE_F32_CSIM = 1,   /* FLOAT build: bit-exact C-simulation accumulate order     */
    SGRACE_MODE_F16_CSIM = 2,   /* HALF build (matrix_mult.h:80): bit-exact, buffers fp16   */
    SGRACE_MODE_FIX16_CSIM = 3, /* EIGHTBIT build ap_fixed<16,2>: bit-exact, buffers int16  */
    SGRACE_MODE_FULL = 4        /* closed full design (quantise/GAT), float32 buffers,
                                   semantics of sgrace.py:563-681                          */
};

/* ---- options: the reference's build-time knobs as runtime fields ---- */
enum {
    SGRACE_OPT_MODE = 1,            /* SGRACE_MODE_*                                        */
    SGRACE_OPT_SPMM_BLOCK = 2,      /* SPMM_BLOCK        matrix_mult.h:169 (csim modes)     */
    SGRACE_OPT_LAT_FEA = 3,         /* FTYPE_LATENCY_FEA matrix_mult.h:118,138,150          */
    SGRACE_OPT_LAT_ADJ = 4,         /* FTYPE_LATENCY_ADJ                                    */
    SGRACE_OPT_FEA_THREADS = 5,     /* FEA_THREADS       matrix_mult.h:166                  */
    SGRACE_OPT_ADJ_THREADS = 6,     /* ADJ_THREADS       matrix_mult.h:167                  */
    SGRACE_OPT_USE_SBLOCKS = 7,     /* USE_SBLOCKS       matrix_mult.h:170                  */
    SGRACE_OPT_INDEX_FORMAT = 8,    /* 0: rowPtr buffers hold CSR pointers (open design)
                                       1: rowPtr buffers hold sorted COO row indices and
                                          nnz_fea1 / nnz_adj1 give the counts (full design,
                                          sgrace.py:1221-1249)                              */
    SGRACE_OPT_QBITS = 9,           /* config.w_qbits: 8,4,2,1; 0 = no quantisation (FULL)  */
    SGRACE_OPT_STAGING = 10,        /* 1: buffers from sgrace_alloc are copied host->device
                                          before and device->host after every start (PYNQ
                                          shared-memory behaviour); 0: device-resident      */
    SGRACE_OPT_LONG_ROW = 11,       /* rows with more non-zeros go to the CTA-per-row kernel */
    SGRACE_OPT_LEAKY_ALPHA_BITS = 12, /* float bits of the GAT LeakyReLU slope (default 0.2) */
    SGRACE_OPT_VALIDATE = 13,       /* 1: check CSR structure on the host mirror at start   */
    SGRACE_OPT_DENSE_TC = 14,       /* 1: allow the tcgen05 path for wide dense FEA (FAST)  */
    SGRACE_OPT_STREAM_KERNEL = 15,  /* 1 (default): TMA-staged persistent SpMM kernel (FAST);
                                       0: the row-strided kernel (any pointer alignment)    */
    SGRACE_OPT_ROW_OFFSET = 19,     /* row-partitioned GAT (sgrace_adj_run on a slice of the adjacency rows against the
                                       full feature-stage result): global index of the slice's first row; default 0 */
    SGRACE_OPT_FUSED_SMALL = 18,    /* sparse-feature FAST layers with at most this many rows run as ONE cooperative
                                       launch (W transpose | FEA | ADJ with grid barriers); default 65536, 0 = off */
    SGRACE_OPT_ACCUMULATE = 17,     /* 1: the ADJ stage computes D = act(D + A.XW): the second pass over an
                                       adjacency split by column ownership (multi-GPU halo)  */
    SGRACE_OPT_ADJ_PLAN = 20,       /* FAST-mode ADJ on block-diagonal adjacencies (batched graphs) with the XW window of a
                                       panel of graphs in shared memory (csrc/sgrace_spmm_panel.cuh).  0 (default): off --
                                       measured slower than the gather kernel when a graph's window leaves little room for
                                       the CSR rings (Cora-size blocks: 0.23 ms against 0.17 ms);
                                       1: the panel plan of an adjacency is found once per (rowPtr, columnIndex, sizes)
                                       and reused; 2: re-analysed at every launch.  A stale plan only costs speed:
                                       columns outside a panel's window are gathered from global memory, results are
                                       bit-equal to the gather kernel                                                   */
    SGRACE_OPT_OVERLAP = 23,        /* staging mode, float32 layers whose D is at least 16 MB: 1 (default) the adjacency goes up in
                                       row panels on a second stream, the aggregation runs panel by panel and each panel of D
                                       goes down on a third stream under the next upload; 0: all copies in, kernels, copies out */
    SGRACE_OPT_PUSH_CTAS = 26,      /* CTAs of sgrace_halo_push (0 = default, four per SM).  One per SM when the push runs beside
                                       the aggregation of the owned columns: it then leaves the SMs to that kernel              */
    SGRACE_OPT_PIPELINED_STARTS = 25,  /* read-only: of those, starts pipelined in row chunks (banded / block-diagonal adjacency) */
    SGRACE_OPT_OVERLAPPED_STARTS = 24, /* read-only: starts that took the overlapped path                                   */
    SGRACE_OPT_PANEL_LAUNCHES = 21, /* read-only: launches of the panel kernel so far                                  */
    SGRACE_OPT_PLAN_BUILDS = 22,    /* read-only: panel plans analysed so far                                          */
    SGRACE_OPT_AGG_FIRST = 16       /* 1: dense layers with M_fea < P_w run as act((A.X).W) --
                                       equal up to float rounding, gathers narrower rows;
                                       0 (default): the reference's order act(A.(X.W))      */
};

/* ---- register offsets: the AXI-Lite map of gat_all_unsigned.hwh:16153-18563 ---- */
enum {
    SGRACE_REG_CTRL = 0x00,         /* bit0 AP_START, bit1 AP_DONE, bit2 AP_IDLE, bit3 AP_READY */
    SGRACE_REG_LOAD_WEIGHTS = 0x10, SGRACE_REG_BETA_QU = 0x18, SGRACE_REG_F_ALIGN = 0x20,
    SGRACE_REG_QSCALE_ADJ = 0x28, SGRACE_REG_QSCALE_FEA = 0x30, SGRACE_REG_QSCALE_W = 0x38,
    SGRACE_REG_DEQ_FACTOR = 0x40, SGRACE_REG_STREAM_MODE = 0x48, SGRACE_REG_GAT_MODE = 0x50,
    SGRACE_REG_GEMM_MODE = 0x58, SGRACE_REG_RELU = 0x60, SGRACE_REG_SCALE_FEA = 0x68,
    SGRACE_REG_MAX_FEA = 0x70,      /* read-only */
    SGRACE_REG_LAYER_COUNT = 0x80, SGRACE_REG_QUANTIZED_MULTIPLIER = 0x88,
    SGRACE_REG_SHIFT = 0x90, SGRACE_REG_BIAS = 0x9c, SGRACE_REG_BIAS_COUNT = 0xa8,
    SGRACE_REG_PROFILING = 0xb0, SGRACE_REG_ZP_LHS = 0xbc, SGRACE_REG_ZP_RHS = 0xc4,
    SGRACE_REG_ZP_DST = 0xcc, SGRACE_REG_CLAMP_MAX = 0xd4, SGRACE_REG_CLAMP_MIN = 0xdc,
    SGRACE_REG_N_ADJ = 0xe4, SGRACE_REG_M_ADJ = 0xec, SGRACE_REG_M_FEA = 0xf4, SGRACE_REG_P_W = 0xfc,
    SGRACE_REG_B = 0x104,           /* pointer registers: <name>_offset_1 (low 32 bits) at the  */
    SGRACE_REG_D1 = 0x110, SGRACE_REG_D2 = 0x11c, SGRACE_REG_D3 = 0x128, SGRACE_REG_D4 = 0x134,
    SGRACE_REG_E1 = 0x140, SGRACE_REG_S1 = 0x14c, SGRACE_REG_ATE_M = 0x158, /* listed offset,    */
    SGRACE_REG_ARRAY_C_ADJUST = 0x164,                      /* <name>_offset_2 (high) at +4    */
    SGRACE_REG_NNZ_FEA1 = 0x16c, SGRACE_REG_NNZ_FEA2 = 0x174, SGRACE_REG_NNZ_FEA3 = 0x17c,
    SGRACE_REG_NNZ_FEA4 = 0x184,
    SGRACE_REG_ROWPTR_FEA1 = 0x18c, SGRACE_REG_ROWPTR_FEA2 = 0x198, SGRACE_REG_ROWPTR_FEA3 = 0x1a4,
    SGRACE_REG_ROWPTR_FEA4 = 0x1b0,
    SGRACE_REG_COLIDX_FEA1 = 0x1bc, SGRACE_REG_COLIDX_FEA2 = 0x1c8, SGRACE_REG_COLIDX_FEA3 = 0x1d4,
    SGRACE_REG_COLIDX_FEA4 = 0x1e0,
    SGRACE_REG_VALUES_FEA1 = 0x1ec, SGRACE_REG_VALUES_FEA2 = 0x1f8, SGRACE_REG_VALUES_FEA3 = 0x204,
    SGRACE_REG_VALUES_FEA4 = 0x210,
    SGRACE_REG_NNZ_ADJ1 = 0x21c, SGRACE_REG_NNZ_ADJ2 = 0x224, SGRACE_REG_NNZ_ADJ3 = 0x22c,
    SGRACE_REG_NNZ_ADJ4 = 0x234,
    SGRACE_REG_ROWPTR_ADJ1 = 0x23c, SGRACE_REG_ROWPTR_ADJ2 = 0x248, SGRACE_REG_ROWPTR_ADJ3 = 0x254,
    SGRACE_REG_ROWPTR_ADJ4 = 0x260,
    SGRACE_REG_COLIDX_ADJ1 = 0x26c, SGRACE_REG_COLIDX_ADJ2 = 0x278, SGRACE_REG_COLIDX_ADJ3 = 0x284,
    SGRACE_REG_COLIDX_ADJ4 = 0x290,
    SGRACE_REG_VALUES_ADJ1 = 0x29c, SGRACE_REG_VALUES_ADJ2 = 0x2a8, SGRACE_REG_VALUES_ADJ3 = 0x2b4,
    SGRACE_REG_VALUES_ADJ4 = 0x2c0,
    /* open-design-only pointer register (mmult-master.ipynb 31:34); accepted, unused
     * (the requantiser `scale()` is dead code, kernelMatrixmult_all.cpp:3485) */
    SGRACE_REG_QUANTIZED_MULTIPLIER_PTR = 0x400,
    SGRACE_REG_FILE_BYTES = 0x480
};

/* ---- lifecycle ---- */
int sgrace_create(int device, sgrace_handle** out);
int sgrace_destroy(sgrace_handle* h);
const char* sgrace_last_error(sgrace_handle* h);
const char* sgrace_version(void);

/* ---- buffers (pynq.allocate look-alike): pinned host mirror + device buffer ---- */
int sgrace_alloc(sgrace_handle* h, size_t bytes, void** host_ptr, uint64_t* device_addr);
int sgrace_free(sgrace_handle* h, uint64_t device_addr);
/* explicit copies between a mirror and its device buffer (offsets relative to device_addr) */
int sgrace_sync_to_device(sgrace_handle* h, uint64_t device_addr, size_t bytes);
int sgrace_sync_from_device(sgrace_handle* h, uint64_t device_addr, size_t bytes);

/* ---- register file ---- */
int sgrace_write_reg(sgrace_handle* h, uint32_t offset, uint32_t value);
int sgrace_read_reg(sgrace_handle* h, uint32_t offset, uint32_t* value);
int sgrace_write_reg64(sgrace_handle* h, uint32_t offset, uint64_t value); /* _offset_1 + _offset_2 */
int sgrace_reg_offset(const char* name, uint32_t* offset);                 /* "N_adj" -> 0xe4 ...   */

int sgrace_set_option(sgrace_handle* h, int key, int64_t value);
int sgrace_get_option(sgrace_handle* h, int key, int64_t* value);
int sgrace_set_stream(sgrace_handle* h, void* cuda_stream);   /* NULL: the handle's own stream;
                                                                 (void*)1 = cudaStreamLegacy (the default stream) */

/* ---- run: one layer per start, as one AP_START pulse does ---- */
int sgrace_start(sgrace_handle* h);
int sgrace_done(sgrace_handle* h, int* done);   /* non-blocking */
int sgrace_wait(sgrace_handle* h);              /* blocks; returns the layer's status */
/* per-stage device times of the last completed layer, milliseconds (after sgrace_wait) */
int sgrace_stage_times(sgrace_handle* h, float* fea_ms, float* adj_ms, float* total_ms);

/* ---- direct call with device pointers: the argument list of mmult_top ---- */
typedef struct {
    int32_t gemm_mode;          /* 0 sparse X, 1 dense X, 2 backward launch (float32 mode): D = values_adj[N_adj x M_adj, dense] . (CSR_fea[M_adj x M_fea] . W) */
    int32_t relu;
    int32_t gat_mode;           /* FULL mode only */
    int32_t N_adj, M_adj, M_fea, P_w;
    int32_t nnz_fea, nnz_adj;   /* required for INDEX_FORMAT 1; else may be 0 (= unknown)    */
    /* FULL-mode per-layer constants (register semantics, sgrace.py:334-365, 476) */
    int32_t scale_fea;
    int32_t internal_quantization;
    float   qscale_fea, qscale_w, qscale_adj;   /* 1/s as float32 */
    float   deq_factor;
    /* device pointers */
    const int32_t* rowPtr_fea; const int32_t* columnIndex_fea; const void* values_fea;
    const int32_t* rowPtr_adj; const int32_t* columnIndex_adj; const void* values_adj;
    const void* B;              /* W transposed, P_w x M_fea */
    const float* attention;     /* 2*P_w, FULL + gat_mode */
    void* D;                    /* N_adj x P_w */
    float* E; float* S;         /* nnz_adj each, FULL + gat_mode, may be NULL */
    void* XW;                   /* optional: caller-provided N_adj x P_w intermediate
                                   (the PIPO tile); NULL = library scratch */
} sgrace_layer_desc;

int sgrace_layer_run(sgrace_handle* h, const sgrace_layer_desc* d);        /* async on the stream */
/* stages separately, for multi-GPU row partitioning (all-gather of XW between them) */
int sgrace_fea_run(sgrace_handle* h, const sgrace_layer_desc* d, void* XW_out);
int sgrace_adj_run(sgrace_handle* h, const sgrace_layer_desc* d, const void* XW_in, int32_t xw_rows);

/* the dense FEA stage on its own with an optional ReLU: out = act(X . W), X dense N x M, B = W
 * transposed (P x M).  What loop_fea does in gemm_mode 1 (kernelMatrixmult_all.cpp:847-865); used by
 * multi-GPU callers of the aggregate-first order, where the activation follows the dense stage */
int sgrace_dense_run(sgrace_handle* h, const void* X, const void* B, void* out, int32_t N, int32_t M, int32_t P,
                     int32_t relu);

/* ---- multi-GPU: the ADJ stage gathering a ROW-PARTITIONED XW (or X) straight over NVLink ----
 * rank r of n_peers holds rows [r*block_rows, (r+1)*block_rows) of the gathered matrix in a buffer
 * from sgrace_peer_alloc; every rank opens the others' buffers with the 64-byte handles (exchanged
 * by the caller, e.g. torch.distributed.all_gather_object) and passes the table of n_peers device
 * addresses, its own included.  No all-gather: only the rows an adjacency row references move. */
int sgrace_peer_alloc(sgrace_handle* h, size_t bytes, uint64_t* device_addr, unsigned char handle_out[64]);
int sgrace_peer_open(sgrace_handle* h, const unsigned char handle_in[64], uint64_t* device_addr);
int sgrace_peer_close(sgrace_handle* h);       /* closes every mapping h imported (call on every rank, then barrier) */
int sgrace_peer_release(sgrace_handle* h);     /* closes the remaining mappings and frees every peer buffer h exported */
int sgrace_adj_run_peer(sgrace_handle* h, const sgrace_layer_desc* d, const uint64_t* bases, int32_t n_peers,
                        int32_t block_rows);
/* halo exchange: copy `n_rows` listed rows (global row ids, int32, device memory) of the partitioned
 * float32 matrix (`width` floats per row, multiple of 4) from their owners into `dst` (device) */
int sgrace_halo_gather(sgrace_handle* h, const uint64_t* bases, int32_t n_peers, int32_t block_rows, const int32_t* rows,
                       int64_t n_rows, int32_t width, void* dst);
/* the same exchange pushed by the owner: for each of n_dst destinations, rows_ptrs[d] (device int32
 * array of counts[d] LOCAL row indices) are copied from `local` to dst_ptrs[d] (peer-mapped address of
 * the destination's first halo slot for this owner), in order */
int sgrace_halo_push(sgrace_handle* h, const void* local, int32_t width, int32_t n_dst, const uint64_t* rows_ptrs,
                     const int64_t* counts, const uint64_t* dst_ptrs);

/* the exchange without any SM: copy-engine transfers between peer-visible buffers, ordered on the
 * handle's stream, and 32-bit flag words for the "my rows have landed" signal.
 *   sgrace_peer_copy    dst/src are device addresses on this GPU or peer-mapped ones
 *   sgrace_peer_signal  writes the 32-bit epoch `value` to the flag word after all earlier work on the stream
 *                       (cuStreamWriteValue32)
 *   sgrace_wait_flag    makes the stream wait (cuStreamWaitValue32, GEQ: wrap-safe) until the flag word, in a
 *                       buffer of this GPU, has reached `value` */
int sgrace_peer_copy(sgrace_handle* h, uint64_t dst, uint64_t src, size_t bytes);
int sgrace_peer_signal(sgrace_handle* h, uint64_t flag_addr, uint32_t value);
int sgrace_wait_flag(sgrace_handle* h, uint64_t flag_addr, uint32_t value);

/* saved-tensor backward helper of the notebook layer (FPYNQ.backward: grad_W = X^T (A g)):
 * out[M x P] = X^T . Y for X [N x M], Y [N x P] row-major float32, N >> M, P; deterministic */
int sgrace_xty_run(sgrace_handle* h, const void* X, const void* Y, void* out, int32_t N, int32_t M, int32_t P);

/* ---- graph preparation on the GPU (what the reference does on the host before every forward) ----
 * sgrace_sym_norm: sym_norm2 of demo/sgrace_lib/sgrace.py:18-51. row/col (int32, device, nnz entries) and
 * weight (float, device, or NULL = all ones) describe the edges; nodes without a self-loop get one of weight
 * `fill`, an existing self-loop keeps its weight; the result is sorted by (row, col):
 *   out_val[k] = deg^-1/2[row] * w * deg^-1/2[col],  deg = row sums of w in sorted order.
 * out_row / out_col / out_val (device) must hold `capacity` >= nnz + n_nodes entries; out_rowptr (device,
 * n_nodes + 1 ints, may be NULL) receives the CSR row pointer of the result; *out_nnz (host) its length.
 * Bit-equal to the torch code.  Synchronises the handle's stream (the length is returned to the host). */
int sgrace_sym_norm(sgrace_handle* h, const int32_t* row, const int32_t* col, const float* weight, int64_t nnz,
                    int32_t n_nodes, float fill, int64_t capacity, int32_t* out_row, int32_t* out_col, float* out_val,
                    int32_t* out_rowptr, int64_t* out_nnz);
/* sgrace_dense_to_csr: `to_sparse` of the feature matrix (sgrace.py:1218-1227, Graph_Classification.ipynb cell
 * 18:53-58): X [n x m] row-major float32 (device) -> rowptr (n + 1), col / val (capacity entries), entries
 * with X != 0 in row-major order. *out_nnz is always set; SGRACE_EBOUNDS if it exceeds `capacity` (col / val are
 * then not written). Synchronises the handle's stream. */
int sgrace_dense_to_csr(sgrace_handle* h, const float* X, int32_t n, int32_t m, int64_t capacity, int32_t* rowptr,
                        int32_t* col, float* val, int64_t* out_nnz);

/* sgrace_prune_adjacency: the adaptive pruning of the full design made explicit.  The quantiser maps small adjacency
 * entries to code 0 (quantization_ufbits, demo/sgrace_lib/sgrace.py:253-265, applied at :626-629; "adaptive
 * pruning", demo/emulation/demo_sgrace.py:32-36) and a zero code contributes nothing to the aggregation or to the
 * GAT softmax (mask at :640).  This call drops those entries once, on the device: CSR in (rowptr n + 1, col / val
 * nnz, device), CSR out (out_rowptr n + 1; out_col / out_val with room for nnz entries), survivors in their original
 * order with their ORIGINAL float values, so a layer run on the compacted arrays quantises them to the same non-zero
 * codes while streaming fewer bytes: the GCN aggregation returns the same D bit for bit, the GAT one the same
 * logits E and, because its softmax sums are grouped by edge position, the same S / D to rounding.  `kept` (optional, nnz ints) receives the
 * input index of every survivor (E / S of a GAT layer on the compacted graph map back through it).  qscale_adj =
 * 1 / a_s as written to quantization_scale_adj; the zero point is 0 as the driver programs it.  Synchronises the
 * handle's stream (*out_nnz is returned to the host). */
int sgrace_prune_adjacency(sgrace_handle* h, const int32_t* rowptr, const int32_t* col, const float* val, int32_t n, int64_t nnz,
                           float qscale_adj, int32_t qbits, int32_t* out_rowptr, int32_t* out_col, float* out_val, int32_t* kept,
                           int64_t* out_nnz);

/* number of this library's kernels launched on the handle since creation (bench evidence) */
int sgrace_launch_count(sgrace_handle* h, uint64_t* count);

#ifdef __cplusplus
}
#endif
#endif
