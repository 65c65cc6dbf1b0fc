// sgrace_abi.cu -- C ABI of libsgrace_b200.so (see include/sgrace_b200.h).
//
// Host-side replacement for the bottom half of the reference driver: the AXI-Lite register
// file + DMA buffers + AP_START/AP_DONE of `mmult_top` (kernelMatrixmult_all.cpp:3762-3967,
// demo/sgrace_lib/sgrace.py:321-559) become a handle that owns a CUDA stream, a register
// array, pinned-host/device buffer pairs and grow-only scratch, and dispatches the sm_100a
// kernels of sgrace_kernels.cuh.  There is no CPU compute path in this library.
#include "../../include/sgrace_b200.h"
#include "sgrace_kernels.cuh"
#include "sgrace_gemm_tc.cuh"
#include "sgrace_spmm_stream.cuh"
#include "sgrace_spmm_panel.cuh"
#include "sgrace_prep.cuh"

#include <cub/device/device_select.cuh>
#include <cuda/functional>
#include <thrust/iterator/counting_iterator.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

using namespace sgrace;

namespace {

struct Buffer {
    void* host = nullptr;
    void* dev = nullptr;
    size_t bytes = 0;
};

struct Scratch {
    void* p = nullptr;
    size_t bytes = 0;
};

// Panel plan of one adjacency (sgrace_spmm_panel.cuh): found once per (arrays, sizes, window capacity) and reused
struct AdjPlan {
    const int* rp = nullptr;
    const int* ci = nullptr;
    int nrows = 0, cap_fit = 0;
    long long nnz = 0;
    int npanels = 0, windowed_rows = 0;
    bool usable = false;
    int4* panels = nullptr;     // device
    int* info = nullptr;        // device: {npanels, windowed rows}
    uint64_t last_use = 0;
};

}  // namespace

struct sgrace_handle {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    uint32_t regs[SGRACE_REG_FILE_BYTES / 4];
    std::map<uint64_t, Buffer> buffers;   // keyed by device base address
    // options
    int mode = SGRACE_MODE_F32_FAST;
    int spmm_block = 1, lat_fea = 0, lat_adj = 0, fea_threads = 1, adj_threads = 1, use_sblocks = 0;
    int index_format = 0, qbits = 8, staging = 1, long_row = 512, validate = 0, dense_tc = 1, stream_kernel = 1, agg_first = 0, accumulate = 0;
    int overlap = 1;                            // SGRACE_OPT_OVERLAP
    int push_ctas = 0;                          // SGRACE_OPT_PUSH_CTAS
    cudaStream_t s_up = nullptr, s_down = nullptr;   // staging copies beside the kernels (created on first use)
    std::vector<cudaEvent_t> ev_pool;
    uint64_t overlapped_starts = 0, pipelined_starts = 0;
    void* pipe_host = nullptr;                  // pinned: panel starts / maxima of start_pipelined
    Scratch pipe_dev;
    int adj_plan = 0;                           // SGRACE_OPT_ADJ_PLAN (opt-in: measured slower than the gather kernel on Cora-size blocks)
    std::vector<AdjPlan> plans;                 // small cache of panel plans
    uint64_t plan_clock = 0, plan_builds = 0, panel_launches = 0;
    Scratch plan_a, plan_b, plan_c, plan_d, plan_flag, plan_starts, plan_tmp;
    int row_offset = 0;                         // global index of the ADJ stage's local row 0 (row-partitioned GAT)
    int fused_small = 65536;                    // layers with at most this many rows run as one cooperative launch (0: off)
    unsigned counter_phase = 0;                 // which of the two counter sets the next SpMM launch uses
    // set for the duration of sgrace_adj_run_peer
    const char* peer_base[MAX_PEERS] = {nullptr};
    int peer_block = 0, peer_count = 0;
    std::map<uint64_t, size_t> peer_allocs;     // device buffers from sgrace_peer_alloc
    std::vector<uint64_t> peer_opened;          // mappings from sgrace_peer_open
    float leaky_alpha = 0.2f;
    // scratch (grow-only)
    Scratch wrm, ax, long_partial, long_done, xw, wq, s1, s2, rp_fea, rp_adj, lists, counters;
        Scratch prep_keys, prep_ids, prep_misc, prep_tmp;   // graph preparation (sgrace_sym_norm / sgrace_dense_to_csr)
    int smem_optin = 0;      // cudaDevAttrMaxSharedMemoryPerBlockOptin
    // streaming-SpMM geometry overrides (0 = built-in default), read from SGRACE_STREAM_* once when the handle is
    // created; SGRACE_TUNE_LIVE=1 re-reads them at every launch (parameter sweeps from one process)
    struct StreamTune {
        int c_s = 0, c_g = 0, g_s = 0, g_g = 0, s_s = 0, s_g = 0, tr_s = 0, tr_g = 0, threads_s = 0, threads_g = 0;
        int ctas = 0, nosmem = 0, long_noseg = 0, live = 0;
        int p_c = 0, p_g = 0, p_s = 0, p_tr = 0, p_ncw = 0, p_long = 0, p_dbg = 0;
        int push_ctas = 0;                                         // halo push kernel: CTAs (SGRACE_HALO_PUSH_CTAS; 0 = 4 per SM)      // panel kernel (SGRACE_PANEL_*)
    } tune;
    std::map<std::pair<const void*, uint64_t>, int> launch_cfg;   // (kernel, threads << 32 | smem) -> resident CTAs per SM
    int* max_fea_dev = nullptr;
    // state
    bool running = false;
    int last_status = 0;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // kernels start, fea done, adj done, all done, start (incl. staging)
    bool ev_valid = false;
    uint64_t launches = 0;
    // pending device->host copies recorded at start (staging)
    char err[512];
};

namespace {

int fail(sgrace_handle* h, int code, const char* fmt, ...) {
    if (h) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->err, sizeof(h->err), fmt, ap);
        va_end(ap);
        h->last_status = code;
    }
    return code;
}

#define CU(call)                                                                           \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(h, SGRACE_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                               \
    } while (0)

int ensure(sgrace_handle* h, Scratch& s, size_t bytes) {
    if (bytes <= s.bytes) return 0;
    if (s.p) { CU(cudaStreamSynchronize(h->stream)); CU(cudaFree(s.p)); s.p = nullptr; s.bytes = 0; }
    size_t want = bytes + bytes / 4 + 256;
    CU(cudaMalloc(&s.p, want));
    s.bytes = want;
    return 0;
}

size_t elt_bytes(int mode) {
    return (mode == SGRACE_MODE_F16_CSIM || mode == SGRACE_MODE_FIX16_CSIM) ? 2 : 4;
}

int default_lat(int mode) {   // matrix_mult.h:117-118,137-138,149-150
    switch (mode) {
        case SGRACE_MODE_F32_CSIM: return 6;
        case SGRACE_MODE_F16_CSIM: return 4;
        default: return 1;
    }
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

void load_tune(sgrace_handle* h) {
    auto& t = h->tune;
    t.c_s = env_int("SGRACE_STREAM_C", 0); t.c_g = env_int("SGRACE_STREAM_C_G", 0);
    t.g_s = env_int("SGRACE_STREAM_G_S", 0); t.g_g = env_int("SGRACE_STREAM_G_G", 0);
    t.s_s = env_int("SGRACE_STREAM_S_S", 0); t.s_g = env_int("SGRACE_STREAM_S_G", 0);
    t.tr_s = env_int("SGRACE_STREAM_TR", 0); t.tr_g = env_int("SGRACE_STREAM_TR_G", 0);
    t.threads_s = env_int("SGRACE_STREAM_THREADS_S", 0); t.threads_g = env_int("SGRACE_STREAM_THREADS_G", 0);
    t.ctas = env_int("SGRACE_STREAM_CTAS", 0);
    t.nosmem = env_int("SGRACE_STREAM_NOSMEM", 0);
    t.long_noseg = env_int("SGRACE_LONG_NOSEG", 0);
    t.live = env_int("SGRACE_TUNE_LIVE", 0);
    t.p_c = env_int("SGRACE_PANEL_C", 0); t.p_g = env_int("SGRACE_PANEL_G", 0); t.p_s = env_int("SGRACE_PANEL_S", 0);
    t.p_tr = env_int("SGRACE_PANEL_TR", 0); t.p_ncw = env_int("SGRACE_PANEL_NCW", 0); t.p_long = env_int("SGRACE_PANEL_LONG", 0); t.p_dbg = env_int("SGRACE_PANEL_DBG", 0);
    t.push_ctas = env_int("SGRACE_HALO_PUSH_CTAS", 0);
}

// resident CTAs per SM of `kern` at this block size / dynamic shared memory; the attribute call and the occupancy
// query are made once per (kernel, geometry) and handle
template <typename K>
int launch_config(sgrace_handle* h, K kern, int threads, size_t smem, int* per_sm) {
    const auto key = std::make_pair((const void*)kern, ((uint64_t)threads << 32) | (uint64_t)smem);
    auto it = h->launch_cfg.find(key);
    if (it != h->launch_cfg.end()) { *per_sm = it->second; return 0; }
    // the attribute belongs to the kernel (per device, shared by every handle of the process), not to the geometry:
    // it is only ever raised
    {
        static std::mutex mu;
        static std::map<std::pair<int, const void*>, size_t> granted_by_kernel;
        std::lock_guard<std::mutex> lock(mu);
        size_t& granted = granted_by_kernel[std::make_pair(h->device, (const void*)kern)];
        if (smem > granted) {
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            granted = smem;
        }
    }
    int n = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem));
    h->launch_cfg[key] = n;
    *per_sm = n;
    return 0;
}

inline int grid_for(long long work_items, int block, int num_sms, int max_waves = 32) {
    long long g = (work_items + block - 1) / block;
    long long cap = (long long)num_sms * max_waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// lanes per row (LPR) x float4 chunks per lane (NV) for a row of P4 16-byte chunks
#define SGRACE_P4_DISPATCH(P4, CALL)                                                                         \
    do {                                                                                                     \
        if ((P4) == 1) { CALL(1, 1); } else if ((P4) == 2) { CALL(2, 1); } else if ((P4) <= 4) { CALL(4, 1); } \
        else if ((P4) <= 8) { CALL(8, 1); } else if ((P4) <= 16) { CALL(16, 1); } else if ((P4) <= 32) { CALL(32, 1); } \
        else if ((P4) <= 64) { CALL(32, 2); } else if ((P4) <= 128) { CALL(32, 4); } else { CALL(32, 8); }    \
    } while (0)

// Two sets of 16 int counters (long-row count, tile counter, ...) used by alternate launches: the
// long-row kernel that ends a launch zeroes the other set, so no memset sits between the kernels.
int counter_sets(sgrace_handle* h, int** cur, int** next) {
    if (h->counters.bytes < 256) {
        if (int rc = ensure(h, h->counters, 256)) return rc;
        CU(cudaMemsetAsync(h->counters.p, 0, h->counters.bytes, h->stream));
    }
    int* base = (int*)h->counters.p + 16;        // ints 0..15 stay with the GAT kernels
    *cur = base + 16 * (h->counter_phase & 1);
    *next = base + 16 * ((h->counter_phase + 1) & 1);
    h->counter_phase++;
    return 0;
}

// rows the main kernel deferred (longer than the threshold / a stage): segmented CTA-per-segment kernel
// with deterministic last-arriver combine; the row-per-CTA kernel covers lists it cannot hold
template <int NVL>
int launch_long_rows(sgrace_handle* h, const int* rp, const int* ci, const float* va, const float* Bm, float* out, int P4,
                     int relu, int* long_rows, int* long_count, long long nnz_hint, int* next_counters) {
    const size_t lsmem = sizeof(float4) * 8 * (size_t)P4;
    // segments <= nnz/SEG + #long rows, #long rows <= nnz/threshold
    const long long max_rows = nnz_hint > 0 ? nnz_hint / (h->long_row > 0 ? h->long_row : 1) + 1 : 0;
    const long long max_segs = nnz_hint > 0 ? nnz_hint / LONG_SEG + max_rows + 1 : 0;
    const size_t part_bytes = (size_t)max_segs * P4 * sizeof(float4);
    const bool seg = nnz_hint > 0 && part_bytes <= ((size_t)1 << 30) && !h->tune.long_noseg;
    PeerTable pt;
    memset(&pt, 0, sizeof(pt));
    pt.count = h->peer_count; pt.block = h->peer_block; pt.accumulate = h->accumulate;
    for (int r = 0; r < 8; r++) pt.base[r] = h->peer_base[r];
    if (seg) {
        if (int rc = ensure(h, h->long_partial, part_bytes + 16)) return rc;
        // per-row completion counters: zeroed once when (re)allocated, reset by the kernel after use
        const size_t done_bytes = sizeof(int) * (size_t)LONG_LIST_MAX;
        if (h->long_done.bytes < done_bytes) {
            if (int rc = ensure(h, h->long_done, done_bytes)) return rc;
            CU(cudaMemsetAsync(h->long_done.p, 0, h->long_done.bytes, h->stream));
        }
        if (lsmem > 12 * 1024)
            CU(cudaFuncSetAttribute(spmm_long_rows_seg_f32_kernel<NVL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
        spmm_long_rows_seg_f32_kernel<NVL><<<h->num_sms * 4, 256, lsmem, h->stream>>>(
            rp, ci, va, (const float4*)Bm, (float4*)out, P4, relu, long_rows, long_count, (float4*)h->long_partial.p,
            (int*)h->long_done.p, pt, next_counters);
    } else {
        if (lsmem > 48 * 1024)
            CU(cudaFuncSetAttribute(spmm_long_rows_f32_kernel<NVL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
        spmm_long_rows_f32_kernel<NVL><<<h->num_sms * 2, 256, lsmem, h->stream>>>(
            rp, ci, va, (const float4*)Bm, (float4*)out, P4, relu, long_rows, long_count, pt, next_counters);
    }
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}


// ------------------------------------------------------------------------------------
// fast float32 SpMM dispatch
// ------------------------------------------------------------------------------------
template <int LPR, int NV>
int launch_spmm_vec(sgrace_handle* h, const int* rp, const int* ci, const float* va, const float* Bm,
                    float* out, int nrows, int P, int relu) {
    const int P4 = P / 4;
    constexpr int RPW = 32 / LPR;
    // room for the long-row list: nrows ints + counter
    if (int rc = ensure(h, h->lists, sizeof(int) * (size_t)(nrows > 0 ? nrows : 1))) return rc;
    int *cset, *nset;
    if (int rc = counter_sets(h, &cset, &nset)) return rc;
    int* long_rows = (int*)h->lists.p;
    int* long_count = cset;
    const int block = 256;
    long long warps = ((long long)nrows + RPW - 1) / RPW;
    // persistent grid: exactly the CTAs that are resident at once (occupancy x SM count), each
    // striding over the rows, so there is a single wave and no tail of partial waves
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, spmm_csr_f32_kernel<LPR, NV>, block, 0));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    int grid = grid_for(warps * 32, block, h->num_sms, ctas_per_sm);
    spmm_csr_f32_kernel<LPR, NV><<<grid, block, 0, h->stream>>>(
        rp, ci, va, (const float4*)Bm, (float4*)out, nrows, P4, relu, h->long_row, long_rows, long_count, h->accumulate);
    h->launches++;
    CU(cudaGetLastError());
    constexpr int NVL = (LPR * NV + 31) / 32 > 0 ? (LPR * NV + 31) / 32 : 1;
    if (int rc = launch_long_rows<NVL>(h, rp, ci, va, Bm, out, P4, relu, long_rows, long_count, 0, nset)) return rc;
    return 0;
}

// ------------------------------------------------------------------------------------
// streaming (TMA-staged, warp-specialised) float32 SpMM dispatch
// ------------------------------------------------------------------------------------

// b_rows > 0: Bm has b_rows rows and is a candidate for shared-memory staging (FEA: W)
template <int LPR, int NV>
int launch_spmm_stream(sgrace_handle* h, const int* rp, const int* ci, const float* va, const float* Bm, float* out,
                       int nrows, int P, int relu, long long nnz_hint, int b_rows, int final_out, int b_total_rows,
                       const QConst* qadj = nullptr, int quant = 0) {
    const int P4 = P / 4;
    if (int rc = ensure(h, h->lists, sizeof(int) * (size_t)(nrows > 0 ? nrows : 1))) return rc;
    int *cset, *nset;
    if (int rc = counter_sets(h, &cset, &nset)) return rc;
    int* long_rows = (int*)h->lists.p;
    int* long_count = cset;
    int* tile_counter = cset + 2;

    StreamParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.rowptr = rp; sp.col = ci; sp.val = va; sp.out = (float4*)out;
    sp.nrows = nrows; sp.P4 = P4; sp.relu = relu;
    sp.streaming_store = final_out;
    sp.accumulate = h->accumulate;
    sp.long_rows = long_rows; sp.long_count = long_count; sp.tile_counter = tile_counter;
    sp.long_thresh = h->long_row;
    const bool exact = (P4 == LPR * NV);
    if (qadj) {
        if (!(exact && NV == 1) || h->peer_count > 0 || h->accumulate)
            return fail(h, SGRACE_EUNSUPPORTED, "quantised streaming ADJ needs P_w in {4,8,16,32,64,128}");
        sp.q.inv_as = qadj->inv_as; sp.q.a_z = qadj->a_z; sp.q.qbits = qadj->qbits; sp.q.den = qadj->den;
        sp.q.deq_o = qadj->deq_o; sp.q.quant = quant;
    }

    // where Bm rows are gathered from: shared memory when the whole matrix fits beside the stages
    if (h->tune.live) load_tune(h);
    const auto& tn = h->tune;
    const size_t budget = (size_t)h->smem_optin;
    const size_t b_plain = (size_t)b_rows * P * 4;
    int bsrc = BSRC_GLOBAL;
    // smem gathers: one CTA per SM split into 4 pipelines of 1 producer + 7 consumer warps with
    // 1024-non-zero stages; global gathers: 2-3 CTAs per SM of one pipeline each, 2048-non-zero stages
    int C = 2048, G = 1, S = 3;
    if (b_rows > 0 && !tn.nosmem && stream_smem_bytes(4, 2, 128, 1024, (int)b_plain) <= budget) {
        bsrc = BSRC_SMEM;
        C = 1024; G = 4; S = 3;
    }
    const bool glob = bsrc == BSRC_GLOBAL;
    if ((glob ? tn.c_g : tn.c_s) > 0) C = (glob ? tn.c_g : tn.c_s) & ~3;
    if ((glob ? tn.g_g : tn.g_s) > 0) G = glob ? tn.g_g : tn.g_s;
    // stage geometry.  smem gathers: a claimed tile holds ~2.5 stages of non-zeros and is cut into
    // 32 pieces, so a stage is filled to within one piece (~8%) and one claim feeds several stages.
    // global gathers: latency-bound, occupancy matters more than stage fill -> small row-pointer
    // slices (tile ~0.6 stage) so that three CTAs fit an SM.
    const double avg = (nnz_hint > 0 && nrows > 0) ? (double)nnz_hint / nrows : 8.0;
    int TR = (int)((glob ? 0.6 : 2.5) * C / (avg > 1.0 ? avg : 1.0));
    TR = (TR / 32) * 32;
    if (TR < 64) TR = 64;
    if (TR > (glob ? 512 : 1024)) TR = glob ? 512 : 1024;
    if ((glob ? tn.tr_g : tn.tr_s) > 0) TR = ((glob ? tn.tr_g : tn.tr_s) / 32) * 32;
    if (TR < 32) TR = 32;
    sp.tile_rows = TR; sp.stage_nnz = C; sp.groups = G;
    sp.b_bytes = glob ? 0 : (int)b_plain;
    sp.Bm = (const float4*)Bm;
    if ((glob ? tn.s_g : tn.s_s) > 0) S = glob ? tn.s_g : tn.s_s;
    while (S > 2 && stream_smem_bytes(G, S, TR, C, sp.b_bytes) > budget) S--;
    sp.stages = S;
    const size_t smem = stream_smem_bytes(G, S, TR, C, sp.b_bytes);
    if (smem > budget) return fail(h, SGRACE_EUNSUPPORTED, "streaming SpMM needs %zu bytes of shared memory", smem);

#define STREAM_LAUNCH(BS, MT, MB, EX, PEER, QA)                                                                \
    do {                                                                                                       \
        auto kern = spmm_stream_f32_kernel<LPR, NV, BS, MT, MB, EX, PEER, QA>;                                 \
        int threads = (BS == BSRC_GLOBAL && MT > 384 && MT < 1024) ? 384 : MT;                                 \
        if ((BS == BSRC_GLOBAL ? tn.threads_g : tn.threads_s) > 0) threads = BS == BSRC_GLOBAL ? tn.threads_g : tn.threads_s; \
        if (threads > MT) threads = MT;                                                                        \
        threads = (threads / (32 * G)) * (32 * G);                                                             \
        if (threads < 64 * G) threads = 64 * G;                                                                \
        int per_sm = 1;                                                                                        \
        if (int rc = launch_config(h, kern, threads, smem, &per_sm)) return rc;                                \
        if (per_sm < 1) return fail(h, SGRACE_ECUDA, "streaming SpMM does not fit an SM (%zu B smem)", smem);  \
        if (tn.ctas > 0 && per_sm > tn.ctas) per_sm = tn.ctas;                                                 \
        const long long tiles = ((long long)nrows + TR - 1) / TR;                                              \
        long long grid = (long long)h->num_sms * per_sm;                                                       \
        if (grid > tiles) grid = tiles;                                                                        \
        kern<<<(int)grid, threads, smem, h->stream>>>(sp);                                                     \
    } while (0)
    // register budgets: NV == 1 kernels fit 64 registers -> 1024-thread CTAs (smem gathers) or
    // several 512-thread CTAs per SM (global gathers, latency-bound: occupancy matters); wide rows
    // (NV > 1) get 512 x 1 -> 128 registers
    constexpr int MT_SMEM = NV == 1 ? 1024 : 512;
    constexpr int MB_GLOB = NV == 1 ? 2 : 1;
    if (bsrc == BSRC_SMEM) {
        if (exact) STREAM_LAUNCH(BSRC_SMEM, MT_SMEM, 1, true, false, QM_NONE); else STREAM_LAUNCH(BSRC_SMEM, 512, 1, false, false, QM_NONE);
    } else if (h->peer_count > 0) {
        // row-partitioned Bm gathered over NVLink; only wide rows (a full warp per row) are instantiated
        sp.peer_count = h->peer_count; sp.peer_block = h->peer_block;
        for (int r = 0; r < MAX_PEERS; r++) sp.peer_base[r] = h->peer_base[r];
        if (LPR != 32) return fail(h, SGRACE_EUNSUPPORTED, "peer gathers need P_w >= 68 (one warp per row)");
        if (LPR == 32) { if (exact) STREAM_LAUNCH(BSRC_GLOBAL, 512, MB_GLOB, true, true, QM_NONE); else STREAM_LAUNCH(BSRC_GLOBAL, 512, 1, false, true, QM_NONE); }
    } else if (qadj) {
        if (NV == 1) STREAM_LAUNCH(BSRC_GLOBAL, 512, 2, true, false, QM_GCN);
    } else {
        if (exact) STREAM_LAUNCH(BSRC_GLOBAL, 512, MB_GLOB, true, false, QM_NONE); else STREAM_LAUNCH(BSRC_GLOBAL, 512, 1, false, false, QM_NONE);
    }
#undef STREAM_LAUNCH
    h->launches++;
    CU(cudaGetLastError());

    if (qadj) {
        // rows the streaming kernel deferred, in the same multiply-then-add order
        adj_q_gcn_list_kernel<<<h->num_sms * 2, 256, 0, h->stream>>>(rp, ci, va, Bm, out, long_rows, long_count, P, relu, quant, *qadj, nset);
        h->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    constexpr int NVL = (LPR * NV + 31) / 32 > 0 ? (LPR * NV + 31) / 32 : 1;
    if (int rc = launch_long_rows<NVL>(h, rp, ci, va, Bm, out, P4, relu, long_rows, long_count, nnz_hint, nset)) return rc;
    return 0;
}


// ------------------------------------------------------------------------------------
// panel (shared-memory window) ADJ: plan + launch
// ------------------------------------------------------------------------------------
struct PanelGeom { int G, S, C, TR, ncw, threads, cap_fit, hub; size_t smem; int win_bytes; };

// ring geometry of the panel kernel for rows of `rowbytes`; the window gets what the rings leave
bool panel_geometry(const sgrace_handle* h, int rowbytes, double avg_nnz, PanelGeom* g) {
    const auto& tn = h->tune;
    if (avg_nnz < 1.0) avg_nnz = 1.0;
    g->G = tn.p_g > 0 ? tn.p_g : 2;
    g->S = tn.p_s > 0 ? tn.p_s : 3;
    g->C = tn.p_c > 0 ? (tn.p_c & ~3) : 1024;
    g->hub = tn.p_long > 0 ? tn.p_long : 32;
    int TR = tn.p_tr > 0 ? tn.p_tr : (int)(1.25 * g->C / avg_nnz);
    TR = (TR / 32) * 32;
    if (TR < 64) TR = 64;
    if (TR > 512) TR = 512;
    g->TR = TR;
    if (g->G < 1 || g->G > 8) return false;
    g->ncw = tn.p_ncw > 0 ? tn.p_ncw : (23 / g->G) - 1;          // 768 threads: 85 registers each, no spills
    if (g->ncw < 1) return false;
    g->threads = 32 * (g->G * (1 + g->ncw) + 1);
    if (g->threads > 768) return false;
    const size_t fixed = panel_smem_bytes(g->G, g->S, g->TR, g->C, 0);
    if (fixed + 16 * 1024 > (size_t)h->smem_optin) return false;
    g->win_bytes = (int)(((size_t)h->smem_optin - fixed) & ~(size_t)127);
    g->cap_fit = g->win_bytes / rowbytes;
    g->smem = panel_smem_bytes(g->G, g->S, g->TR, g->C, g->win_bytes);
    return g->cap_fit >= 64;
}

// finds (or builds) the plan of this adjacency; *out = nullptr when the panel kernel should not be used
int adj_plan_for(sgrace_handle* h, const int* rp, const int* ci, int nrows, long long nnz, int cap_fit, const AdjPlan** out) {
    *out = nullptr;
    h->plan_clock++;
    if (h->adj_plan != 2) {
        for (auto& pl : h->plans)
            if (pl.rp == rp && pl.ci == ci && pl.nrows == nrows && pl.nnz == nnz && pl.cap_fit == cap_fit) {
                pl.last_use = h->plan_clock;
                if (pl.usable) *out = &pl;
                return 0;
            }
    }
    // building needs a device->host read of two ints: not inside a stream capture
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CU(cudaStreamIsCapturing(h->stream, &cap));
    if (cap != cudaStreamCaptureStatusNone) return 0;
    AdjPlan* slot = nullptr;
    for (auto& pl : h->plans)
        if (pl.rp == rp && pl.ci == ci && pl.nrows == nrows && pl.nnz == nnz && pl.cap_fit == cap_fit) slot = &pl;
    if (!slot) {
        if (h->plans.size() < 8) { h->plans.emplace_back(); slot = &h->plans.back(); }
        else { slot = &h->plans[0]; for (auto& pl : h->plans) if (pl.last_use < slot->last_use) slot = &pl; }
    }
    CU(cudaStreamSynchronize(h->stream));           // a plan that is being replaced may still be in use
    if (slot->panels) { cudaFree(slot->panels); slot->panels = nullptr; }
    if (!slot->info) CU(cudaMalloc(&slot->info, 16));
    const size_t n = (size_t)nrows;
    if (int rc = ensure(h, h->plan_a, 4 * n)) return rc;
    if (int rc = ensure(h, h->plan_b, 4 * n)) return rc;
    if (int rc = ensure(h, h->plan_c, 4 * n)) return rc;
    if (int rc = ensure(h, h->plan_d, 4 * n)) return rc;
    if (int rc = ensure(h, h->plan_flag, n)) return rc;
    if (int rc = ensure(h, h->plan_starts, 4 * n + 16)) return rc;
    int *hi = (int*)h->plan_a.p, *lo_rev = (int*)h->plan_b.p, *pmax = (int*)h->plan_c.p, *smin = (int*)h->plan_d.p;
    unsigned char* flag = (unsigned char*)h->plan_flag.p;
    int* starts = (int*)h->plan_starts.p;
    int* nstarts = starts + n;
    const int blocks = (nrows + 255) / 256;
    plan::row_extent_kernel<<<blocks, 256, 0, h->stream>>>(rp, ci, nrows, hi, lo_rev);
    size_t t1 = 0, t2 = 0, t3 = 0;
    CU(cub::DeviceScan::InclusiveScan(nullptr, t1, hi, pmax, cuda::maximum<>{}, nrows, h->stream));
    CU(cub::DeviceScan::InclusiveScan(nullptr, t2, lo_rev, smin, cuda::minimum<>{}, nrows, h->stream));
    thrust::counting_iterator<int> counting(0);
    CU(cub::DeviceSelect::Flagged(nullptr, t3, counting, flag, starts, nstarts, nrows, h->stream));
    size_t tmp = t1 > t2 ? t1 : t2; if (t3 > tmp) tmp = t3;
    if (int rc = ensure(h, h->plan_tmp, tmp + 16)) return rc;
    CU(cub::DeviceScan::InclusiveScan(h->plan_tmp.p, t1, hi, pmax, cuda::maximum<>{}, nrows, h->stream));
    CU(cub::DeviceScan::InclusiveScan(h->plan_tmp.p, t2, lo_rev, smin, cuda::minimum<>{}, nrows, h->stream));
    plan::block_flag_kernel<<<blocks, 256, 0, h->stream>>>(pmax, smin, nrows, flag);
    CU(cub::DeviceSelect::Flagged(h->plan_tmp.p, t3, counting, flag, starts, nstarts, nrows, h->stream));
    // panels of a few per SM when the graph is small, never more rows than the window holds
    int cap_pack = (int)((long long)nrows / ((long long)h->num_sms * 4) + 1);
    if (cap_pack < 256) cap_pack = 256;
    if (cap_pack > cap_fit) cap_pack = cap_fit;
    int nstarts_h = 0;
    CU(cudaMemcpyAsync(&nstarts_h, nstarts, 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    const size_t max_panels = (size_t)nstarts_h + n / cap_pack + 2;
    CU(cudaMalloc(&slot->panels, sizeof(int4) * max_panels));
    plan::pack_kernel<<<1, 1024, 0, h->stream>>>(starts, nstarts, nrows, cap_fit, cap_pack, slot->panels, slot->info);
    h->launches += 6;
    CU(cudaGetLastError());
    int info_h[2] = {0, 0};
    CU(cudaMemcpyAsync(info_h, slot->info, 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    slot->rp = rp; slot->ci = ci; slot->nrows = nrows; slot->nnz = nnz; slot->cap_fit = cap_fit;
    slot->npanels = info_h[0]; slot->windowed_rows = info_h[1];
    // worth it when most rows get a window and there is a panel for every SM
    slot->usable = info_h[0] >= h->num_sms && (double)info_h[1] >= 0.9 * (double)nrows;
    slot->last_use = h->plan_clock;
    h->plan_builds++;
    if (slot->usable) *out = slot;
    return 0;
}

// returns -100 when the panel kernel does not apply (the caller takes the gather kernel)
template <int LPR, int NV>
int launch_spmm_panel(sgrace_handle* h, const int* rp, const int* ci, const float* va, const float* Bm, float* out,
                      int nrows, int P, int relu, long long nnz_hint, int final_out, int b_total_rows) {
    if (NV > 2) return -100;
    const int rowbytes = P * 4;
    PanelGeom g;
    if (h->tune.live) load_tune(h);
    if (!panel_geometry(h, rowbytes, nrows > 0 ? (double)nnz_hint / nrows : 8.0, &g)) return -100;
    const AdjPlan* pl = nullptr;
    if (int rc = adj_plan_for(h, rp, ci, nrows, nnz_hint, g.cap_fit, &pl)) return rc;
    if (!pl) return -100;
    if (int rc = ensure(h, h->lists, sizeof(int) * (size_t)(nrows > 0 ? nrows : 1))) return rc;
    int *cset, *nset;
    if (int rc = counter_sets(h, &cset, &nset)) return rc;
    PanelParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.rowptr = rp; pp.col = ci; pp.val = va; pp.Bm = (const float4*)Bm; pp.out = (float4*)out;
    pp.nrows = nrows; pp.relu = relu; pp.streaming_store = final_out; pp.bm_rows = b_total_rows;
    pp.long_thresh = h->long_row; pp.hub_thresh = g.hub < h->long_row ? g.hub : h->long_row;
    pp.tile_rows = g.TR; pp.stage_nnz = g.C; pp.stages = g.S; pp.groups = g.G;
    pp.win_bytes = g.win_bytes; pp.panels = pl->panels; pp.npanels = pl->info;
    pp.long_rows = (int*)h->lists.p; pp.long_count = cset; pp.panel_counter = cset + 3;
    pp.dbg = h->tune.p_dbg;
    auto kern = spmm_panel_f32_kernel<LPR, NV, 768>;
    int per_sm = 0;
    if (int rc = launch_config(h, kern, g.threads, g.smem, &per_sm)) return rc;
    if (per_sm < 1) return fail(h, SGRACE_ECUDA, "panel SpMM does not fit an SM (%zu B smem)", g.smem);
    int grid = h->num_sms < pl->npanels ? h->num_sms : pl->npanels;
    kern<<<grid, g.threads, g.smem, h->stream>>>(pp);
    h->launches++;
    h->panel_launches++;
    CU(cudaGetLastError());
    constexpr int NVL = (LPR * NV + 31) / 32 > 0 ? (LPR * NV + 31) / 32 : 1;
    return launch_long_rows<NVL>(h, rp, ci, va, Bm, out, P / 4, relu, pp.long_rows, pp.long_count, nnz_hint, nset);
}

template <int NC>
int launch_spmm_scalar(sgrace_handle* h, const int* rp, const int* ci, const float* va, const float* Bm,
                       float* out, int nrows, int P, int relu) {
    int grid = grid_for((long long)nrows * 32, 256, h->num_sms, 8);
    spmm_csr_f32_scalar_kernel<NC><<<grid, 256, 0, h->stream>>>(rp, ci, va, Bm, out, nrows, P, relu);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

int spmm_f32(sgrace_handle* h, const int* rp, const int* ci, const float* va, const float* Bm, float* out,
             int nrows, int P, int relu, long long nnz_hint = 0, int b_rows = 0, int final_out = 0,
             int b_total_rows = 0) {
    if (nrows <= 0 || P <= 0) return 0;
    const bool aligned = (P % 4 == 0) && (((uintptr_t)Bm & 15) == 0) && (((uintptr_t)out & 15) == 0);
    // the streaming kernel bulk-copies 16-byte groups of the CSR arrays
    const bool csr_aligned = ((((uintptr_t)rp) | ((uintptr_t)ci) | ((uintptr_t)va)) & 15) == 0;
    if (aligned && csr_aligned && h->stream_kernel && h->adj_plan && b_rows == 0 && b_total_rows == nrows && nnz_hint > 0 &&
        !h->accumulate && h->peer_count == 0 && nrows >= 4096) {
        // square adjacency: try the shared-memory window kernel (block-diagonal batches); -100 = does not apply
        const int P4 = P / 4;
        int rc = -100;
#define SGRACE_PANEL(L, V) rc = launch_spmm_panel<L, V>(h, rp, ci, va, Bm, out, nrows, P, relu, nnz_hint, final_out, b_total_rows)
        if (P4 == 1) SGRACE_PANEL(1, 1); else if (P4 == 2) SGRACE_PANEL(2, 1); else if (P4 == 4) SGRACE_PANEL(4, 1);
        else if (P4 == 8) SGRACE_PANEL(8, 1); else if (P4 == 16) SGRACE_PANEL(16, 1); else if (P4 == 32) SGRACE_PANEL(32, 1);
        else if (P4 == 64) SGRACE_PANEL(32, 2);
#undef SGRACE_PANEL
        if (rc != -100) return rc;
    }
    if (aligned && csr_aligned && h->stream_kernel) {
        const int P4 = P / 4;
#define SGRACE_STREAM(L, V) return launch_spmm_stream<L, V>(h, rp, ci, va, Bm, out, nrows, P, relu, nnz_hint, b_rows, final_out, b_total_rows)
        if (P4 == 1) SGRACE_STREAM(1, 1);
        if (P4 == 2) SGRACE_STREAM(2, 1);
        if (P4 <= 4) SGRACE_STREAM(4, 1);
        if (P4 <= 8) SGRACE_STREAM(8, 1);
        if (P4 <= 16) SGRACE_STREAM(16, 1);
        if (P4 <= 32) SGRACE_STREAM(32, 1);
        if (P4 <= 64) SGRACE_STREAM(32, 2);
        if (P4 <= 128) SGRACE_STREAM(32, 4);
        if (P4 <= 256) SGRACE_STREAM(32, 8);
#undef SGRACE_STREAM
        return fail(h, SGRACE_EUNSUPPORTED, "P_w=%d > 1024 not supported", P);
    }
    if ((h->peer_count > 0 || h->accumulate) && !aligned)
        return fail(h, SGRACE_EUNSUPPORTED, "peer gathers / accumulate need 16-byte aligned operands and P_w a multiple of 4");
    if (aligned) {
        const int P4 = P / 4;
        if (P4 == 1) return launch_spmm_vec<1, 1>(h, rp, ci, va, Bm, out, nrows, P, relu);
        if (P4 == 2) return launch_spmm_vec<2, 1>(h, rp, ci, va, Bm, out, nrows, P, relu);
        if (P4 <= 4) return launch_spmm_vec<4, 1>(h, rp, ci, va, Bm, out, nrows, P, relu);
        if (P4 <= 8) return launch_spmm_vec<8, 1>(h, rp, ci, va, Bm, out, nrows, P, relu);
        if (P4 <= 16) return launch_spmm_vec<16, 1>(h, rp, ci, va, Bm, out, nrows, P, relu);
        if (P4 <= 32) return launch_spmm_vec<32, 1>(h, rp, ci, va, Bm, out, nrows, P, relu);
        if (P4 <= 64) return launch_spmm_vec<32, 2>(h, rp, ci, va, Bm, out, nrows, P, relu);
        if (P4 <= 128) return launch_spmm_vec<32, 4>(h, rp, ci, va, Bm, out, nrows, P, relu);
        if (P4 <= 256) return launch_spmm_vec<32, 8>(h, rp, ci, va, Bm, out, nrows, P, relu);
        return fail(h, SGRACE_EUNSUPPORTED, "P_w=%d > 1024 not supported", P);
    }
    if (P <= 32) return launch_spmm_scalar<1>(h, rp, ci, va, Bm, out, nrows, P, relu);
    if (P <= 64) return launch_spmm_scalar<2>(h, rp, ci, va, Bm, out, nrows, P, relu);
    if (P <= 128) return launch_spmm_scalar<4>(h, rp, ci, va, Bm, out, nrows, P, relu);
    if (P <= 256) return launch_spmm_scalar<8>(h, rp, ci, va, Bm, out, nrows, P, relu);
    if (P <= 1024) return launch_spmm_scalar<32>(h, rp, ci, va, Bm, out, nrows, P, relu);
    return fail(h, SGRACE_EUNSUPPORTED, "P_w=%d > 1024 not supported", P);
}

// ------------------------------------------------------------------------------------
// bit-exact C-simulation-order stage dispatch
// ------------------------------------------------------------------------------------
template <typename Ops>
int stage_exact(sgrace_handle* h, int lat, const int* rp, const int* ci, const void* va, const void* Bm,
                void* out, int nrows, int P, int hw_threads, int sblock, int dense_M, int relu) {
    typedef typename Ops::T T;
    if (nrows <= 0 || P <= 0) return 0;
    long long items = (long long)nrows * P;
    int block = 256;
    long long g = (items + block - 1) / block;
    if (g > 0x7fffffffLL) return fail(h, SGRACE_EUNSUPPORTED, "problem too large for exact kernel");
    int grid = (int)g;
#define LAUNCH_LAT(L)                                                                              \
    stage_exact_kernel<Ops, L><<<grid, block, 0, h->stream>>>(rp, ci, (const T*)va, (const T*)Bm, \
                                                              (T*)out, nrows, P, hw_threads, sblock, dense_M, relu)
    switch (lat) {
        case 1: LAUNCH_LAT(1); break;
        case 2: LAUNCH_LAT(2); break;
        case 3: LAUNCH_LAT(3); break;
        case 4: LAUNCH_LAT(4); break;
        case 5: LAUNCH_LAT(5); break;
        case 6: LAUNCH_LAT(6); break;
        case 7: LAUNCH_LAT(7); break;
        case 8: LAUNCH_LAT(8); break;
        default: return fail(h, SGRACE_EUNSUPPORTED, "FADD latency %d not in 1..8", lat);
    }
#undef LAUNCH_LAT
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

int stage_exact_any(sgrace_handle* h, int mode, int lat, const int* rp, const int* ci, const void* va,
                    const void* Bm, void* out, int nrows, int P, int hw_threads, int sblock, int dense_M,
                    int relu) {
    switch (mode) {
        case SGRACE_MODE_F32_CSIM:
            return stage_exact<OpsF32>(h, lat, rp, ci, va, Bm, out, nrows, P, hw_threads, sblock, dense_M, relu);
        case SGRACE_MODE_F16_CSIM:
            if (P % 2 == 0 && ((((uintptr_t)Bm) | ((uintptr_t)out)) & 3) == 0 && nrows > 0 && lat >= 1 && lat <= 8) {
                // two columns per thread on the packed half pipes, same order bit for bit
                const long long items = (long long)nrows * (P / 2);
                const long long g = (items + 255) / 256;
                if (g > 0x7fffffffLL) return fail(h, SGRACE_EUNSUPPORTED, "problem too large for exact kernel");
#define LAUNCH_X2(L)                                                                                                          \
    do {                                                                                                                      \
        if (dense_M > 0) stage_exact_f16x2_kernel<L, true><<<(int)g, 256, 0, h->stream>>>(rp, ci, (const unsigned short*)va, (const unsigned*)Bm, \
                                                                                          (unsigned*)out, nrows, P / 2, hw_threads, sblock, dense_M, relu); \
        else stage_exact_f16x2_kernel<L, false><<<(int)g, 256, 0, h->stream>>>(rp, ci, (const unsigned short*)va, (const unsigned*)Bm, \
                                                                               (unsigned*)out, nrows, P / 2, hw_threads, sblock, dense_M, relu); \
    } while (0)
                switch (lat) {
                    case 1: LAUNCH_X2(1); break; case 2: LAUNCH_X2(2); break; case 3: LAUNCH_X2(3); break; case 4: LAUNCH_X2(4); break;
                    case 5: LAUNCH_X2(5); break; case 6: LAUNCH_X2(6); break; case 7: LAUNCH_X2(7); break; default: LAUNCH_X2(8); break;
                }
#undef LAUNCH_X2
                h->launches++;
                CU(cudaGetLastError());
                return 0;
            }
            return stage_exact<OpsF16>(h, lat, rp, ci, va, Bm, out, nrows, P, hw_threads, sblock, dense_M, relu);
        case SGRACE_MODE_FIX16_CSIM:
            return stage_exact<OpsFix16>(h, 1, rp, ci, va, Bm, out, nrows, P, hw_threads, sblock, dense_M, relu);
    }
    return fail(h, SGRACE_EINVAL, "bad exact mode %d", mode);
}

template <typename T>
int transpose_b(sgrace_handle* h, const void* B, void* Wrm, int M, int P) {
    dim3 grid((M + 31) / 32, (P + 31) / 32), block(32, 8);
    transpose_b_kernel<T><<<grid, block, 0, h->stream>>>((const T*)B, (T*)Wrm, M, P);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

int make_rowptr(sgrace_handle* h, Scratch& s, const int* rows, int nnz, int n, const int** out) {
    if (int rc = ensure(h, s, sizeof(int) * (size_t)(n + 1))) return rc;
    int grid = (n + 1 + 255) / 256;
    coo_rows_to_rowptr_kernel<<<grid, 256, 0, h->stream>>>(rows, nnz, n, (int*)s.p);
    h->launches++;
    CU(cudaGetLastError());
    *out = (const int*)s.p;
    return 0;
}

// dense X (N x M) times W given as the B buffer (W transposed, P x M): tensor cores when the shape
// is a real contraction, CUDA cores otherwise.  Needs h->wrm = W row-major for the CUDA-core kernel.
int dense_f32(sgrace_handle* h, const float* X, const float* B, float* out, int N, int M, int P, int relu) {
    if (h->dense_tc && fea_dense_tc_supported(N, M, P)) {
        int rc = fea_dense_tc_launch(X, B, out, N, M, P, h->num_sms, h->stream, 0, relu);
        if (rc == 0) { h->launches++; return 0; }
        if (rc != -100) return fail(h, SGRACE_ECUDA, "tcgen05 dense FEA launch failed (%d)", rc);
    }
    dim3 grid((N + 63) / 64, (P + 63) / 64);
    fea_dense_f32_kernel<<<grid, 256, 0, h->stream>>>(X, (const float*)h->wrm.p, out, N, M, P, relu);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

// one cooperative launch per small layer; -100 = not eligible (grid does not fit / no cooperative launch)
template <int LPR, int NV, bool SPLIT>
int launch_fused_small_k(sgrace_handle* h, const sgrace_layer_desc* d, const int* rp_fea, const int* rp_adj, void* XW) {
    const int N = d->N_adj, M = d->M_fea, P = d->P_w;
    if (int rc = ensure(h, h->wrm, sizeof(float) * (size_t)M * P)) return rc;
    { int *c0, *c1; if (int rc = counter_sets(h, &c0, &c1)) return rc; h->counter_phase--; }   // allocation only
    auto kern = fused_small_layer_f32_kernel<LPR, NV, SPLIT>;
    static int per_sm = -1;                     // per instantiation; one device kind per process
    if (per_sm < 0) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
    if (per_sm < 1) return -100;
    constexpr int RPW = SPLIT ? 1 : 32 / LPR;
    static const int cap_per_sm = env_int("SGRACE_FUSED_CTAS_PER_SM", 4);
    long long want = ((long long)N + RPW * 8 - 1) / (RPW * 8);        // 8 warps per CTA
    long long cap = (long long)h->num_sms * (per_sm > cap_per_sm ? cap_per_sm : per_sm);
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    const int* ci_f = d->columnIndex_fea; const float* va_f = (const float*)d->values_fea;
    const int* ci_a = d->columnIndex_adj; const float* va_a = (const float*)d->values_adj;
    const float* Bp = (const float*)d->B; float* Wrm = (float*)h->wrm.p;
    float4* XWp = (float4*)XW; float4* Dp = (float4*)d->D;
    int Nn = N, Mm = M, Pp = P, relu = d->relu != 0;
    static int phases = env_int("SGRACE_FUSED_PHASES", 7);
    unsigned* bar = (unsigned*)h->counters.p + 48;      // ints 48,49: grid barrier {count, generation}
    void* args[] = {&rp_fea, &ci_f, &va_f, &rp_adj, &ci_a, &va_a, &Bp, &Wrm, &XWp, &Dp, &Nn, &Mm, &Pp, &relu, &bar, &phases};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(256), args, 0, h->stream);
    if (e != cudaSuccess) { cudaGetLastError(); return -100; }
    h->launches++;
    return 0;
}

template <int LPR, int NV>
int launch_fused_small_t(sgrace_handle* h, const sgrace_layer_desc* d, const int* rp_fea, const int* rp_adj, void* XW) {
    // below this many rows a warp per row (lane groups share the row's non-zeros) shortens the critical path
    static const int split_rows = env_int("SGRACE_FUSED_SPLIT_ROWS", 16384);
    if (LPR < 32 && d->N_adj <= split_rows) return launch_fused_small_k<LPR, NV, true>(h, d, rp_fea, rp_adj, XW);
    return launch_fused_small_k<LPR, NV, false>(h, d, rp_fea, rp_adj, XW);
}

int launch_fused_small(sgrace_handle* h, const sgrace_layer_desc* d, const int* rp_fea, const int* rp_adj, void* XW) {
    const int P4 = d->P_w / 4;
    if (P4 == 1) return launch_fused_small_t<1, 1>(h, d, rp_fea, rp_adj, XW);
    if (P4 == 2) return launch_fused_small_t<2, 1>(h, d, rp_fea, rp_adj, XW);
    if (P4 <= 4) return launch_fused_small_t<4, 1>(h, d, rp_fea, rp_adj, XW);
    if (P4 <= 8) return launch_fused_small_t<8, 1>(h, d, rp_fea, rp_adj, XW);
    if (P4 <= 16) return launch_fused_small_t<16, 1>(h, d, rp_fea, rp_adj, XW);
    if (P4 <= 32) return launch_fused_small_t<32, 1>(h, d, rp_fea, rp_adj, XW);
    if (P4 <= 64) return launch_fused_small_t<32, 2>(h, d, rp_fea, rp_adj, XW);
    if (P4 <= 128) return launch_fused_small_t<32, 4>(h, d, rp_fea, rp_adj, XW);
    return launch_fused_small_t<32, 8>(h, d, rp_fea, rp_adj, XW);
}

QConst make_qconst(const sgrace_handle* h, const sgrace_layer_desc* d) {
    QConst q;
    memset(&q, 0, sizeof(q));
    q.inv_fs = d->qscale_fea; q.inv_ws = d->qscale_w; q.inv_as = d->qscale_adj;
    q.f_z = q.w_z = q.a_z = 0;   // the driver never programs zero points (sgrace.py:334-365)
    q.qbits = h->qbits;
    q.den = h->qbits == 1 ? 2.0f : (float)(1 << (h->qbits > 0 ? h->qbits - 1 : 0));
    q.wh_den = q.den * q.den;
    q.wh_scale = (float)(1 << d->scale_fea);
    const int iq = d->internal_quantization;
    q.a_hi = (float)((ldexp(1.0, iq) - 1.0) / ldexp(1.0, iq));
    q.round_T = (float)pow(10.0, (double)(iq - 1));
    q.deq_o = d->deq_factor;
    q.alpha = h->leaky_alpha;
    return q;
}

// ------------------------------------------------------------------------------------
// the two stages
// ------------------------------------------------------------------------------------
int resolve_rowptrs(sgrace_handle* h, const sgrace_layer_desc* d, const int** rp_fea, const int** rp_adj,
                    bool need_fea, bool need_adj) {
    *rp_fea = d->rowPtr_fea;
    *rp_adj = d->rowPtr_adj;
    if (h->index_format == 1) {
        if (need_fea && d->gemm_mode == 0)
            if (int rc = make_rowptr(h, h->rp_fea, d->rowPtr_fea, d->nnz_fea, d->N_adj, rp_fea)) return rc;
        if (need_adj)
            if (int rc = make_rowptr(h, h->rp_adj, d->rowPtr_adj, d->nnz_adj, d->N_adj, rp_adj)) return rc;
    }
    return 0;
}

int run_fea(sgrace_handle* h, const sgrace_layer_desc* d, const int* rp_fea, void* XW) {
    const int N = d->N_adj, M = d->M_fea, P = d->P_w;
    if (N <= 0 || P <= 0) return 0;
    if (!d->B || !d->values_fea) return fail(h, SGRACE_EINVAL, "B / values_fea pointer not set");
    if (d->gemm_mode == 0 && (!rp_fea || !d->columnIndex_fea))
        return fail(h, SGRACE_EINVAL, "rowPtr_fea / columnIndex_fea pointer not set");
    const size_t esz = elt_bytes(h->mode);
    switch (h->mode) {
        case SGRACE_MODE_F32_FAST: {
            if (int rc = ensure(h, h->wrm, esz * (size_t)M * P)) return rc;
            if (int rc = transpose_b<float>(h, d->B, h->wrm.p, M, P)) return rc;
            if (d->gemm_mode == 0)
                return spmm_f32(h, rp_fea, d->columnIndex_fea, (const float*)d->values_fea,
                                (const float*)h->wrm.p, (float*)XW, N, P, 0, d->nnz_fea, M, 0);
            return dense_f32(h, (const float*)d->values_fea, (const float*)d->B, (float*)XW, N, M, P, 0);
        }
        case SGRACE_MODE_F32_CSIM:
        case SGRACE_MODE_F16_CSIM:
        case SGRACE_MODE_FIX16_CSIM: {
            if (int rc = ensure(h, h->wrm, esz * (size_t)M * P)) return rc;
            if (esz == 4) { if (int rc = transpose_b<float>(h, d->B, h->wrm.p, M, P)) return rc; }
            else          { if (int rc = transpose_b<unsigned short>(h, d->B, h->wrm.p, M, P)) return rc; }
            const int lat = h->lat_fea > 0 ? h->lat_fea : default_lat(h->mode);
            return stage_exact_any(h, h->mode, lat, rp_fea, d->columnIndex_fea, d->values_fea, h->wrm.p, XW, N, P,
                                   h->fea_threads, h->spmm_block, d->gemm_mode ? M : 0, 0);
        }
        case SGRACE_MODE_FULL: {
            if (h->qbits == 0) {   // fake_quantization == 0: float32, multiply-then-add in row order
                if (int rc = ensure(h, h->wrm, 4 * (size_t)M * P)) return rc;
                if (int rc = transpose_b<float>(h, d->B, h->wrm.p, M, P)) return rc;
                return stage_exact<OpsF32>(h, 1, rp_fea, d->columnIndex_fea, d->values_fea, h->wrm.p, XW, N, P, 1, 1,
                                           d->gemm_mode ? M : 0, 0);
            }
            const QConst qc = make_qconst(h, d);
            const int Pp = (P + 3) & ~3;
            if (int rc = ensure(h, h->wq, (size_t)M * Pp)) return rc;
            quantize_w_kernel<<<(M * Pp + 255) / 256, 256, 0, h->stream>>>((const float*)d->B, (signed char*)h->wq.p,
                                                                           M, P, Pp, qc);
            h->launches++;
            CU(cudaGetLastError());
            if (!h->max_fea_dev) CU(cudaMalloc(&h->max_fea_dev, sizeof(int)));
            CU(cudaMemsetAsync(h->max_fea_dev, 0, sizeof(int), h->stream));
            long long items = (long long)N * (Pp / 4);
            int grid = (int)((items + 255) / 256);
            if (d->gemm_mode == 0)
                fea_q_csr_kernel<<<grid, 256, 0, h->stream>>>(rp_fea, d->columnIndex_fea, (const float*)d->values_fea,
                                                              (const signed char*)h->wq.p, (float*)XW, N, P, Pp, qc,
                                                              h->max_fea_dev);
            else
                fea_q_dense_kernel<<<grid, 256, 0, h->stream>>>((const float*)d->values_fea,
                                                                (const signed char*)h->wq.p, (float*)XW, N, M, P, Pp,
                                                                qc, h->max_fea_dev);
            h->launches++;
            CU(cudaGetLastError());
            return 0;
        }
    }
    return fail(h, SGRACE_EINVAL, "bad mode %d", h->mode);
}

int run_adj(sgrace_handle* h, const sgrace_layer_desc* d, const int* rp_adj, const void* XW, int xw_rows) {
    const int N = d->N_adj, P = d->P_w;
    if (N <= 0 || P <= 0) return 0;
    if (!rp_adj || !d->columnIndex_adj || !d->values_adj || !d->D)
        return fail(h, SGRACE_EINVAL, "adjacency / D pointer not set");
    // USE_SBLOCKS == 1 drops the ReLU in the reference's write stage (kernelMatrixmult_all.cpp:748-786)
    const int relu = (h->use_sblocks && h->mode != SGRACE_MODE_FULL && h->mode != SGRACE_MODE_F32_FAST)
                         ? 0 : (d->relu != 0);
    switch (h->mode) {
        case SGRACE_MODE_F32_FAST:
            return spmm_f32(h, rp_adj, d->columnIndex_adj, (const float*)d->values_adj, (const float*)XW,
                            (float*)d->D, N, P, relu, d->nnz_adj, 0, 1, xw_rows);
        case SGRACE_MODE_F32_CSIM:
        case SGRACE_MODE_F16_CSIM:
        case SGRACE_MODE_FIX16_CSIM: {
            const int lat = h->lat_adj > 0 ? h->lat_adj : default_lat(h->mode);
            return stage_exact_any(h, h->mode, lat, rp_adj, d->columnIndex_adj, d->values_adj, XW, d->D, N, P,
                                   h->adj_threads, h->spmm_block, 0, relu);
        }
        case SGRACE_MODE_FULL: {
            const QConst qc = make_qconst(h, d);
            const int quant = h->qbits > 0;
            const bool vec = P % 4 == 0 && P <= 1024 && ((((uintptr_t)XW) | ((uintptr_t)d->D)) & 15) == 0;
            if (!d->gat_mode) {
                // the streaming kernel (TMA-staged CSR slices) with the quantised arithmetic, for the widths whose row
                // is exactly LPR lanes x one float4
                const int P4s = P / 4;
                const bool csr_al = ((((uintptr_t)rp_adj) | ((uintptr_t)d->columnIndex_adj) | ((uintptr_t)d->values_adj)) & 15) == 0;
                if (vec && csr_al && h->stream_kernel && !h->accumulate && h->peer_count == 0 &&
                    (P4s == 1 || P4s == 2 || P4s == 4 || P4s == 8 || P4s == 16 || P4s == 32)) {
#define SGRACE_QSTREAM(L) return launch_spmm_stream<L, 1>(h, rp_adj, d->columnIndex_adj, (const float*)d->values_adj, (const float*)XW, \
                                                          (float*)d->D, N, P, relu, d->nnz_adj, 0, 1, xw_rows, &qc, quant)
                    if (P4s == 1) SGRACE_QSTREAM(1);
                    if (P4s == 2) SGRACE_QSTREAM(2);
                    if (P4s == 4) SGRACE_QSTREAM(4);
                    if (P4s == 8) SGRACE_QSTREAM(8);
                    if (P4s == 16) SGRACE_QSTREAM(16);
                    SGRACE_QSTREAM(32);
#undef SGRACE_QSTREAM
                }
                if (vec) {
                    const int P4 = P / 4;
#define SGRACE_ADJQ(L, V) adj_q_gcn_vec_kernel<L, V><<<grid_for((long long)N * L, 256, h->num_sms, 64), 256, 0, h->stream>>>( \
                        rp_adj, d->columnIndex_adj, (const float*)d->values_adj, (const float4*)XW, (float4*)d->D, N, P4, relu, quant, qc)
                    SGRACE_P4_DISPATCH(P4, SGRACE_ADJQ);
#undef SGRACE_ADJQ
                } else {
                    long long items = (long long)N * P;
                    adj_q_gcn_kernel<<<(int)((items + 255) / 256), 256, 0, h->stream>>>(
                        rp_adj, d->columnIndex_adj, (const float*)d->values_adj, (const float*)XW, (float*)d->D, N, P,
                        relu, quant, qc);
                }
                h->launches++;
                CU(cudaGetLastError());
                return 0;
            }
            if (!d->attention) return fail(h, SGRACE_EINVAL, "gat_mode set but ate_m (attention) pointer missing");
            if (h->row_offset < 0 || (long long)h->row_offset + N > xw_rows)
                return fail(h, SGRACE_EINVAL, "row_offset %d + N_adj %d exceeds the %d rows of the feature-stage result", h->row_offset, N, xw_rows);
            if (int rc = ensure(h, h->s1, sizeof(float) * (size_t)xw_rows)) return rc;
            if (int rc = ensure(h, h->s2, sizeof(float) * (size_t)xw_rows)) return rc;
            if (int rc = ensure(h, h->lists, sizeof(int) * (size_t)N)) return rc;
            { int *c0, *c1; if (int rc = counter_sets(h, &c0, &c1)) return rc; h->counter_phase--; }
            int* empty_count = (int*)h->counters.p + 1;
            CU(cudaMemsetAsync(empty_count, 0, sizeof(int), h->stream));
            if (vec) {
                const int P4 = P / 4;
                gat_scores_vec_kernel<<<(xw_rows + 255) / 256, 256, sizeof(float) * 2 * P, h->stream>>>(
                    (const float4*)XW, d->attention, (float*)h->s1.p, (float*)h->s2.p, xw_rows, P4, quant, qc);
                h->launches++;
                CU(cudaGetLastError());
#define SGRACE_GATQ(L, V) gat_aggregate_vec_kernel<L, V><<<(unsigned)(((long long)N * L + 255) / 256), 256, 0, h->stream>>>( \
                    rp_adj, d->columnIndex_adj, (const float*)d->values_adj, (const float4*)XW, (const float*)h->s1.p,       \
                    (const float*)h->s2.p, (float4*)d->D, d->E, d->S, N, P4, relu, quant, qc, (int*)h->lists.p, empty_count, h->row_offset)
                SGRACE_P4_DISPATCH(P4, SGRACE_GATQ);
#undef SGRACE_GATQ
            } else {
                gat_scores_kernel<<<(xw_rows + 255) / 256, 256, 0, h->stream>>>((const float*)XW, d->attention,
                                                                                (float*)h->s1.p, (float*)h->s2.p, xw_rows,
                                                                                P, quant, qc);
                h->launches++;
                CU(cudaGetLastError());
                int grid = grid_for((long long)N * 32, 256, h->num_sms, 8);
                gat_aggregate_kernel<<<grid, 256, 0, h->stream>>>(rp_adj, d->columnIndex_adj, (const float*)d->values_adj,
                                                                  (const float*)XW, (const float*)h->s1.p,
                                                                  (const float*)h->s2.p, (float*)d->D, d->E, d->S, N, P,
                                                                  relu, quant, qc, (int*)h->lists.p, empty_count, h->row_offset);
            }
            h->launches++;
            CU(cudaGetLastError());
            gat_empty_rows_kernel<<<(P + 127) / 128, 128, 0, h->stream>>>((const float*)XW, (float*)d->D, xw_rows, P,
                                                                          relu, quant, qc, (const int*)h->lists.p,
                                                                          empty_count);
            h->launches++;
            CU(cudaGetLastError());
            return 0;
        }
    }
    return fail(h, SGRACE_EINVAL, "bad mode %d", h->mode);
}

int check_desc(sgrace_handle* h, const sgrace_layer_desc* d) {
    if (!d) return fail(h, SGRACE_EINVAL, "null descriptor");
    if (d->N_adj < 0 || d->M_fea < 0 || d->P_w < 0) return fail(h, SGRACE_EINVAL, "negative dimension");
    if (d->gemm_mode == 2) {
        // the full design's backward launch (sgrace.py:717-760): dense "adjacency" operand, sparse "feature" operand
        if (h->mode != SGRACE_MODE_F32_FAST)
            return fail(h, SGRACE_EUNSUPPORTED,
                        "gemm_mode=2 (hardware backward, sgrace.py:717) runs in the float32 mode only: the fixed-point "
                        "arithmetic of that launch is not specified by the open sources");
        if (d->M_adj < 0 || d->P_w % 4) return fail(h, SGRACE_EINVAL, "gemm_mode=2 needs M_adj >= 0 and P_w a multiple of 4");
        if (h->index_format == 1 && d->nnz_fea < 0) return fail(h, SGRACE_EINVAL, "COO index format needs the non-zero count");
        return 0;
    }
    if (d->gemm_mode != 0 && d->gemm_mode != 1) return fail(h, SGRACE_EINVAL, "gemm_mode=%d", d->gemm_mode);
    if (h->index_format == 1 && ((d->gemm_mode == 0 && d->nnz_fea < 0) || d->nnz_adj < 0))
        return fail(h, SGRACE_EINVAL, "COO index format needs nnz_fea1 / nnz_adj1");
    if (h->mode == SGRACE_MODE_FULL && h->qbits > 0) {
        if (d->internal_quantization < 1 || d->internal_quantization > 30 || d->scale_fea < 0 || d->scale_fea > 30)
            return fail(h, SGRACE_EINVAL, "scale_fea / quantized_multiplier registers out of range");
        if (!(d->qscale_fea > 0.f) || !(d->qscale_w > 0.f) || !(d->qscale_adj > 0.f))
            return fail(h, SGRACE_EINVAL, "quantization_scale_* registers not programmed");
    }
    return 0;
}

int layer_run_impl(sgrace_handle* h, const sgrace_layer_desc* d, bool timed) {
    if (int rc = check_desc(h, d)) return rc;
    const size_t esz = elt_bytes(h->mode);
    void* XW = d->XW;
    if (!XW) {
        if (int rc = ensure(h, h->xw, esz * (size_t)d->N_adj * (size_t)d->P_w + 16)) return rc;
        XW = h->xw.p;
    }
    const int *rp_fea, *rp_adj;
    if (timed) CU(cudaEventRecord(h->ev[0], h->stream));
    if (d->gemm_mode == 2) {
        // D[N_adj x P] = act( Adense[N_adj x M_adj] . ( S[M_adj x M_fea] . W ) ):  grad_W = X^T (A g) with
        // Adense = X^T, S = A, W = g (the B buffer holds g transposed, as always)
        if (!d->values_adj || !d->D) return fail(h, SGRACE_EINVAL, "values_adj / D pointer not set");
        sgrace_layer_desc f = *d;
        f.gemm_mode = 0;
        f.N_adj = d->M_adj;
        if (int rc = ensure(h, h->xw, esz * (size_t)d->M_adj * (size_t)d->P_w + 16)) return rc;
        XW = h->xw.p;
        if (int rc = resolve_rowptrs(h, &f, &rp_fea, &rp_adj, true, false)) return rc;
        if (int rc = run_fea(h, &f, rp_fea, XW)) return rc;
        if (timed) CU(cudaEventRecord(h->ev[1], h->stream));
        const int R = d->N_adj, K = d->M_adj, P4 = d->P_w / 4;
        if (R > 0 && P4 > 0) {
            if ((((uintptr_t)d->D) & 15) != 0) return fail(h, SGRACE_EINVAL, "gemm_mode=2 needs a 16-byte aligned D");
            const long long warps = (long long)R * ((P4 + 31) / 32);
            dense_adj_f32_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, h->stream>>>(
                (const float*)d->values_adj, (const float4*)XW, (float4*)d->D, R, K, P4, d->relu != 0);
            h->launches++;
            CU(cudaGetLastError());
        }
        if (timed) CU(cudaEventRecord(h->ev[2], h->stream));
        return 0;
    }
    if (int rc = resolve_rowptrs(h, d, &rp_fea, &rp_adj, true, true)) return rc;
    // Small sparse-feature layers: one cooperative launch for the whole layer (W transpose | FEA | ADJ)
    if (h->mode == SGRACE_MODE_F32_FAST && d->gemm_mode == 0 && h->fused_small && d->N_adj > 0 && d->N_adj <= h->fused_small &&
        d->P_w % 4 == 0 && d->P_w <= 1024 && d->M_fea > 0 && !h->accumulate &&
        (((uintptr_t)XW | (uintptr_t)d->D) & 15) == 0 && d->B && d->values_fea && d->columnIndex_fea && rp_fea && rp_adj &&
        d->columnIndex_adj && d->values_adj && d->D) {
        const int rc = launch_fused_small(h, d, rp_fea, rp_adj, XW);
        if (rc != -100) {
            if (rc) return rc;
            if (timed) { CU(cudaEventRecord(h->ev[1], h->stream)); CU(cudaEventRecord(h->ev[2], h->stream)); }
            return 0;
        }
    }
    // Opt-in aggregate-first order for a dense layer that widens (M_fea < P_w):  D = act((A.X).W).
    // Equal to act(A.(X.W)) up to float rounding; gathers M_fea-wide rows instead of P_w-wide ones.
    if (h->mode == SGRACE_MODE_F32_FAST && h->agg_first && d->gemm_mode == 1 && d->M_fea < d->P_w && d->M_fea % 4 == 0 &&
        d->N_adj > 0 && d->P_w > 0) {
        const int N = d->N_adj, M = d->M_fea, P = d->P_w;
        if (!d->B || !d->values_fea || !d->D) return fail(h, SGRACE_EINVAL, "B / values_fea / D pointer not set");
        if (int rc = ensure(h, h->ax, sizeof(float) * (size_t)N * M + 16)) return rc;
        if (int rc = ensure(h, h->wrm, sizeof(float) * (size_t)M * P)) return rc;
        if (int rc = transpose_b<float>(h, d->B, h->wrm.p, M, P)) return rc;
        if (int rc = spmm_f32(h, rp_adj, d->columnIndex_adj, (const float*)d->values_adj, (const float*)d->values_fea,
                              (float*)h->ax.p, N, M, 0, d->nnz_adj, 0, 0, N)) return rc;
        if (timed) CU(cudaEventRecord(h->ev[1], h->stream));
        if (int rc = dense_f32(h, (const float*)h->ax.p, (const float*)d->B, (float*)d->D, N, M, P, d->relu != 0)) return rc;
        if (timed) CU(cudaEventRecord(h->ev[2], h->stream));
        return 0;
    }
    if (int rc = run_fea(h, d, rp_fea, XW)) return rc;
    if (timed) CU(cudaEventRecord(h->ev[1], h->stream));
    if (int rc = run_adj(h, d, rp_adj, XW, d->N_adj)) return rc;
    if (timed) CU(cudaEventRecord(h->ev[2], h->stream));
    return 0;
}

// ------------------------------------------------------------------------------------
// register file helpers
// ------------------------------------------------------------------------------------
struct RegName { const char* name; uint32_t off; };
const RegName kRegNames[] = {
    {"CTRL", 0x00}, {"GIER", 0x04}, {"IP_IER", 0x08}, {"IP_ISR", 0x0c},
    {"load_weights", 0x10}, {"beta_qu", 0x18}, {"f_align", 0x20},
    {"quantization_scale_adj", 0x28}, {"quantization_scale_fea", 0x30}, {"quantization_scale_w", 0x38},
    {"deq_factor", 0x40}, {"stream_mode", 0x48}, {"gat_mode", 0x50}, {"gemm_mode", 0x58}, {"relu", 0x60},
    {"scale_fea", 0x68}, {"max_fea", 0x70}, {"max_fea_ctrl", 0x74}, {"layer_count", 0x80},
    {"quantized_multiplier", 0x88}, {"shift_offset_1", 0x90}, {"shift_offset_2", 0x94},
    {"bias_offset_1", 0x9c}, {"bias_offset_2", 0xa0}, {"bias_count", 0xa8},
    {"profiling_offset_1", 0xb0}, {"profiling_offset_2", 0xb4},
    {"zero_point_lhs", 0xbc}, {"zero_point_rhs", 0xc4}, {"zero_point_dst", 0xcc},
    {"clamp_max", 0xd4}, {"clamp_min", 0xdc},
    {"N_adj", 0xe4}, {"M_adj", 0xec}, {"M_fea", 0xf4}, {"P_w", 0xfc},
    {"B_offset_1", 0x104}, {"B_offset_2", 0x108},
    {"D1_offset_1", 0x110}, {"D1_offset_2", 0x114}, {"D2_offset_1", 0x11c}, {"D2_offset_2", 0x120},
    {"D3_offset_1", 0x128}, {"D3_offset_2", 0x12c}, {"D4_offset_1", 0x134}, {"D4_offset_2", 0x138},
    {"E1_offset_1", 0x140}, {"E1_offset_2", 0x144}, {"S1_offset_1", 0x14c}, {"S1_offset_2", 0x150},
    {"ate_m_offset_1", 0x158}, {"ate_m_offset_2", 0x15c}, {"array_c_adjust", 0x164},
    {"nnz_fea1", 0x16c}, {"nnz_fea2", 0x174}, {"nnz_fea3", 0x17c}, {"nnz_fea4", 0x184},
    {"rowPtr_fea1_offset_1", 0x18c}, {"rowPtr_fea1_offset_2", 0x190},
    {"rowPtr_fea2_offset_1", 0x198}, {"rowPtr_fea2_offset_2", 0x19c},
    {"rowPtr_fea3_offset_1", 0x1a4}, {"rowPtr_fea3_offset_2", 0x1a8},
    {"rowPtr_fea4_offset_1", 0x1b0}, {"rowPtr_fea4_offset_2", 0x1b4},
    {"columnIndex_fea1_offset_1", 0x1bc}, {"columnIndex_fea1_offset_2", 0x1c0},
    {"columnIndex_fea2_offset_1", 0x1c8}, {"columnIndex_fea2_offset_2", 0x1cc},
    {"columnIndex_fea3_offset_1", 0x1d4}, {"columnIndex_fea3_offset_2", 0x1d8},
    {"columnIndex_fea4_offset_1", 0x1e0}, {"columnIndex_fea4_offset_2", 0x1e4},
    {"values_fea1_offset_1", 0x1ec}, {"values_fea1_offset_2", 0x1f0},
    {"values_fea2_offset_1", 0x1f8}, {"values_fea2_offset_2", 0x1fc},
    {"values_fea3_offset_1", 0x204}, {"values_fea3_offset_2", 0x208},
    {"values_fea4_offset_1", 0x210}, {"values_fea4_offset_2", 0x214},
    {"nnz_adj1", 0x21c}, {"nnz_adj2", 0x224}, {"nnz_adj3", 0x22c}, {"nnz_adj4", 0x234},
    {"rowPtr_adj1_offset_1", 0x23c}, {"rowPtr_adj1_offset_2", 0x240},
    {"rowPtr_adj2_offset_1", 0x248}, {"rowPtr_adj2_offset_2", 0x24c},
    {"rowPtr_adj3_offset_1", 0x254}, {"rowPtr_adj3_offset_2", 0x258},
    {"rowPtr_adj4_offset_1", 0x260}, {"rowPtr_adj4_offset_2", 0x264},
    {"columnIndex_adj1_offset_1", 0x26c}, {"columnIndex_adj1_offset_2", 0x270},
    {"columnIndex_adj2_offset_1", 0x278}, {"columnIndex_adj2_offset_2", 0x27c},
    {"columnIndex_adj3_offset_1", 0x284}, {"columnIndex_adj3_offset_2", 0x288},
    {"columnIndex_adj4_offset_1", 0x290}, {"columnIndex_adj4_offset_2", 0x294},
    {"values_adj1_offset_1", 0x29c}, {"values_adj1_offset_2", 0x2a0},
    {"values_adj2_offset_1", 0x2a8}, {"values_adj2_offset_2", 0x2ac},
    {"values_adj3_offset_1", 0x2b4}, {"values_adj3_offset_2", 0x2b8},
    {"values_adj4_offset_1", 0x2c0}, {"values_adj4_offset_2", 0x2c4},
    {"quantized_multiplier_offset_1", 0x400}, {"quantized_multiplier_offset_2", 0x404},
};

inline uint64_t reg64(const sgrace_handle* h, uint32_t off) {
    return (uint64_t)h->regs[off / 4] | ((uint64_t)h->regs[off / 4 + 1] << 32);
}
inline float regf(const sgrace_handle* h, uint32_t off) {
    float f;
    memcpy(&f, &h->regs[off / 4], 4);
    return f;
}

// find the sgrace_alloc buffer containing device address a
const Buffer* find_buffer(const sgrace_handle* h, uint64_t a, size_t* offset) {
    if (h->buffers.empty() || a == 0) return nullptr;
    auto it = h->buffers.upper_bound(a);
    if (it == h->buffers.begin()) return nullptr;
    --it;
    if (a >= it->first && a < it->first + it->second.bytes) {
        *offset = (size_t)(a - it->first);
        return &it->second;
    }
    return nullptr;
}

const void* host_view(const sgrace_handle* h, uint64_t addr) {
    size_t off;
    const Buffer* b = find_buffer(h, addr, &off);
    return b ? (const char*)b->host + off : nullptr;
}

// largest column index referenced by each row panel of a CSR adjacency: panel i = non-zeros [k0[i], k0[i+1])
__global__ void panel_maxcol_kernel(const int* __restrict__ col, const long long* __restrict__ k0, int npanels, int* __restrict__ out) {
    const int i = blockIdx.y;
    if (i >= npanels) return;
    int m = -1;
    for (long long k = k0[i] + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < k0[i + 1]; k += (long long)gridDim.x * blockDim.x)
        m = max(m, __ldg(col + k));
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m >= 0) atomicMax(out + i, m);
}

// host mirror -> device for bytes [skip, skip + bytes) behind device address addr, on stream st (0: the handle's)
int stage_in(sgrace_handle* h, uint64_t addr, size_t bytes, size_t skip = 0, cudaStream_t st = nullptr) {
    size_t off;
    const Buffer* b = find_buffer(h, addr, &off);
    if (!b || bytes == 0) return 0;
    if (off + skip + bytes > b->bytes)
        return fail(h, SGRACE_EBOUNDS, "layer needs %zu bytes at buffer offset %zu but the allocation has %zu", skip + bytes,
                    off, b->bytes);
    CU(cudaMemcpyAsync((char*)b->dev + off + skip, (char*)b->host + off + skip, bytes, cudaMemcpyHostToDevice, st ? st : h->stream));
    return 0;
}
int stage_out(sgrace_handle* h, uint64_t addr, size_t bytes, size_t skip = 0, cudaStream_t st = nullptr) {
    size_t off;
    const Buffer* b = find_buffer(h, addr, &off);
    if (!b || bytes == 0) return 0;
    if (off + skip + bytes > b->bytes)
        return fail(h, SGRACE_EBOUNDS, "layer writes %zu bytes at buffer offset %zu but the allocation has %zu", skip + bytes,
                    off, b->bytes);
    CU(cudaMemcpyAsync((char*)b->host + off + skip, (char*)b->dev + off + skip, bytes, cudaMemcpyDeviceToHost, st ? st : h->stream));
    return 0;
}

// Banded / block-diagonal adjacencies (batched graphs): the whole layer is pipelined in row chunks, so that D comes down
// while the features are still going up.  The column array goes up first and a small kernel finds the largest column each
// row panel references (exact, on the device); the host reads those K numbers, then queues: feature chunk c up -> feature
// stage on chunk c -> every adjacency panel whose columns are now covered -> that panel of D down.  The order of the
// launches follows the exact maxima, so the result does not depend on the guess that made us come here (a host sample of
// the column array); a general graph simply runs all its panels after the last chunk.  -100: does not qualify.
int start_pipelined(sgrace_handle* h, sgrace_layer_desc& d, uint64_t a_rpf, uint64_t a_cif, uint64_t a_vf, uint64_t a_rpa,
                    uint64_t a_cia, uint64_t a_va, uint64_t a_b, uint64_t a_d, long long nnz_fea, long long nnz_adj) {
    const size_t N = (size_t)d.N_adj, M = (size_t)d.M_fea, P = (size_t)d.P_w;
    if (!h->overlap || h->mode != SGRACE_MODE_F32_FAST || h->index_format != 0 || h->agg_first || h->accumulate ||
        h->peer_count > 0 || d.gemm_mode == 2 || d.gat_mode || nnz_adj <= 0)
        return -100;
    const size_t d_bytes = N * P * 4;
    if (d_bytes < ((size_t)16 << 20) || (h->fused_small && d.N_adj <= h->fused_small)) return -100;
    // host mirrors of the index arrays (the chunk boundaries are read from them); every operand must sit in a buffer
    // from sgrace_alloc
    const int* rpa = (const int*)host_view(h, a_rpa);
    const int* cia = (const int*)host_view(h, a_cia);
    if (!rpa || !cia || !host_view(h, a_va) || !host_view(h, a_d) || !host_view(h, a_vf) || !host_view(h, a_b)) return -100;
    const int* rpf = nullptr;
    if (d.gemm_mode == 0) {
        rpf = (const int*)host_view(h, a_rpf);
        if (!rpf || !host_view(h, a_cif) || rpf[0] != 0) return -100;
    }
    if (rpa[0] != 0) return -100;
    // panels of D of about 24 MB (12 MB panels were measured no faster: twice the launches for half the tail)
    int K = (int)((d_bytes + ((size_t)24 << 20) - 1) / ((size_t)24 << 20));
    if (K < 4) K = 4;
    if (K > 16) K = 16;
    const size_t rows_per = ((N + K - 1) / K + 127) & ~(size_t)127;
    K = (int)((N + rows_per - 1) / rows_per);
    if (K < 3) return -100;
    // host sample: does a panel look past the chunk after its own?  256 entries per panel, a few microseconds
    for (int i = 0; i + 2 < K; i++) {
        const size_t r0 = (size_t)i * rows_per, r1 = r0 + rows_per;
        const long long k0 = rpa[r0], k1 = rpa[r1];
        if (k0 < 0 || k1 < k0 || k1 > nnz_adj) return -100;
        const long long limit = (long long)(r1 + rows_per);
        for (int sidx = 0; sidx < 256 && k1 > k0; sidx++) {
            const long long k = k0 + (long long)(((unsigned long long)sidx * 2654435761ull) % (unsigned long long)(k1 - k0));
            if (cia[k] >= limit) return -100;
        }
    }
    if (!h->s_up) CU(cudaStreamCreateWithFlags(&h->s_up, cudaStreamNonBlocking));
    if (!h->s_down) CU(cudaStreamCreateWithFlags(&h->s_down, cudaStreamNonBlocking));
    while ((int)h->ev_pool.size() < 2 * 16 + 4) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->ev_pool.push_back(e);
    }
    cudaEvent_t ev_prev = h->ev_pool[32], ev_max = h->ev_pool[33], ev_down = h->ev_pool[34], ev_b = h->ev_pool[35];
    if (!h->pipe_host) CU(cudaHostAlloc(&h->pipe_host, 512, cudaHostAllocDefault));
    if (int rc = ensure(h, h->pipe_dev, 512)) return rc;
    long long* k0_host = (long long*)h->pipe_host;              // [K + 1] panel starts, then [K] maxima (ints)
    int* max_host = (int*)(k0_host + 17);
    for (int i = 0; i <= K; i++) { const size_t r = (size_t)i * rows_per < N ? (size_t)i * rows_per : N; k0_host[i] = rpa[r]; }
    long long* k0_dev = (long long*)h->pipe_dev.p;
    int* max_dev = (int*)(k0_dev + 17);
    CU(cudaEventRecord(ev_prev, h->stream));
    CU(cudaStreamWaitEvent(h->s_up, ev_prev, 0));
    CU(cudaStreamWaitEvent(h->s_down, ev_prev, 0));
    // ---- uploads, all on s_up in the order they are needed ----
    if (int rc = stage_in(h, a_cia, (size_t)nnz_adj * 4, 0, h->s_up)) return rc;
    CU(cudaMemcpyAsync(k0_dev, k0_host, 17 * 8, cudaMemcpyHostToDevice, h->s_up));
    CU(cudaMemsetAsync(max_dev, 0xff, 16 * 4, h->s_up));
    panel_maxcol_kernel<<<dim3(64, K), 256, 0, h->s_up>>>(d.columnIndex_adj, k0_dev, K, max_dev);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(max_host, max_dev, 16 * 4, cudaMemcpyDeviceToHost, h->s_up));
    CU(cudaEventRecord(ev_max, h->s_up));
    if (int rc = stage_in(h, a_rpa, (N + 1) * 4, 0, h->s_up)) return rc;
    if (int rc = stage_in(h, a_va, (size_t)nnz_adj * 4, 0, h->s_up)) return rc;
    if (int rc = stage_in(h, a_b, M * P * 4, 0, h->s_up)) return rc;
    CU(cudaEventRecord(ev_b, h->s_up));
    if (d.gemm_mode == 0) if (int rc = stage_in(h, a_rpf, (N + 1) * 4, 0, h->s_up)) return rc;
    for (int c = 0; c < K; c++) {
        const size_t r0 = (size_t)c * rows_per, r1 = r0 + rows_per < N ? r0 + rows_per : N;
        if (d.gemm_mode == 0) {
            const long long f0 = rpf[r0], f1 = rpf[r1];
            if (f0 < 0 || f1 < f0 || f1 > nnz_fea) return fail(h, SGRACE_EBOUNDS, "feature row pointers not monotonic around row %zu", r0);
            if (int rc = stage_in(h, a_cif, (size_t)(f1 - f0) * 4, (size_t)f0 * 4, h->s_up)) return rc;
            if (int rc = stage_in(h, a_vf, (size_t)(f1 - f0) * 4, (size_t)f0 * 4, h->s_up)) return rc;
        } else {
            if (int rc = stage_in(h, a_vf, (r1 - r0) * M * 4, r0 * M * 4, h->s_up)) return rc;
        }
        CU(cudaEventRecord(h->ev_pool[c], h->s_up));
    }
    void* XW = d.XW;
    if (!XW) {
        if (int rc = ensure(h, h->xw, 4 * N * P + 16)) return rc;
        XW = h->xw.p;
    }
    if (int rc = ensure(h, h->wrm, 4 * M * P)) return rc;
    // ---- the K maxima (the column array has landed; the other uploads keep the copy engine busy meanwhile) ----
    CU(cudaEventSynchronize(ev_max));
    int need[16];                       // chunk that must be done before panel i may run
    for (int i = 0; i < K; i++) {
        const int mc = max_host[i];
        if (mc >= (int)N) return fail(h, SGRACE_EBOUNDS, "adjacency column index %d out of range [0,%zu)", mc, N);
        need[i] = mc < 0 ? 0 : (int)((size_t)mc / rows_per);
    }
    h->ev_valid = false;
    CU(cudaEventRecord(h->ev[0], h->stream));
    CU(cudaStreamWaitEvent(h->stream, ev_b, 0));                  // B has landed
    if (int rc = transpose_b<float>(h, d.B, h->wrm.p, (int)M, (int)P)) return rc;
    bool launched[16] = {false};
    for (int c = 0; c < K; c++) {
        const size_t r0 = (size_t)c * rows_per, r1 = r0 + rows_per < N ? r0 + rows_per : N;
        CU(cudaStreamWaitEvent(h->stream, h->ev_pool[c], 0));
        if (d.gemm_mode == 0) {
            if (int rc = spmm_f32(h, d.rowPtr_fea + r0, d.columnIndex_fea, (const float*)d.values_fea, (const float*)h->wrm.p,
                                  (float*)XW + r0 * P, (int)(r1 - r0), (int)P, 0, (long long)rpf[r1] - rpf[r0], (int)M, 0)) return rc;
        } else {
            if (int rc = dense_f32(h, (const float*)d.values_fea + r0 * M, (const float*)d.B, (float*)XW + r0 * P, (int)(r1 - r0),
                                   (int)M, (int)P, 0)) return rc;
        }
        if (c == K - 1) CU(cudaEventRecord(h->ev[1], h->stream));
        for (int i = 0; i < K; i++) {
            if (launched[i] || need[i] > c) continue;
            launched[i] = true;
            const size_t a0 = (size_t)i * rows_per, a1 = a0 + rows_per < N ? a0 + rows_per : N;
            if (int rc = spmm_f32(h, d.rowPtr_adj + a0, d.columnIndex_adj, (const float*)d.values_adj, (const float*)XW,
                                  (float*)d.D + a0 * P, (int)(a1 - a0), (int)P, d.relu != 0, k0_host[i + 1] - k0_host[i], 0, 1, 0)) return rc;
            CU(cudaEventRecord(h->ev_pool[16 + i], h->stream));
            CU(cudaStreamWaitEvent(h->s_down, h->ev_pool[16 + i], 0));
            if (int rc = stage_out(h, a_d, (a1 - a0) * P * 4, a0 * P * 4, h->s_down)) return rc;
        }
    }
    CU(cudaEventRecord(h->ev[2], h->stream));
    CU(cudaEventRecord(ev_down, h->s_down));
    CU(cudaStreamWaitEvent(h->stream, ev_down, 0));
    h->overlapped_starts++;
    h->pipelined_starts++;
    return 0;
}

// Large float32 layers with host buffers: the copies run beside the kernels.  The feature operand and B go up on the
// handle's stream and the feature stage starts behind them; the adjacency goes up in row panels on a second stream,
// the aggregation runs panel by panel as the slices land, and each panel of D goes down on a third stream while the
// next panel's slices are still going up (host->device and device->host use different copy engines).
// Returns -100 when the layer does not qualify (the caller takes the serial path).
int start_overlapped(sgrace_handle* h, sgrace_layer_desc& d, uint64_t a_rpf, uint64_t a_cif, uint64_t a_vf, uint64_t a_rpa,
                     uint64_t a_cia, uint64_t a_va, uint64_t a_b, uint64_t a_d, long long nnz_fea, long long nnz_adj) {
    const size_t N = (size_t)d.N_adj, M = (size_t)d.M_fea, P = (size_t)d.P_w;
    if (!h->overlap || h->mode != SGRACE_MODE_F32_FAST || h->index_format != 0 || h->agg_first || h->accumulate ||
        h->peer_count > 0 || d.gemm_mode == 2 || d.gat_mode)
        return -100;
    const size_t d_bytes = N * P * 4;
    if (d_bytes < ((size_t)16 << 20) || (h->fused_small && d.N_adj <= h->fused_small)) return -100;
    const int* rp_host = (const int*)host_view(h, a_rpa);
    if (!rp_host || !host_view(h, a_cia) || !host_view(h, a_va) || !host_view(h, a_d)) return -100;
    if (rp_host[0] != 0) return -100;
    // panels of D of about 24 MB, row counts a multiple of 128 (the row-pointer slice of a panel stays 16-byte aligned)
    int K = (int)((d_bytes + ((size_t)24 << 20) - 1) / ((size_t)24 << 20));
    if (K < 2) K = 2;
    if (K > 16) K = 16;
    size_t rows_per = ((N + K - 1) / K + 127) & ~(size_t)127;
    K = (int)((N + rows_per - 1) / rows_per);
    if (K < 2) return -100;
    if (!h->s_up) CU(cudaStreamCreateWithFlags(&h->s_up, cudaStreamNonBlocking));
    if (!h->s_down) CU(cudaStreamCreateWithFlags(&h->s_down, cudaStreamNonBlocking));
    while ((int)h->ev_pool.size() < 2 * K + 3) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->ev_pool.push_back(e);
    }
    cudaEvent_t ev_prev = h->ev_pool[2 * K], ev_fup = h->ev_pool[2 * K + 1], ev_down = h->ev_pool[2 * K + 2];
    // whatever was queued on the handle's stream before this start comes first on the side streams too
    CU(cudaEventRecord(ev_prev, h->stream));
    CU(cudaStreamWaitEvent(h->s_up, ev_prev, 0));
    CU(cudaStreamWaitEvent(h->s_down, ev_prev, 0));
    // feature operand + B on the handle's stream, then the feature stage
    if (d.gemm_mode == 0) {
        if (int rc = stage_in(h, a_rpf, (N + 1) * 4)) return rc;
        if (int rc = stage_in(h, a_cif, (size_t)nnz_fea * 4)) return rc;
        if (int rc = stage_in(h, a_vf, (size_t)nnz_fea * 4)) return rc;
    } else {
        if (int rc = stage_in(h, a_vf, N * M * 4)) return rc;
    }
    if (int rc = stage_in(h, a_b, M * P * 4)) return rc;
    CU(cudaEventRecord(ev_fup, h->stream));
    // adjacency: the row pointers at once, then (column, value) slices panel by panel, behind the feature upload so
    // that the feature stage is not delayed
    CU(cudaStreamWaitEvent(h->s_up, ev_fup, 0));
    if (int rc = stage_in(h, a_rpa, (N + 1) * 4, 0, h->s_up)) return rc;
    for (int i = 0; i < K; i++) {
        const size_t r0 = (size_t)i * rows_per, r1 = r0 + rows_per < N ? r0 + rows_per : N;
        const long long k0 = rp_host[r0], k1 = rp_host[r1];
        if (k0 < 0 || k1 < k0 || k1 > nnz_adj) return fail(h, SGRACE_EBOUNDS, "adjacency row pointers not monotonic around row %zu", r0);
        if (int rc = stage_in(h, a_cia, (size_t)(k1 - k0) * 4, (size_t)k0 * 4, h->s_up)) return rc;
        if (int rc = stage_in(h, a_va, (size_t)(k1 - k0) * 4, (size_t)k0 * 4, h->s_up)) return rc;
        CU(cudaEventRecord(h->ev_pool[i], h->s_up));
    }
    void* XW = d.XW;
    if (!XW) {
        if (int rc = ensure(h, h->xw, 4 * N * P + 16)) return rc;
        XW = h->xw.p;
    }
    h->ev_valid = false;
    CU(cudaEventRecord(h->ev[0], h->stream));
    if (int rc = run_fea(h, &d, d.rowPtr_fea, XW)) return rc;
    CU(cudaEventRecord(h->ev[1], h->stream));
    for (int i = 0; i < K; i++) {
        const size_t r0 = (size_t)i * rows_per, r1 = r0 + rows_per < N ? r0 + rows_per : N;
        const long long k0 = rp_host[r0], k1 = rp_host[r1];
        CU(cudaStreamWaitEvent(h->stream, h->ev_pool[i], 0));
        if (int rc = spmm_f32(h, d.rowPtr_adj + r0, d.columnIndex_adj, (const float*)d.values_adj, (const float*)XW,
                              (float*)d.D + r0 * P, (int)(r1 - r0), (int)P, d.relu != 0, k1 - k0, 0, 1, 0)) return rc;
        CU(cudaEventRecord(h->ev_pool[K + i], h->stream));
        CU(cudaStreamWaitEvent(h->s_down, h->ev_pool[K + i], 0));
        if (int rc = stage_out(h, a_d, (r1 - r0) * P * 4, r0 * P * 4, h->s_down)) return rc;
    }
    CU(cudaEventRecord(h->ev[2], h->stream));
    CU(cudaEventRecord(ev_down, h->s_down));
    CU(cudaStreamWaitEvent(h->stream, ev_down, 0));     // the handle's stream ends when the last panel of D is on the host
    h->overlapped_starts++;
    return 0;
}

int validate_csr_host(sgrace_handle* h, const int* rp, const int* ci, int n, int ncols, const char* what) {
    for (int i = 0; i < n; i++)
        if (rp[i + 1] < rp[i]) return fail(h, SGRACE_EBOUNDS, "%s: rowPtr not monotonic at row %d", what, i);
    for (long long k = rp[0]; k < rp[n]; k++)
        if (ci[k] < 0 || ci[k] >= ncols)
            return fail(h, SGRACE_EBOUNDS, "%s: column index %d out of range [0,%d) at %lld", what, ci[k], ncols, k);
    return 0;
}

}  // namespace

// ====================================================================================
// exported functions
// ====================================================================================
extern "C" {

const char* sgrace_version(void) { return "sgrace_b200 0.1 (sm_100a)"; }

int sgrace_create(int device, sgrace_handle** out) {
    if (!out) return SGRACE_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return SGRACE_ECUDA;
    sgrace_handle* h = new (std::nothrow) sgrace_handle();
    if (!h) return SGRACE_ENOMEM;
    memset(h->regs, 0, sizeof(h->regs));
    h->err[0] = 0;
    h->device = device;
    h->regs[0] = 0x4;   // AP_IDLE
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return SGRACE_ECUDA;
    }
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    for (int i = 0; i < 5; i++) cudaEventCreate(&h->ev[i]);
    load_tune(h);
    *out = h;
    return SGRACE_OK;
}

int sgrace_destroy(sgrace_handle* h) {
    if (!h) return SGRACE_EINVAL;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (auto& kv : h->buffers) {
        cudaFree(kv.second.dev);
        cudaFreeHost(kv.second.host);
    }
    for (auto& pl : h->plans) { if (pl.panels) cudaFree(pl.panels); if (pl.info) cudaFree(pl.info); }
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->s_up) { cudaStreamSynchronize(h->s_up); cudaStreamDestroy(h->s_up); }
    if (h->s_down) { cudaStreamSynchronize(h->s_down); cudaStreamDestroy(h->s_down); }
    if (h->pipe_host) cudaFreeHost(h->pipe_host);
    if (h->pipe_dev.p) cudaFree(h->pipe_dev.p);
    Scratch* all[] = {&h->plan_a, &h->plan_b, &h->plan_c, &h->plan_d, &h->plan_flag, &h->plan_starts, &h->plan_tmp, &h->wrm, &h->ax, &h->long_partial, &h->long_done, &h->xw, &h->wq, &h->s1, &h->s2, &h->rp_fea, &h->rp_adj, &h->lists, &h->counters, &h->prep_keys, &h->prep_ids, &h->prep_misc, &h->prep_tmp};
    for (Scratch* s : all) if (s->p) cudaFree(s->p);
    if (h->max_fea_dev) cudaFree(h->max_fea_dev);
    for (int i = 0; i < 5; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return SGRACE_OK;
}

const char* sgrace_last_error(sgrace_handle* h) { return h ? h->err : "null handle"; }

int sgrace_alloc(sgrace_handle* h, size_t bytes, void** host_ptr, uint64_t* device_addr) {
    if (!h || !host_ptr || !device_addr) return SGRACE_EINVAL;
    CU(cudaSetDevice(h->device));
    Buffer b;
    b.bytes = bytes ? bytes : 1;
    size_t padded = (b.bytes + 255) & ~(size_t)255;
    CU(cudaMalloc(&b.dev, padded));
    cudaError_t e = cudaHostAlloc(&b.host, padded, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaFree(b.dev);
        return fail(h, SGRACE_ENOMEM, "cudaHostAlloc(%zu) failed: %s", padded, cudaGetErrorString(e));
    }
    memset(b.host, 0, padded);
    CU(cudaMemsetAsync(b.dev, 0, padded, h->stream));
    h->buffers[(uint64_t)(uintptr_t)b.dev] = b;
    *host_ptr = b.host;
    *device_addr = (uint64_t)(uintptr_t)b.dev;
    return SGRACE_OK;
}

int sgrace_free(sgrace_handle* h, uint64_t device_addr) {
    if (!h) return SGRACE_EINVAL;
    auto it = h->buffers.find(device_addr);
    if (it == h->buffers.end()) return fail(h, SGRACE_EINVAL, "sgrace_free: unknown buffer 0x%llx",
                                            (unsigned long long)device_addr);
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(it->second.dev);
    cudaFreeHost(it->second.host);
    h->buffers.erase(it);
    return SGRACE_OK;
}

int sgrace_sync_to_device(sgrace_handle* h, uint64_t device_addr, size_t bytes) {
    if (!h) return SGRACE_EINVAL;
    return stage_in(h, device_addr, bytes);
}
int sgrace_sync_from_device(sgrace_handle* h, uint64_t device_addr, size_t bytes) {
    if (!h) return SGRACE_EINVAL;
    if (int rc = stage_out(h, device_addr, bytes)) return rc;
    CU(cudaStreamSynchronize(h->stream));
    return SGRACE_OK;
}

int sgrace_reg_offset(const char* name, uint32_t* offset) {
    if (!name || !offset) return SGRACE_EINVAL;
    for (const RegName& r : kRegNames)
        if (strcmp(r.name, name) == 0) { *offset = r.off; return SGRACE_OK; }
    return SGRACE_EINVAL;
}

int sgrace_write_reg(sgrace_handle* h, uint32_t offset, uint32_t value) {
    if (!h) return SGRACE_EINVAL;
    if (offset % 4 || offset >= SGRACE_REG_FILE_BYTES) return fail(h, SGRACE_EINVAL, "bad register offset 0x%x", offset);
    if (offset == SGRACE_REG_CTRL) {
        if (value & 1u) return sgrace_start(h);   // CTRL.AP_START = 1
        return SGRACE_OK;
    }
    if (offset == SGRACE_REG_MAX_FEA) return SGRACE_OK;   // read-only
    h->regs[offset / 4] = value;
    return SGRACE_OK;
}

int sgrace_write_reg64(sgrace_handle* h, uint32_t offset, uint64_t value) {
    if (int rc = sgrace_write_reg(h, offset, (uint32_t)(value & 0xffffffffu))) return rc;
    return sgrace_write_reg(h, offset + 4, (uint32_t)(value >> 32));
}

int sgrace_read_reg(sgrace_handle* h, uint32_t offset, uint32_t* value) {
    if (!h || !value) return SGRACE_EINVAL;
    if (offset % 4 || offset >= SGRACE_REG_FILE_BYTES) return fail(h, SGRACE_EINVAL, "bad register offset 0x%x", offset);
    if (offset == SGRACE_REG_CTRL) {
        int done = 0;
        sgrace_done(h, &done);
        // AP_DONE (bit1) and AP_READY (bit3) when finished, AP_IDLE (bit2) when nothing runs
        *value = done ? 0xEu : 0x0u;
        return SGRACE_OK;
    }
    *value = h->regs[offset / 4];
    return SGRACE_OK;
}

int sgrace_set_option(sgrace_handle* h, int key, int64_t v) {
    if (!h) return SGRACE_EINVAL;
    switch (key) {
        case SGRACE_OPT_MODE:
            if (v < 0 || v > SGRACE_MODE_FULL) return fail(h, SGRACE_EINVAL, "bad mode %lld", (long long)v);
            h->mode = (int)v; break;
        case SGRACE_OPT_SPMM_BLOCK: if (v < 1) return fail(h, SGRACE_EINVAL, "spmm_block < 1"); h->spmm_block = (int)v; break;
        case SGRACE_OPT_LAT_FEA: if (v < 0 || v > 8) return fail(h, SGRACE_EINVAL, "lat_fea not in 0..8"); h->lat_fea = (int)v; break;
        case SGRACE_OPT_LAT_ADJ: if (v < 0 || v > 8) return fail(h, SGRACE_EINVAL, "lat_adj not in 0..8"); h->lat_adj = (int)v; break;
        case SGRACE_OPT_FEA_THREADS: if (v != 1 && v != 2 && v != 4) return fail(h, SGRACE_EINVAL, "fea_threads must be 1,2,4"); h->fea_threads = (int)v; break;
        case SGRACE_OPT_ADJ_THREADS: if (v != 1 && v != 2 && v != 4) return fail(h, SGRACE_EINVAL, "adj_threads must be 1,2,4"); h->adj_threads = (int)v; break;
        case SGRACE_OPT_USE_SBLOCKS: h->use_sblocks = v != 0; break;
        case SGRACE_OPT_INDEX_FORMAT: if (v != 0 && v != 1) return fail(h, SGRACE_EINVAL, "index_format must be 0/1"); h->index_format = (int)v; break;
        case SGRACE_OPT_QBITS: if (v != 0 && v != 1 && v != 2 && v != 4 && v != 8) return fail(h, SGRACE_EINVAL, "qbits must be 0,1,2,4,8"); h->qbits = (int)v; break;
        case SGRACE_OPT_STAGING: h->staging = v != 0; break;
        case SGRACE_OPT_LONG_ROW: if (v < 1) return fail(h, SGRACE_EINVAL, "long_row < 1"); h->long_row = (int)v; break;
        case SGRACE_OPT_LEAKY_ALPHA_BITS: { uint32_t b = (uint32_t)v; memcpy(&h->leaky_alpha, &b, 4); break; }
        case SGRACE_OPT_VALIDATE: h->validate = v != 0; break;
        case SGRACE_OPT_DENSE_TC: h->dense_tc = v != 0; break;
        case SGRACE_OPT_STREAM_KERNEL: h->stream_kernel = v != 0; break;
        case SGRACE_OPT_AGG_FIRST: h->agg_first = v != 0; break;
        case SGRACE_OPT_ACCUMULATE: h->accumulate = v != 0; break;
        case SGRACE_OPT_FUSED_SMALL: if (v < 0) return fail(h, SGRACE_EINVAL, "fused_small < 0"); h->fused_small = (int)v; break;
        case SGRACE_OPT_ROW_OFFSET: if (v < 0) return fail(h, SGRACE_EINVAL, "row_offset < 0"); h->row_offset = (int)v; break;
        case SGRACE_OPT_ADJ_PLAN: if (v < 0 || v > 2) return fail(h, SGRACE_EINVAL, "adj_plan must be 0,1,2"); h->adj_plan = (int)v; break;
        case SGRACE_OPT_OVERLAP: h->overlap = v != 0; break;
        case SGRACE_OPT_PUSH_CTAS: if (v < 0 || v > 65535) return fail(h, SGRACE_EINVAL, "push_ctas out of range"); h->push_ctas = (int)v; break;
        default: return fail(h, SGRACE_EINVAL, "unknown option %d", key);
    }
    return SGRACE_OK;
}

int sgrace_get_option(sgrace_handle* h, int key, int64_t* v) {
    if (!h || !v) return SGRACE_EINVAL;
    switch (key) {
        case SGRACE_OPT_MODE: *v = h->mode; break;
        case SGRACE_OPT_SPMM_BLOCK: *v = h->spmm_block; break;
        case SGRACE_OPT_LAT_FEA: *v = h->lat_fea; break;
        case SGRACE_OPT_LAT_ADJ: *v = h->lat_adj; break;
        case SGRACE_OPT_FEA_THREADS: *v = h->fea_threads; break;
        case SGRACE_OPT_ADJ_THREADS: *v = h->adj_threads; break;
        case SGRACE_OPT_USE_SBLOCKS: *v = h->use_sblocks; break;
        case SGRACE_OPT_INDEX_FORMAT: *v = h->index_format; break;
        case SGRACE_OPT_QBITS: *v = h->qbits; break;
        case SGRACE_OPT_STAGING: *v = h->staging; break;
        case SGRACE_OPT_LONG_ROW: *v = h->long_row; break;
        case SGRACE_OPT_LEAKY_ALPHA_BITS: { uint32_t b; memcpy(&b, &h->leaky_alpha, 4); *v = b; break; }
        case SGRACE_OPT_VALIDATE: *v = h->validate; break;
        case SGRACE_OPT_DENSE_TC: *v = h->dense_tc; break;
        case SGRACE_OPT_STREAM_KERNEL: *v = h->stream_kernel; break;
        case SGRACE_OPT_AGG_FIRST: *v = h->agg_first; break;
        case SGRACE_OPT_ACCUMULATE: *v = h->accumulate; break;
        case SGRACE_OPT_FUSED_SMALL: *v = h->fused_small; break;
        case SGRACE_OPT_ROW_OFFSET: *v = h->row_offset; break;
        case SGRACE_OPT_ADJ_PLAN: *v = h->adj_plan; break;
        case SGRACE_OPT_OVERLAP: *v = h->overlap; break;
        case SGRACE_OPT_PUSH_CTAS: *v = h->push_ctas; break;
        case SGRACE_OPT_OVERLAPPED_STARTS: *v = (int64_t)h->overlapped_starts; break;
        case SGRACE_OPT_PIPELINED_STARTS: *v = (int64_t)h->pipelined_starts; break;
        case SGRACE_OPT_PANEL_LAUNCHES: *v = (int64_t)h->panel_launches; break;
        case SGRACE_OPT_PLAN_BUILDS: *v = (int64_t)h->plan_builds; break;
        default: return fail(h, SGRACE_EINVAL, "unknown option %d", key);
    }
    return SGRACE_OK;
}

int sgrace_set_stream(sgrace_handle* h, void* s) {
    if (!h) return SGRACE_EINVAL;
    CU(cudaStreamSynchronize(h->stream));
    if (s) {
        if (h->own_stream) cudaStreamDestroy(h->stream);
        h->stream = (cudaStream_t)s;
        h->own_stream = false;
    } else if (!h->own_stream) {
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    return SGRACE_OK;
}

int sgrace_layer_run(sgrace_handle* h, const sgrace_layer_desc* d) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    return layer_run_impl(h, d, false);
}

int sgrace_fea_run(sgrace_handle* h, const sgrace_layer_desc* d, void* XW_out) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (int rc = check_desc(h, d)) return rc;
    if (!XW_out) return fail(h, SGRACE_EINVAL, "XW_out is null");
    const int *rp_fea, *rp_adj;
    if (int rc = resolve_rowptrs(h, d, &rp_fea, &rp_adj, true, false)) return rc;
    return run_fea(h, d, rp_fea, XW_out);
}

int sgrace_adj_run(sgrace_handle* h, const sgrace_layer_desc* d, const void* XW_in, int32_t xw_rows) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (int rc = check_desc(h, d)) return rc;
    if (!XW_in) return fail(h, SGRACE_EINVAL, "XW_in is null");
    const int *rp_fea, *rp_adj;
    if (int rc = resolve_rowptrs(h, d, &rp_fea, &rp_adj, false, true)) return rc;
    return run_adj(h, d, rp_adj, XW_in, xw_rows);
}

int sgrace_dense_run(sgrace_handle* h, const void* X, const void* B, void* out, int32_t N, int32_t M, int32_t P, int32_t relu) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (h->mode != SGRACE_MODE_F32_FAST) return fail(h, SGRACE_EUNSUPPORTED, "sgrace_dense_run is a float32 fast-path entry");
    if (!X || !B || !out || N < 0 || M <= 0 || P <= 0) return fail(h, SGRACE_EINVAL, "bad argument");
    if (N == 0) return SGRACE_OK;
    if (int rc = ensure(h, h->wrm, sizeof(float) * (size_t)M * P)) return rc;
    if (int rc = transpose_b<float>(h, B, h->wrm.p, M, P)) return rc;
    return dense_f32(h, (const float*)X, (const float*)B, (float*)out, N, M, P, relu != 0);
}

int sgrace_peer_alloc(sgrace_handle* h, size_t bytes, uint64_t* device_addr, unsigned char handle_out[64]) {
    if (!h || !device_addr || !handle_out) return SGRACE_EINVAL;
    CU(cudaSetDevice(h->device));
    void* p = nullptr;
    CU(cudaMalloc(&p, bytes ? bytes : 1));
    cudaIpcMemHandle_t ipc;
    cudaError_t e = cudaIpcGetMemHandle(&ipc, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(h, SGRACE_ECUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
    static_assert(sizeof(ipc) == 64, "IPC handle size");
    memcpy(handle_out, &ipc, 64);
    h->peer_allocs[(uint64_t)(uintptr_t)p] = bytes;
    *device_addr = (uint64_t)(uintptr_t)p;
    return SGRACE_OK;
}

int sgrace_peer_open(sgrace_handle* h, const unsigned char handle_in[64], uint64_t* device_addr) {
    if (!h || !handle_in || !device_addr) return SGRACE_EINVAL;
    CU(cudaSetDevice(h->device));
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, handle_in, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->peer_opened.push_back((uint64_t)(uintptr_t)p);
    *device_addr = (uint64_t)(uintptr_t)p;
    return SGRACE_OK;
}

// Teardown in two steps with a barrier between them on the caller's side: every rank first closes the mappings it
// imported (sgrace_peer_close), then -- once all ranks have done so -- frees what it exported (sgrace_peer_release).
// Freeing exported memory that a peer still has mapped is undefined behaviour.
int sgrace_peer_close(sgrace_handle* h) {
    if (!h) return SGRACE_EINVAL;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    for (uint64_t a : h->peer_opened) cudaIpcCloseMemHandle((void*)(uintptr_t)a);
    h->peer_opened.clear();
    return SGRACE_OK;
}

int sgrace_peer_release(sgrace_handle* h) {
    if (!h) return SGRACE_EINVAL;
    if (int rc = sgrace_peer_close(h)) return rc;
    for (auto& kv : h->peer_allocs) cudaFree((void*)(uintptr_t)kv.first);
    h->peer_allocs.clear();
    return SGRACE_OK;
}

int sgrace_peer_copy(sgrace_handle* h, uint64_t dst, uint64_t src, size_t bytes) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (bytes == 0) return SGRACE_OK;
    if (!dst || !src) return fail(h, SGRACE_EINVAL, "peer copy: null address");
    CU(cudaMemcpyAsync((void*)(uintptr_t)dst, (const void*)(uintptr_t)src, bytes, cudaMemcpyDeviceToDevice, h->stream));
    return SGRACE_OK;
}

// Flag protocol of the copy-engine exchange: a sender bumps a 32-bit epoch word in the receiver's peer-visible
// memory with a stream memory operation ordered after its copies; the receiver's stream waits until the word has
// reached the epoch (CU_STREAM_WAIT_VALUE_GEQ compares (int32)(*addr - value) >= 0, so it is wrap-safe and a sender
// that has already posted a later epoch never strands a wait).
extern "C++" {
template <typename Fn>
static Fn driver_entry(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        return reinterpret_cast<Fn>(p);
    return nullptr;
}
}

int sgrace_peer_signal(sgrace_handle* h, uint64_t flag_addr, uint32_t value) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (!flag_addr || (flag_addr & 3)) return fail(h, SGRACE_EINVAL, "peer signal: bad flag address");
    typedef CUresult (*WriteValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    static WriteValueFn fn = driver_entry<WriteValueFn>("cuStreamWriteValue32");
    if (!fn) return fail(h, SGRACE_EUNSUPPORTED, "cuStreamWriteValue32 is not available in this driver");
    const CUresult r = fn((CUstream)h->stream, (CUdeviceptr)flag_addr, (cuuint32_t)value, CU_STREAM_WRITE_VALUE_DEFAULT);
    if (r != CUDA_SUCCESS) return fail(h, SGRACE_ECUDA, "cuStreamWriteValue32 failed (%d)", (int)r);
    return SGRACE_OK;
}

int sgrace_wait_flag(sgrace_handle* h, uint64_t flag_addr, uint32_t value) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (!flag_addr || (flag_addr & 3)) return fail(h, SGRACE_EINVAL, "wait flag: bad flag address");
    typedef CUresult (*WaitValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    static WaitValueFn fn = driver_entry<WaitValueFn>("cuStreamWaitValue32");
    if (!fn) return fail(h, SGRACE_EUNSUPPORTED, "cuStreamWaitValue32 is not available in this driver");
    const CUresult r = fn((CUstream)h->stream, (CUdeviceptr)flag_addr, (cuuint32_t)value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) return fail(h, SGRACE_ECUDA, "cuStreamWaitValue32 failed (%d)", (int)r);
    return SGRACE_OK;
}

int sgrace_halo_gather(sgrace_handle* h, const uint64_t* bases, int32_t n_peers, int32_t block_rows, const int32_t* rows,
                       int64_t n_rows, int32_t width, void* dst) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (!bases || n_peers < 1 || n_peers > MAX_PEERS || block_rows < 1 || width < 4 || width % 4 || n_rows < 0 || (n_rows && (!rows || !dst)))
        return fail(h, SGRACE_EINVAL, "bad halo-gather argument");
    if (n_rows == 0) return SGRACE_OK;
    PeerTable pt;
    memset(&pt, 0, sizeof(pt));
    pt.count = n_peers; pt.block = block_rows;
    for (int r = 0; r < n_peers; r++) pt.base[r] = (const char*)(uintptr_t)bases[r];
    const long long total = (long long)n_rows * (width / 4);
    long long grid = (total + 255) / 256;
    if (grid > (long long)h->num_sms * 8) grid = (long long)h->num_sms * 8;
    halo_gather_kernel<<<(int)grid, 256, 0, h->stream>>>(pt, rows, n_rows, width / 4, (float4*)dst);
    h->launches++;
    CU(cudaGetLastError());
    return SGRACE_OK;
}

int sgrace_halo_push(sgrace_handle* h, const void* local, int32_t width, int32_t n_dst, const uint64_t* rows_ptrs,
                     const int64_t* counts, const uint64_t* dst_ptrs) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (!local || width < 4 || width % 4 || n_dst < 0 || n_dst > 8 || (n_dst && (!rows_ptrs || !counts || !dst_ptrs)))
        return fail(h, SGRACE_EINVAL, "bad halo-push argument");
    PushTable pt;
    memset(&pt, 0, sizeof(pt));
    pt.count = n_dst;
    long long total_rows = 0;
    for (int d = 0; d < n_dst; d++) {
        if (counts[d] < 0) return fail(h, SGRACE_EINVAL, "negative halo-push count");
        pt.rows[d] = (const int*)(uintptr_t)rows_ptrs[d];
        pt.dst[d] = (float4*)(uintptr_t)dst_ptrs[d];
        pt.start[d] = total_rows;
        total_rows += counts[d];
    }
    pt.start[n_dst] = total_rows;
    pt.rot = n_dst > 0 ? h->device % n_dst : 0;
    if (total_rows == 0) return SGRACE_OK;
    const size_t smem = (size_t)2 * HALO_CHUNK * width * 4;
    if (smem > 200 * 1024) return fail(h, SGRACE_EUNSUPPORTED, "halo push: rows wider than 400 floats are not supported");
    CU(cudaFuncSetAttribute(halo_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (h->tune.live) load_tune(h);
    long long grid = total_rows / HALO_CHUNK + n_dst;
    const long long cap = h->tune.push_ctas > 0 ? h->tune.push_ctas : (h->push_ctas > 0 ? h->push_ctas : (long long)h->num_sms * 4);
    if (grid > cap) grid = cap;
    halo_push_kernel<<<(int)grid, 256, smem, h->stream>>>(pt, (const float4*)local, width / 4);
    h->launches++;
    CU(cudaGetLastError());
    return SGRACE_OK;
}

int sgrace_adj_run_peer(sgrace_handle* h, const sgrace_layer_desc* d, const uint64_t* bases, int32_t n_peers,
                        int32_t block_rows) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (int rc = check_desc(h, d)) return rc;
    if (!bases || n_peers < 1 || n_peers > MAX_PEERS || block_rows < 1) return fail(h, SGRACE_EINVAL, "bad peer table");
    if (h->mode != SGRACE_MODE_F32_FAST) return fail(h, SGRACE_EUNSUPPORTED, "peer gathers are a float32 fast-path feature");
    const int *rp_fea, *rp_adj;
    if (int rc = resolve_rowptrs(h, d, &rp_fea, &rp_adj, false, true)) return rc;
    for (int r = 0; r < MAX_PEERS; r++) h->peer_base[r] = r < n_peers ? (const char*)(uintptr_t)bases[r] : nullptr;
    h->peer_block = block_rows; h->peer_count = n_peers;
    const int rc = run_adj(h, d, rp_adj, (const void*)(uintptr_t)bases[0], n_peers * block_rows);
    h->peer_count = 0;
    return rc;
}

int sgrace_xty_run(sgrace_handle* h, const void* X, const void* Y, void* out, int32_t N, int32_t M, int32_t P) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (!X || !Y || !out || N < 0 || M <= 0 || P <= 0) return fail(h, SGRACE_EINVAL, "bad argument");
    const int tiles = ((M + 63) / 64) * ((P + 63) / 64);
    int chunks = h->num_sms * 2 / tiles;
    if (chunks < 1) chunks = 1;
    int rows_per = (N + chunks - 1) / chunks;
    rows_per = ((rows_per + 31) / 32) * 32;
    if (rows_per < 32) rows_per = 32;
    chunks = N > 0 ? (N + rows_per - 1) / rows_per : 1;
    if (int rc = ensure(h, h->long_partial, sizeof(float) * 4096 * (size_t)chunks * tiles)) return rc;
    xty_partial_kernel<<<dim3(chunks, tiles), 256, 0, h->stream>>>((const float*)X, (const float*)Y, (float*)h->long_partial.p, N, M,
                                                                    P, rows_per);
    xty_reduce_kernel<<<dim3(tiles, 16), 256, 0, h->stream>>>((const float*)h->long_partial.p, (float*)out, M, P, chunks, tiles);
    h->launches += 2;
    CU(cudaGetLastError());
    return SGRACE_OK;
}

int sgrace_start(sgrace_handle* h) {
    if (!h) return SGRACE_EINVAL;
    CU(cudaSetDevice(h->device));
    h->last_status = 0;
    // bias_count > 0 is the parameter-preload-only call of the open design: nothing to compute
    // (kernelMatrixmult_all.cpp:3876-3888)
    if ((int32_t)h->regs[SGRACE_REG_BIAS_COUNT / 4] > 0) { h->running = true; h->ev_valid = false;
        CU(cudaEventRecord(h->ev[3], h->stream)); return SGRACE_OK; }

    // layer_count > 1 / stream_mode != 0 (demo/emulation/config.py:6,17, sgrace.py:1862): several layers per start with
    // the activations kept inside the device.  Every shipped configuration programs 1 / 0, and the open sources do not
    // say where the closed design takes the next layer's weights and sizes from: refused rather than guessed.
    if ((int32_t)h->regs[SGRACE_REG_LAYER_COUNT / 4] > 1 || (h->regs[SGRACE_REG_STREAM_MODE / 4] & 1u))
        return fail(h, SGRACE_EUNSUPPORTED, "layer_count=%d stream_mode=%u: multi-layer streaming of the closed design is "
                    "not specified by the open sources (config.py:6,17); program layer_count=1, stream_mode=0",
                    (int)h->regs[SGRACE_REG_LAYER_COUNT / 4], h->regs[SGRACE_REG_STREAM_MODE / 4] & 1u);

    sgrace_layer_desc d;
    memset(&d, 0, sizeof(d));
    d.gemm_mode = (int32_t)h->regs[SGRACE_REG_GEMM_MODE / 4];
    d.relu = (int32_t)(h->regs[SGRACE_REG_RELU / 4] & 1u);
    d.gat_mode = (int32_t)(h->regs[SGRACE_REG_GAT_MODE / 4] & 1u);
    d.N_adj = (int32_t)h->regs[SGRACE_REG_N_ADJ / 4];
    d.M_adj = (int32_t)h->regs[SGRACE_REG_M_ADJ / 4];
    d.M_fea = (int32_t)h->regs[SGRACE_REG_M_FEA / 4];
    d.P_w = (int32_t)h->regs[SGRACE_REG_P_W / 4];
    d.nnz_fea = (int32_t)h->regs[SGRACE_REG_NNZ_FEA1 / 4];
    d.nnz_adj = (int32_t)h->regs[SGRACE_REG_NNZ_ADJ1 / 4];
    d.scale_fea = (int32_t)h->regs[SGRACE_REG_SCALE_FEA / 4];
    d.internal_quantization = (int32_t)h->regs[SGRACE_REG_QUANTIZED_MULTIPLIER / 4];
    d.qscale_fea = regf(h, SGRACE_REG_QSCALE_FEA);
    d.qscale_w = regf(h, SGRACE_REG_QSCALE_W);
    d.qscale_adj = regf(h, SGRACE_REG_QSCALE_ADJ);
    d.deq_factor = regf(h, SGRACE_REG_DEQ_FACTOR);
    const uint64_t a_rpf = reg64(h, SGRACE_REG_ROWPTR_FEA1), a_cif = reg64(h, SGRACE_REG_COLIDX_FEA1),
                   a_vf = reg64(h, SGRACE_REG_VALUES_FEA1), a_rpa = reg64(h, SGRACE_REG_ROWPTR_ADJ1),
                   a_cia = reg64(h, SGRACE_REG_COLIDX_ADJ1), a_va = reg64(h, SGRACE_REG_VALUES_ADJ1),
                   a_b = reg64(h, SGRACE_REG_B), a_d = reg64(h, SGRACE_REG_D1), a_e = reg64(h, SGRACE_REG_E1),
                   a_s = reg64(h, SGRACE_REG_S1), a_att = reg64(h, SGRACE_REG_ATE_M);
    d.rowPtr_fea = (const int32_t*)(uintptr_t)a_rpf;
    d.columnIndex_fea = (const int32_t*)(uintptr_t)a_cif;
    d.values_fea = (const void*)(uintptr_t)a_vf;
    d.rowPtr_adj = (const int32_t*)(uintptr_t)a_rpa;
    d.columnIndex_adj = (const int32_t*)(uintptr_t)a_cia;
    d.values_adj = (const void*)(uintptr_t)a_va;
    d.B = (const void*)(uintptr_t)a_b;
    d.D = (void*)(uintptr_t)a_d;
    const bool full = h->mode == SGRACE_MODE_FULL;
    const bool gat = full && d.gat_mode;
    d.attention = gat ? (const float*)(uintptr_t)a_att : nullptr;
    d.E = gat ? (float*)(uintptr_t)a_e : nullptr;
    d.S = gat ? (float*)(uintptr_t)a_s : nullptr;
    if (int rc = check_desc(h, &d)) return rc;

    const size_t esz = elt_bytes(h->mode);
    const size_t N = (size_t)d.N_adj, M = (size_t)d.M_fea, P = (size_t)d.P_w;
    if (d.gemm_mode == 2) {
        // backward launch: the sparse operand (M_adj rows) sits behind the *_fea registers, the dense one
        // (N_adj x M_adj) behind values_adj.  The driver does not rewrite the count registers for this launch
        // (sgrace.py:717-790): in the COO format the count is the one the forward left in nnz_adj1.
        const size_t R = (size_t)d.M_adj;
        long long nnz = h->index_format == 1 ? (long long)d.nnz_adj : -1;
        if (h->index_format == 0) {
            size_t off;
            const Buffer* b = find_buffer(h, a_rpf, &off);
            if (b && off + (R + 1) * 4 > b->bytes) return fail(h, SGRACE_EBOUNDS, "rowPtr buffer too small for M_adj=%zu", R);
            if (h->staging && b) nnz = ((const int*)((const char*)b->host + off))[R];
            else if (d.rowPtr_fea && R) {
                int v = 0;
                CU(cudaMemcpyAsync(&v, d.rowPtr_fea + R, 4, cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
                nnz = v;
            }
        }
        if (nnz < 0) return fail(h, SGRACE_EBOUNDS, "negative non-zero count");
        d.nnz_fea = (int32_t)nnz;
        CU(cudaEventRecord(h->ev[4], h->stream));
        if (h->staging) {
            if (int rc = stage_in(h, a_rpf, h->index_format == 0 ? (R + 1) * 4 : (size_t)nnz * 4)) return rc;
            if (int rc = stage_in(h, a_cif, (size_t)nnz * 4)) return rc;
            if (int rc = stage_in(h, a_vf, (size_t)nnz * esz)) return rc;
            if (int rc = stage_in(h, a_va, N * R * esz)) return rc;
            if (int rc = stage_in(h, a_b, M * P * esz)) return rc;
        }
        h->ev_valid = false;
        if (int rc = layer_run_impl(h, &d, true)) return rc;
        if (h->staging) if (int rc = stage_out(h, a_d, N * P * esz)) return rc;
        CU(cudaEventRecord(h->ev[3], h->stream));
        h->ev_valid = true;
        h->running = true;
        return SGRACE_OK;
    }
    // non-zero counts: registers (COO format) or the last row pointer read from the host mirror
    long long nnz_fea = d.nnz_fea, nnz_adj = d.nnz_adj;
    if (h->index_format == 0) {
        // true CSR: the count is the last row pointer -- read from the host mirror when there is
        // one (bounds-checked), else from the device
        auto last_ptr = [&](uint64_t addr, const int32_t* dptr, long long* out) -> int {
            size_t off;
            const Buffer* b = find_buffer(h, addr, &off);
            if (b && off + (N + 1) * 4 > b->bytes)
                return fail(h, SGRACE_EBOUNDS, "rowPtr buffer holds %zu bytes but N_adj=%zu needs %zu", b->bytes - off,
                            N, (N + 1) * 4);
            if (h->staging && b) { *out = ((const int*)((const char*)b->host + off))[N]; return 0; }
            if (*out <= 0 && dptr && N) {
                int v = 0;
                CU(cudaMemcpyAsync(&v, dptr + N, 4, cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
                *out = v;
            }
            return 0;
        };
        if (int rc = last_ptr(a_rpa, d.rowPtr_adj, &nnz_adj)) return rc;
        if (d.gemm_mode == 0) if (int rc = last_ptr(a_rpf, d.rowPtr_fea, &nnz_fea)) return rc;
        d.nnz_adj = (int32_t)nnz_adj;
        d.nnz_fea = (int32_t)nnz_fea;
    }
    if (nnz_adj < 0 || nnz_fea < 0) return fail(h, SGRACE_EBOUNDS, "negative non-zero count");

    if (h->validate && h->staging && h->index_format == 0) {
        auto span_ok = [&](uint64_t addr, size_t bytes) {
            size_t off;
            const Buffer* b = find_buffer(h, addr, &off);
            return b && off + bytes <= b->bytes;
        };
        if (span_ok(a_rpa, (N + 1) * 4) && span_ok(a_cia, (size_t)nnz_adj * 4))
            if (int rc = validate_csr_host(h, (const int*)host_view(h, a_rpa), (const int*)host_view(h, a_cia),
                                           (int)N, (int)N, "adjacency")) return rc;
        if (d.gemm_mode == 0 && span_ok(a_rpf, (N + 1) * 4) && span_ok(a_cif, (size_t)nnz_fea * 4))
            if (int rc = validate_csr_host(h, (const int*)host_view(h, a_rpf), (const int*)host_view(h, a_cif),
                                           (int)N, (int)M, "features")) return rc;
    }

    CU(cudaEventRecord(h->ev[4], h->stream));
    if (h->staging && !(full && h->qbits > 0)) {
        // the side-stream paths wait on an event from the host: not inside a stream capture
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        CU(cudaStreamIsCapturing(h->stream, &cap));
        int rc = -100;
        if (cap == cudaStreamCaptureStatusNone) {
            rc = start_pipelined(h, d, a_rpf, a_cif, a_vf, a_rpa, a_cia, a_va, a_b, a_d, nnz_fea, nnz_adj);
            if (rc == -100) rc = start_overlapped(h, d, a_rpf, a_cif, a_vf, a_rpa, a_cia, a_va, a_b, a_d, nnz_fea, nnz_adj);
        }
        if (rc != -100) {
            if (rc) {
                // an error after copies were queued on the side streams: let them drain before anything else touches the buffers
                if (h->s_up) cudaStreamSynchronize(h->s_up);
                if (h->s_down) cudaStreamSynchronize(h->s_down);
                return rc;
            }
            const uint64_t a_prof = reg64(h, SGRACE_REG_PROFILING);
            size_t off;
            const Buffer* pb = find_buffer(h, a_prof, &off);
            if (pb && off + 15 * 8 <= pb->bytes) memset((char*)pb->host + off, 0, 15 * 8);
            CU(cudaEventRecord(h->ev[3], h->stream));
            h->ev_valid = true;
            h->running = true;
            return SGRACE_OK;
        }
    }
    if (h->staging) {
        // host mirror -> device, only the ranges this layer reads
        const size_t rp_bytes_a = h->index_format == 0 ? (N + 1) * 4 : (size_t)nnz_adj * 4;
        if (int rc = stage_in(h, a_rpa, rp_bytes_a)) return rc;
        if (int rc = stage_in(h, a_cia, (size_t)nnz_adj * 4)) return rc;
        if (int rc = stage_in(h, a_va, (size_t)nnz_adj * esz)) return rc;
        if (d.gemm_mode == 0) {
            const size_t rp_bytes_f = h->index_format == 0 ? (N + 1) * 4 : (size_t)nnz_fea * 4;
            if (int rc = stage_in(h, a_rpf, rp_bytes_f)) return rc;
            if (int rc = stage_in(h, a_cif, (size_t)nnz_fea * 4)) return rc;
            if (int rc = stage_in(h, a_vf, (size_t)nnz_fea * esz)) return rc;
        } else {
            if (int rc = stage_in(h, a_vf, N * M * esz)) return rc;
        }
        if (int rc = stage_in(h, a_b, M * P * esz)) return rc;
        if (gat) if (int rc = stage_in(h, a_att, 2 * P * 4)) return rc;
    }

    h->ev_valid = false;
    if (int rc = layer_run_impl(h, &d, true)) return rc;

    if (full && h->qbits > 0 && h->max_fea_dev) {
        // max |X_q W_q| as a 16-fractional-bit integer (sgrace.py:506-520 divides by 2^frac_bits_o)
        int m = 0;
        CU(cudaMemcpyAsync(&m, h->max_fea_dev, 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        const double den = (h->qbits == 1 ? 2.0 : (double)(1 << (h->qbits - 1)));
        double v = (double)m / (den * den) * 65536.0;
        h->regs[SGRACE_REG_MAX_FEA / 4] = v > 2147483647.0 ? 2147483647u : (uint32_t)v;
    }
    if (h->staging) {
        if (int rc = stage_out(h, a_d, N * P * esz)) return rc;
        if (gat) {
            if (int rc = stage_out(h, a_e, (size_t)nnz_adj * 4)) return rc;
            if (int rc = stage_out(h, a_s, (size_t)nnz_adj * 4)) return rc;
        }
        // the open design's FIFO counters read 0 as built (mmult-master.ipynb cells 39-40)
        const uint64_t a_prof = reg64(h, SGRACE_REG_PROFILING);
        size_t off;
        const Buffer* pb = find_buffer(h, a_prof, &off);
        if (pb && off + 15 * 8 <= pb->bytes) memset((char*)pb->host + off, 0, 15 * 8);
    }
    CU(cudaEventRecord(h->ev[3], h->stream));
    h->ev_valid = true;
    h->running = true;
    return SGRACE_OK;
}

int sgrace_done(sgrace_handle* h, int* done) {
    if (!h || !done) return SGRACE_EINVAL;
    if (!h->running) { *done = 0; return SGRACE_OK; }
    cudaError_t e = cudaEventQuery(h->ev[3]);
    if (e == cudaSuccess) { *done = 1; return SGRACE_OK; }
    if (e == cudaErrorNotReady) { *done = 0; return SGRACE_OK; }
    *done = 1;
    return fail(h, SGRACE_ECUDA, "layer failed: %s", cudaGetErrorString(e));
}

int sgrace_wait(sgrace_handle* h) {
    if (!h) return SGRACE_EINVAL;
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return h->last_status;
}

int sgrace_stage_times(sgrace_handle* h, float* fea_ms, float* adj_ms, float* total_ms) {
    if (!h) return SGRACE_EINVAL;
    if (!h->ev_valid) return fail(h, SGRACE_EINVAL, "no completed layer to time");
    CU(cudaEventSynchronize(h->ev[3]));
    float a = 0, b = 0, c = 0;
    CU(cudaEventElapsedTime(&a, h->ev[0], h->ev[1]));
    CU(cudaEventElapsedTime(&b, h->ev[1], h->ev[2]));
    CU(cudaEventElapsedTime(&c, h->ev[4], h->ev[3]));
    if (fea_ms) *fea_ms = a;
    if (adj_ms) *adj_ms = b;
    if (total_ms) *total_ms = c;
    return SGRACE_OK;
}

int sgrace_sym_norm(sgrace_handle* h, const int32_t* row, const int32_t* col, const float* weight, int64_t nnz,
                    int32_t n_nodes, float fill, int64_t capacity, int32_t* out_row, int32_t* out_col, float* out_val,
                    int32_t* out_rowptr, int64_t* out_nnz) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (nnz < 0 || n_nodes < 0 || (nnz && (!row || !col)) || !out_row || !out_col || !out_val || !out_nnz)
        return fail(h, SGRACE_EINVAL, "sym_norm: bad argument");
    const long long total = (long long)nnz + n_nodes;
    if (total >= (1ll << 31)) return fail(h, SGRACE_EUNSUPPORTED, "sym_norm: more than 2^31 entries");
    *out_nnz = 0;
    if (total == 0) return SGRACE_OK;
    const int n = n_nodes;
    using namespace prep;
    // scratch: keys in|out, ids in|out, misc = loop_edge[n] | counters[2] | dis[n] | rowptr[n+1] | w[total]
    if (int rc = ensure(h, h->prep_keys, 16 * (size_t)total)) return rc;
    if (int rc = ensure(h, h->prep_ids, 8 * (size_t)total)) return rc;
    const size_t misc_ints = (size_t)n + 2 + n + (n + 1) + total;
    if (int rc = ensure(h, h->prep_misc, 4 * misc_ints)) return rc;
    unsigned long long* k0 = (unsigned long long*)h->prep_keys.p; unsigned long long* k1 = k0 + total;
    int* i0 = (int*)h->prep_ids.p; int* i1 = i0 + total;
    int* loop_edge = (int*)h->prep_misc.p; int* counters = loop_edge + n;
    float* dis = (float*)(counters + 2); int* rowptr = (int*)(dis + n); float* w_sorted = (float*)(rowptr + n + 1);
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, i0, i1, (int)total, 0, 64, h->stream));
    if (int rc = ensure(h, h->prep_tmp, tmp_bytes)) return rc;
    CU(cudaMemsetAsync(loop_edge, 0xff, sizeof(int) * (size_t)n, h->stream));
    CU(cudaMemsetAsync(counters, 0, 2 * sizeof(int), h->stream));
    const int blocks = (int)((total + 255) / 256);
    sym_keys_kernel<<<blocks, 256, 0, h->stream>>>(row, col, nnz, n, k0, i0, loop_edge, counters);
    CU(cudaGetLastError());
    tmp_bytes = h->prep_tmp.bytes;
    CU(cub::DeviceRadixSort::SortPairs(h->prep_tmp.p, tmp_bytes, k0, k1, i0, i1, (int)total, 0, 64, h->stream));
    int host_counters[2] = {0, 0};
    CU(cudaMemcpyAsync(host_counters, counters, sizeof(host_counters), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (host_counters[1]) return fail(h, SGRACE_EBOUNDS, "sym_norm: %d edges with a node index outside [0, %d)", host_counters[1], n);
    const long long kept = total - host_counters[0];
    *out_nnz = kept;
    if (kept > capacity) return fail(h, SGRACE_EBOUNDS, "sym_norm: result has %lld entries, capacity %lld", kept, (long long)capacity);
    const int oblocks = (int)((kept + 255) / 256);
    sym_emit_kernel<<<oblocks, 256, 0, h->stream>>>(k1, i1, weight, loop_edge, fill, nnz, kept, out_row, out_col, w_sorted);
    int* rp = out_rowptr ? out_rowptr : rowptr;
    coo_rows_to_rowptr_kernel<<<(n + 1 + 255) / 256, 256, 0, h->stream>>>(out_row, (int)kept, n, rp);
    sym_deg_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(rp, w_sorted, n, dis);
    sym_norm_kernel<<<oblocks, 256, 0, h->stream>>>(out_row, out_col, w_sorted, dis, kept, out_val);
    CU(cudaGetLastError());
    h->launches += 5;
    return SGRACE_OK;
}

int sgrace_dense_to_csr(sgrace_handle* h, const float* X, int32_t n, int32_t m, int64_t capacity, int32_t* rowptr,
                        int32_t* col, float* val, int64_t* out_nnz) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (n < 0 || m < 0 || !rowptr || !out_nnz || (n > 0 && m > 0 && !X)) return fail(h, SGRACE_EINVAL, "dense_to_csr: bad argument");
    using namespace prep;
    *out_nnz = 0;
    CU(cudaMemsetAsync(rowptr, 0, sizeof(int) * ((size_t)n + 1), h->stream));
    if (n == 0) { CU(cudaStreamSynchronize(h->stream)); return SGRACE_OK; }
    const int blocks = (int)(((long long)n * 32 + 255) / 256);
    dense_count_kernel<<<blocks, 256, 0, h->stream>>>(X, n, m, rowptr);
    CU(cudaGetLastError());
    size_t tmp_bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, rowptr, rowptr, n + 1, h->stream));
    if (int rc = ensure(h, h->prep_tmp, tmp_bytes)) return rc;
    tmp_bytes = h->prep_tmp.bytes;
    CU(cub::DeviceScan::ExclusiveSum(h->prep_tmp.p, tmp_bytes, rowptr, rowptr, n + 1, h->stream));
    int total = 0;
    CU(cudaMemcpyAsync(&total, rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *out_nnz = total;
    h->launches += 2;
    if (total > capacity) return fail(h, SGRACE_EBOUNDS, "dense_to_csr: %d non-zeros, capacity %lld", total, (long long)capacity);
    if (total == 0) return SGRACE_OK;
    if (!col || !val) return fail(h, SGRACE_EINVAL, "dense_to_csr: col / val missing");
    dense_fill_kernel<<<blocks, 256, 0, h->stream>>>(X, n, m, rowptr, capacity, col, val);
    CU(cudaGetLastError());
    h->launches++;
    return SGRACE_OK;
}

int sgrace_prune_adjacency(sgrace_handle* h, const int32_t* rowptr, const int32_t* col, const float* val, int32_t n, int64_t nnz,
                           float qscale_adj, int32_t qbits, int32_t* out_rowptr, int32_t* out_col, float* out_val, int32_t* kept,
                           int64_t* out_nnz) {
    if (!h) return SGRACE_EINVAL;
    h->last_status = 0;
    CU(cudaSetDevice(h->device));
    if (n < 0 || nnz < 0 || nnz >= 0x7fffffffLL || !rowptr || !out_rowptr || !out_nnz || (nnz > 0 && (!col || !val || !out_col || !out_val)))
        return fail(h, SGRACE_EINVAL, "prune_adjacency: bad argument");
    if (qbits < 1 || qbits > 8 || !(qscale_adj > 0.f)) return fail(h, SGRACE_EINVAL, "prune_adjacency: qbits in 1..8 and a positive scale");
    using namespace prep;
    *out_nnz = 0;
    const size_t pos_bytes = sizeof(int) * ((size_t)nnz + 1);
    if (int rc = ensure(h, h->prep_ids, pos_bytes)) return rc;
    int* pos = (int*)h->prep_ids.p;
    const int blocks = (int)((nnz + 1 + 255) / 256);
    prune_flag_kernel<<<blocks, 256, 0, h->stream>>>(val, nnz, qscale_adj, 0, qbits, pos);
    CU(cudaGetLastError());
    size_t tmp_bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, pos, pos, (int)(nnz + 1), h->stream));
    if (int rc = ensure(h, h->prep_tmp, tmp_bytes)) return rc;
    tmp_bytes = h->prep_tmp.bytes;
    CU(cub::DeviceScan::ExclusiveSum(h->prep_tmp.p, tmp_bytes, pos, pos, (int)(nnz + 1), h->stream));
    if (nnz > 0) {
        prune_scatter_kernel<<<(int)((nnz + 255) / 256), 256, 0, h->stream>>>(col, val, pos, nnz, out_col, out_val, kept);
        CU(cudaGetLastError());
    }
    prune_rowptr_kernel<<<(n + 1 + 255) / 256, 256, 0, h->stream>>>(rowptr, pos, n, out_rowptr);
    CU(cudaGetLastError());
    int total = 0;
    CU(cudaMemcpyAsync(&total, pos + nnz, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *out_nnz = total;
    h->launches += 4;
    return SGRACE_OK;
}

int sgrace_launch_count(sgrace_handle* h, uint64_t* count) {
    if (!h || !count) return SGRACE_EINVAL;
    *count = h->launches;
    return SGRACE_OK;
}

}  // extern "C"
