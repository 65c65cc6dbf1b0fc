// sgrace_spmm_panel.cuh -- ADJ stage for block-diagonal adjacencies (batched graphs: cora_x1024, molecule batches)
// with the XW gathers served from SHARED MEMORY, float32:
//
//     D[r,:] = act( sum_k val[k] * XW[col[k],:] )
//
// The rows of one diagonal block only reference the XW rows of the same block.  A PANEL is a run of whole
// blocks whose XW rows (the "window") fit shared memory beside the CSR stage rings.  A CTA claims a panel,
// bulk-copies the window once (TMA, cp.async.bulk -> UBLKCP), and streams the panel's CSR arrays through
// rings of stages as spmm_stream_f32_kernel does; the consumers' gathers are LDS.128 from the window instead
// of L2 round trips.  This is the PIPO hand-off of the reference (the XW tile held on chip between loop_fea and
// loop_adj, kernelMatrixmult_all.cpp:3651-3713) at the granularity of a graph.  A column outside the window
// falls back to a global load and panels always partition the rows, so the result never depends on the plan --
// a stale or coarse plan costs speed, never correctness.
//
// Reference behaviour being replaced (not ported): loop_adj / dsp_kernel_wrapper_adj_* (kernelMatrixmult_all.cpp:
// 3339-3627, 1778-1957).  Accumulation is float FMA in CSR order within a row, lane l of a row group holding
// float4 chunk l of the output row as in spmm_stream_f32_kernel: bit-equal to it (tests/test_gpu_panel.py).
//
// CTA layout (one persistent CTA per SM, 32 * (G * (1 + ncw) + 1) threads):
//   * G pipelines of 1 PRODUCER warp + ncw CONSUMER warps.  A producer claims row tiles inside the current panel from
//     a shared-memory counter, samples 33 row pointers per tile (one tile ahead), packs the tile's pieces into stages
//     and streams each stage's rowPtr / columnIndex / values slices into its ring with cp.async.bulk; every panel
//     ends with a marker stage.
//   * the consumer warps of a pipeline share every stage: the stage's rows are cut into chunks of 32, handed to the
//     warps round-robin across stages.  A warp counting-sorts its 32 rows by length with ballots (one row per lane,
//     the order in a 32-byte shared array) and runs lock-step passes of 32/LPR rows of (nearly) equal length, so the
//     per-non-zero loop -- two LDS.32, one LDS.128 gather, four FFMA, two non-zeros per step with the next pair's
//     loads under the current pair's gathers -- carries no row bookkeeping and few idle slots (aligned 4-blocks with
//     rows fixed to groups fill 32 % of the issued slots on Cora-shape rows).
//   * a row longer than `hub_thresh` is taken by the whole warp at once: the row groups gather the XW rows of 32
//     non-zeros together and run the FMA chain group after group, passing the accumulator by shuffle -- the CSR
//     order of additions exactly, at a latency of one gather per 32 non-zeros instead of per 4.
//   * the WINDOW warp (one lane) claims panels from a global counter (one ahead), publishes the panel descriptor,
//     waits until every consumer warp has left the previous panel (`wfree`) and issues the next window copy onto
//     `bfull`.  The producers run one panel ahead of the window.
#pragma once
#include "sgrace_spmm_stream.cuh"

namespace sgrace {

enum { PANEL_END = 1 };

struct PanelParams {
    const int* rowptr;
    const int* col;
    const float* val;
    const float4* Bm;          // XW row-major (global), row stride LPR*NV float4
    float4* out;
    int nrows, relu, streaming_store;
    int bm_rows;               // rows of Bm
    int long_thresh;           // rows with more non-zeros go to the long-row kernel (global list)
    int hub_thresh;            // rows with more non-zeros (up to long_thresh) are taken by a whole warp
    int tile_rows;             // TR: rows per claimed tile, multiple of 32, <= 1024
    int stage_nnz;             // C: non-zeros a stage can hold (multiple of 4)
    int stages;                // S
    int groups;                // G
    int win_bytes;             // capacity of the window, bytes (multiple of 128)
    const int4* panels;        // {row_begin, row_end, win_base, win_rows}
    const int* npanels;        // device memory, written by the planner
    int* long_rows;
    int* long_count;
    int* panel_counter;
    int dbg;                   // measurement only (SGRACE_PANEL_DBG): 1 no window copies, 2 no passes, 4 no stores, 8 no sort/scan
};

struct PanelHeader {           // 32 bytes
    int row_begin;             // first row of the stage
    int nrows;                 // rows in the stage; < 0: no more work; 0 with PANEL_END: end-of-panel marker
    int kbase;                 // global index of the non-zero stored at col_s[0] / val_s[0] (multiple of 4)
    int roff;                  // rp_s[roff + i] is rowptr[row_begin + i]
    int flags;                 // PANEL_END | chunk_base << 8 (32-row chunks of this pipeline before this stage, mod ncw)
    int pseq;                  // panel sequence number within this CTA
    int win_base;              // first XW row held in the window
    int win_rows;
};

struct TrueTag { static constexpr bool value = true; };
struct FalseTag { static constexpr bool value = false; };

// bytes in front of the window: mbarriers, control words, panel descriptors, stage headers, per-warp sort arrays
__host__ __device__ inline int panel_ctrl_bytes(int groups, int stages) {
    const int gs = groups * stages;
    int off = 8 * (2 * gs + 2);            // full[G][S], empty[G][S], bfull, wfree
    off += 16;                             // pub_seq, tile_ctr[2], pad
    off += 32;                             // pdesc[2]
    off += 32 * gs;                        // headers
    off += 32 * 32;                        // order[32 warps][32]
    return (off + 127) & ~127;
}
inline size_t panel_smem_bytes(int groups, int stages, int tile_rows, int stage_nnz, int win_bytes) {
    return (size_t)panel_ctrl_bytes(groups, stages) + (size_t)win_bytes +
           (size_t)groups * stages * stream_stage_bytes(tile_rows, stage_nnz) + 128;
}

// blockDim.x = 32 * (G * (1 + ncw) + 1); the last warp is the window warp
template <int LPR, int NV, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
spmm_panel_f32_kernel(const PanelParams p) {
    constexpr int RPW = 32 / LPR;                 // row groups per warp
    constexpr int SLOTS = LPR < 4 ? LPR : 4;      // non-zeros a group takes per step of the hub chain
    constexpr uint32_t rowbytes = LPR * NV * 16;
    extern __shared__ __align__(128) unsigned char smem[];
    const int S = p.stages, TR = p.tile_rows, C = p.stage_nnz, G = p.groups;
    const int warp_all = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const int wpg = (nwarps - 1) / G;             // warps per pipeline (1 producer + ncw consumers)
    const int ncw = wpg - 1;
    const bool is_window_warp = warp_all == nwarps - 1;
    const int grp = is_window_warp ? 0 : warp_all / wpg, warp = is_window_warp ? 0 : warp_all % wpg;
    // ---- carve shared memory ----
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* full = bars + grp * S;
    uint64_t* empty = bars + (G + grp) * S;
    uint64_t* bfull = bars + 2 * G * S;
    uint64_t* wfree = bfull + 1;
    volatile int* pub_seq = reinterpret_cast<volatile int*>(wfree + 1);       // panels published so far
    int* tile_ctr = const_cast<int*>(pub_seq) + 1;                            // [2]
    int4* pdesc = reinterpret_cast<int4*>(const_cast<int*>(pub_seq) + 4);     // [2]
    PanelHeader* hdr0 = reinterpret_cast<PanelHeader*>(pdesc + 2);
    PanelHeader* hdr = hdr0 + grp * S;
    unsigned char* order = reinterpret_cast<unsigned char*>(hdr0 + G * S) + 32 * warp_all;
    unsigned char* win = smem + panel_ctrl_bytes(G, S);
    const int stage_bytes = stream_stage_bytes(TR, C);
    unsigned char* stage0 = win + p.win_bytes + (size_t)grp * S * stage_bytes;
    const int rp_bytes = (TR + 8) * 4, arr_bytes = (C + 8) * 4;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G * S; s++) { mbar_init(bars + s, 1); mbar_init(bars + G * S + s, ncw); }
        mbar_init(bfull, 1);
        mbar_init(wfree, G * ncw);
        *pub_seq = 0;
        tile_ctr[0] = tile_ctr[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (!is_window_warp && warp_all >= G * wpg) return;   // warps beyond G whole pipelines (none for the shipped geometries)

    if (is_window_warp) {
        // =========================== WINDOW WARP (one lane) ===========================
        if (lane != 0) return;
        const int npanels = __ldg(p.npanels);
        int pid = atomicAdd(p.panel_counter, 1);
        int4 pd = pid < npanels ? __ldg(p.panels + pid) : make_int4(-1, -1, 0, 0);
        for (int q = 0;; q++) {
            const int4 cur = pd;
            // slot q&1 was last used by panel q-2, which every producer has left (wfree of q-2 was seen below)
            pdesc[q & 1] = cur;
            tile_ctr[q & 1] = 0;
            __threadfence_block();
            *pub_seq = q + 1;
            if (cur.x < 0) return;
            // claim the panel after this one while this one is being processed
            pid = atomicAdd(p.panel_counter, 1);
            pd = pid < npanels ? __ldg(p.panels + pid) : make_int4(-1, -1, 0, 0);
            if (q > 0) mbar_wait(wfree, (uint32_t)((q - 1) & 1));     // the previous window is no longer read
            const uint32_t bytes = (uint32_t)cur.w * rowbytes;
            if (bytes > 0 && !(p.dbg & 1)) {
                // generic-proxy reads of the old window happen-before (via wfree) the async-proxy writes of the new one
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive_expect_tx(bfull, bytes);
                const unsigned char* src = reinterpret_cast<const unsigned char*>(p.Bm) + (size_t)(unsigned)cur.z * rowbytes;
                uint32_t off = 0;
                while (off < bytes) {
                    const uint32_t n = min(bytes - off, 32768u);
                    bulk_g2s(win + off, src + off, n, bfull);
                    off += n;
                }
            } else {
                mbar_arrive(bfull);
            }
        }
    }

    if (warp == 0) {
        // =========================== PRODUCER ===========================
        const int nnz_total = __ldg(p.rowptr + p.nrows);
        const int SUB = TR / 32;
        const int long_thresh = min(p.long_thresh, C - 4);
        int stage = 0;
        uint32_t ephase = 1;
        int pseq = 0, win_base = 0, win_rows = 0, chunk_base = 0;

        auto emit = [&](int rb, int re, int kb, int ke, int flags) {
            mbar_wait(empty + stage, ephase);
            unsigned char* st = stage0 + (size_t)stage * stage_bytes;
            int* rp_s = reinterpret_cast<int*>(st);
            int* col_s = reinterpret_cast<int*>(st + rp_bytes);
            float* val_s = reinterpret_cast<float*>(st + rp_bytes + arr_bytes);
            uint32_t tx = 0;
            const int rb_al = rb & ~3, kb_al = kb & ~3;
            if (re > rb) {
                {
                    const int tot_safe = (p.nrows + 1) & ~3;
                    const int want = ((re + 1 - rb_al) + 3) & ~3;
                    const int bulk = max(0, min(want, tot_safe - rb_al));
                    const int rem_lo = rb_al + bulk;
                    if (lane <= re - rem_lo && lane < 4) rp_s[bulk + lane] = __ldg(p.rowptr + rem_lo + lane);
                    if (lane == 0 && bulk > 0) bulk_g2s(rp_s, p.rowptr + rb_al, bulk * 4, full + stage);
                    tx += bulk * 4;
                }
                if (ke > kb) {
                    const int tot_safe = nnz_total & ~3;
                    const int want = ((ke - kb_al) + 3) & ~3;
                    const int bulk = max(0, min(want, tot_safe - kb_al));
                    const int rem_lo = kb_al + bulk;
                    if (lane < ke - rem_lo && lane < 4) {
                        col_s[bulk + lane] = __ldg(p.col + rem_lo + lane);
                        val_s[bulk + lane] = __ldg(p.val + rem_lo + lane);
                    }
                    if (lane == 0 && bulk > 0) {
                        bulk_g2s(col_s, p.col + kb_al, bulk * 4, full + stage);
                        bulk_g2s(val_s, p.val + kb_al, bulk * 4, full + stage);
                    }
                    tx += bulk * 8;
                }
            }
            if (lane == 0) {
                PanelHeader h;
                h.row_begin = rb; h.nrows = re - rb; h.kbase = kb_al; h.roff = rb - rb_al;
                h.flags = flags | (chunk_base << 8); h.pseq = pseq; h.win_base = win_base; h.win_rows = win_rows;
                hdr[stage] = h;
            }
            chunk_base = (chunk_base + (re - rb + 31) / 32) % ncw;
            __syncwarp();
            if (lane == 0) { if (tx) mbar_arrive_expect_tx(full + stage, tx); else mbar_arrive(full + stage); }
            if (++stage == S) { stage = 0; ephase ^= 1; }
        };

        for (;; pseq++) {
            if (lane == 0) { while (*pub_seq < pseq + 1) __nanosleep(32); }
            __syncwarp();
            __threadfence_block();
            const int4 pd = pdesc[pseq & 1];
            if (pd.x < 0) break;
            const int pb = pd.x, pe = pd.y;
            win_base = pd.z; win_rows = pd.w;
            int* tctr = tile_ctr + (pseq & 1);
            const int ntiles = (pe - pb + TR - 1) / TR;
            // piece j of tile t covers rows [a + j*SUB, a + (j+1)*SUB) clipped to the tile; the samples of the next
            // tile are loaded while this one is being streamed
            auto sample = [&](int t, int& r_lo, int& r_hi, int& s_lo, int& s_hi) {
                const int a = t < ntiles ? pb + t * TR : pe;
                const int tile_end = min(a + TR, pe);
                r_lo = min(a + lane * SUB, tile_end);
                r_hi = min(r_lo + SUB, tile_end);
                s_lo = __ldg(p.rowptr + r_lo);
                s_hi = __ldg(p.rowptr + r_hi);
            };
            int t_cur = __shfl_sync(0xffffffffu, lane == 0 ? atomicAdd(tctr, 1) : 0, 0);
            int n_rlo, n_rhi, n_slo, n_shi;
            sample(t_cur, n_rlo, n_rhi, n_slo, n_shi);
            while (t_cur < ntiles) {
                const int a = pb + t_cur * TR;
                const int tile_end = min(a + TR, pe);
                const int r_lo = n_rlo, r_hi = n_rhi, s_lo = n_slo, s_hi = n_shi;
                t_cur = __shfl_sync(0xffffffffu, lane == 0 ? atomicAdd(tctr, 1) : 0, 0);
                sample(t_cur, n_rlo, n_rhi, n_slo, n_shi);
                int piece = 0;
                while (piece < 32) {
                    const int pr_lo = __shfl_sync(0xffffffffu, r_lo, piece);
                    if (pr_lo >= tile_end) break;
                    const int kb = __shfl_sync(0xffffffffu, s_lo, piece);
                    // pieces piece..e fit one stage if their span (from the aligned-down start) does
                    const bool fits = lane >= piece && (s_hi - (kb & ~3)) <= C - 4;
                    const unsigned nofit = ~__ballot_sync(0xffffffffu, fits) & (0xffffffffu << piece);
                    const int e = nofit ? (__ffs(nofit) - 1) : 32;
                    if (e > piece) {
                        const int re = __shfl_sync(0xffffffffu, r_hi, e - 1);
                        const int ke = __shfl_sync(0xffffffffu, s_hi, e - 1);
                        emit(pr_lo, re, kb, ke, 0);
                        piece = e;
                        continue;
                    }
                    // slow path: this piece alone overflows a stage -> row by row; a row that fits no stage goes to the
                    // long-row kernel
                    const int pr_hi = __shfl_sync(0xffffffffu, r_hi, piece);
                    int run_rb = pr_lo, run_kb = kb;
                    for (int r = pr_lo; r < pr_hi; r++) {
                        const int k0 = __ldg(p.rowptr + r), k1 = __ldg(p.rowptr + r + 1);
                        const bool is_long = (k1 - k0) > long_thresh;
                        if (is_long || (k1 - (run_kb & ~3)) > C - 4) {
                            if (r > run_rb) emit(run_rb, r, run_kb, k0, 0);
                            if (is_long) {
                                if (lane == 0) p.long_rows[atomicAdd(p.long_count, 1)] = r;
                                run_rb = r + 1; run_kb = k1;
                            } else {
                                run_rb = r; run_kb = k0;
                            }
                        }
                    }
                    if (pr_hi > run_rb) emit(run_rb, pr_hi, run_kb, __ldg(p.rowptr + pr_hi), 0);
                    piece++;
                }
            }
            emit(pb, pb, 0, 0, PANEL_END);          // every pipeline ends every panel, with or without tiles of its own
        }
        mbar_wait(empty + stage, ephase);
        if (lane == 0) {
            PanelHeader h;
            h.row_begin = 0; h.nrows = -1; h.kbase = 0; h.roff = 0; h.flags = 0; h.pseq = pseq; h.win_base = 0; h.win_rows = 0;
            hdr[stage] = h;
            mbar_arrive(full + stage);
        }
        return;
    }

    // =========================== CONSUMERS ===========================
    const int cw = warp - 1;
    const int g = lane / LPR, l = lane % LPR;
    const uint32_t ws_lane = smem_u32(win) + l * 16;
    const char* bg_lane = reinterpret_cast<const char*>(p.Bm + l);
    int cur_pseq = -1;
    // One predicated LDS (column inside the window) and one predicated LDG (outside: never for a clean panel) per
    // gather, no branch.
    auto gather = [&](int c, int v, bool live, int win_base, unsigned win_rows) -> float4 {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        const unsigned w = (unsigned)(c - win_base);
        const bool in = w < win_rows;
        const uint32_t sa = ws_lane + w * rowbytes + v * (LPR * 16);
        const float4* ga = reinterpret_cast<const float4*>(bg_lane + (size_t)(unsigned)c * rowbytes) + v * LPR;
        asm volatile(
            "{\n\t.reg .pred ps, pg;\n\t"
            "setp.ne.u32 ps, %6, 0;\n\t"
            "setp.ne.u32 pg, %7, 0;\n\t"
            "@ps ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n\t"
            "@pg ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%5];\n\t}"
            : "+f"(r.x), "+f"(r.y), "+f"(r.z), "+f"(r.w)
            : "r"(sa), "l"(ga), "r"((unsigned)(live && in)), "r"((unsigned)(live && !in)));
        return r;
    };
    auto store_row = [&](int row, const float4* acc) {
        float4* orow = p.out + (size_t)row * (LPR * NV) + l;
#pragma unroll
        for (int v = 0; v < NV; v++) {
            float4 r = acc[v];
            if (p.relu) r = relu4(r);
            if (p.dbg & 4) continue;
            if (p.streaming_store) __stcs(orow + v * LPR, r); else orow[v * LPR] = r;
        }
    };

    int stage = 0;
    uint32_t fphase = 0;
    for (;;) {
        mbar_wait(full + stage, fphase);
        const int4 h0 = *reinterpret_cast<const int4*>(hdr + stage);
        if (h0.y < 0) break;
        const int4 h1 = *(reinterpret_cast<const int4*>(hdr + stage) + 1);
        if (h1.y != cur_pseq) {
            // always wait, marker stages included: a parity wait may only ever be one phase behind the barrier
            cur_pseq = h1.y;
            mbar_wait(bfull, (uint32_t)(cur_pseq & 1));
        }
        const int win_base = h1.z;
        const unsigned win_rows = (unsigned)h1.w;
        const int row_begin = h0.x, nrows = h0.y;
        const unsigned char* st = stage0 + (size_t)stage * stage_bytes;
        const int* rp_s = reinterpret_cast<const int*>(st) + h0.w;
        const int* col_k = reinterpret_cast<const int*>(st + rp_bytes) - h0.z;
        const float* val_k = reinterpret_cast<const float*>(st + rp_bytes + arr_bytes) - h0.z;
        const int nchunks = (nrows + 31) >> 5;
        const int chunk_base = h1.x >> 8;

        // chunk j of the stage belongs to warp (chunk_base + j) mod ncw
        for (int j = (cw - chunk_base + ncw) % ncw; j < nchunks; j += ncw) {
            const int r0 = j * 32;                        // first row of the chunk (stage-relative)
            const int nr = min(32, nrows - r0);
            const bool have = lane < nr;
            if (p.dbg & 8) continue;
            const int beg = have ? rp_s[r0 + lane] : 0;
            const int len = have ? rp_s[r0 + lane + 1] - beg : 0;
            // rows for the long-row kernel (as in the gather kernel) and rows the whole warp takes at once
            if (len > p.long_thresh) p.long_rows[atomicAdd(p.long_count, 1)] = row_begin + r0 + lane;
            const bool skip = len > p.hub_thresh;
            unsigned hm = __ballot_sync(0xffffffffu, skip && len <= p.long_thresh);
            // is every column of the chunk inside the window?  (always, unless the plan is stale)
            bool outside = false;
            {
                const int kb = __shfl_sync(0xffffffffu, beg, 0);
                const int ke = __shfl_sync(0xffffffffu, beg + len, nr - 1);
                for (int kk = kb + lane; kk < ke; kk += 32) outside |= (unsigned)(col_k[kk] - win_base) >= win_rows;
            }
            const bool clean = !__any_sync(0xffffffffu, outside);
            // Rows too long for one group: STEP non-zeros per pass (one per lane), every group gathers its SLOTS of
            // them, then the FMA chain runs group after group with the accumulator passed by shuffle: the CSR order of
            // additions, one gather latency per STEP non-zeros.
            while (hm) {
                const int src = __ffs(hm) - 1;
                hm &= hm - 1;
                const int hk = __shfl_sync(0xffffffffu, beg, src), hend = hk + __shfl_sync(0xffffffffu, len, src);
                constexpr int STEP = RPW * SLOTS;
                float4 hacc[NV];
#pragma unroll
                for (int v = 0; v < NV; v++) hacc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int kk = hk; kk < hend; kk += STEP) {
                    int c = 0; float a = 0.f;
                    if (lane < STEP && kk + lane < hend) { c = col_k[kk + lane]; a = val_k[kk + lane]; }
                    float4 b[SLOTS][NV];
                    float as[SLOTS];
#pragma unroll
                    for (int s = 0; s < SLOTS; s++) {
                        const int cs = __shfl_sync(0xffffffffu, c, g * SLOTS + s);
                        as[s] = __shfl_sync(0xffffffffu, a, g * SLOTS + s);
                        const bool live = kk + g * SLOTS + s < hend;
#pragma unroll
                        for (int v = 0; v < NV; v++) b[s][v] = gather(cs, v, live, win_base, win_rows);
                    }
#pragma unroll
                    for (int gg = 0; gg < RPW; gg++) {
                        // the chain so far sits in group gg-1 (for gg = 0: in the last group of the previous pass)
                        const int from = ((gg + RPW - 1) % RPW) * LPR + l;
                        if (RPW > 1) {
#pragma unroll
                            for (int v = 0; v < NV; v++) {
                                hacc[v].x = __shfl_sync(0xffffffffu, hacc[v].x, from);
                                hacc[v].y = __shfl_sync(0xffffffffu, hacc[v].y, from);
                                hacc[v].z = __shfl_sync(0xffffffffu, hacc[v].z, from);
                                hacc[v].w = __shfl_sync(0xffffffffu, hacc[v].w, from);
                            }
                        }
                        if (g == gg) {
#pragma unroll
                            for (int s = 0; s < SLOTS; s++)
                                if (kk + gg * SLOTS + s < hend) {
#pragma unroll
                                    for (int v = 0; v < NV; v++) fma4s(hacc[v], as[s], b[s][v]);
                                }
                        }
                    }
                }
                if (g == RPW - 1) store_row(row_begin + r0 + src, hacc);       // the last group holds the chain
            }
            // ---- counting sort of the chunk's rows by length (clipped): lock-step passes then run over rows of equal
            //      length ----
            constexpr int NB = 12;
            const int bk = skip ? 0 : min(len, NB - 1);
            const unsigned lt = (1u << lane) - 1u;
            int base = 0, pos = 0;
#pragma unroll
            for (int bkt = 0; bkt < NB; bkt++) {
                const unsigned m = __ballot_sync(0xffffffffu, have && bk == bkt);
                if (bk == bkt) pos = base + __popc(m & lt);
                base += __popc(m);
            }
            __syncwarp();
            if (have) order[pos] = (unsigned char)(lane | (skip ? 0x80 : 0));
            __syncwarp();
            // ---- passes: group g of pass ps takes the row at sorted position ps * RPW + g.  Two non-zeros per step,
            //      the next pair's column / value loads issued under the current pair's gathers ----
            auto run_passes = [&](auto clean_tag) {
                constexpr bool CLEAN = decltype(clean_tag)::value;
                auto fetch = [&](int c, int v, bool live) -> float4 {
                    if (CLEAN) {
                        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                        const uint32_t sa = ws_lane + (uint32_t)(c - win_base) * rowbytes + v * (LPR * 16);
                        asm volatile(
                            "{\n\t.reg .pred ps;\n\t"
                            "setp.ne.u32 ps, %5, 0;\n\t"
                            "@ps ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
                            : "+f"(r.x), "+f"(r.y), "+f"(r.z), "+f"(r.w)
                            : "r"(sa), "r"((unsigned)live));
                        return r;
                    }
                    return gather(c, v, live, win_base, win_rows);
                };
                for (int ps = 0; ps * RPW < nr; ps++) {
                    const int sp = ps * RPW + g;
                    int row = -1, k0 = 0, ln = 0;
                    if (sp < nr) {
                        const int o = order[sp];
                        if (!(o & 0x80)) { row = o; k0 = rp_s[r0 + o]; ln = rp_s[r0 + o + 1] - k0; }
                    }
                    const int maxlen = __reduce_max_sync(0xffffffffu, ln);
                    float4 acc[NV];
#pragma unroll
                    for (int v = 0; v < NV; v++) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int* cp = col_k + k0;
                    const float* vp = val_k + k0;
                    int c0 = 0, c1 = 0; float a0 = 0.f, a1 = 0.f;
                    if (0 < ln) { c0 = cp[0]; a0 = vp[0]; }
                    if (1 < ln) { c1 = cp[1]; a1 = vp[1]; }
                    for (int jj = 0; jj < maxlen; jj += 2) {
                        float4 b0[NV], b1[NV];
#pragma unroll
                        for (int v = 0; v < NV; v++) { b0[v] = fetch(c0, v, jj < ln); b1[v] = fetch(c1, v, jj + 1 < ln); }
                        const float x0 = a0, x1 = a1;
                        c0 = c1 = 0; a0 = a1 = 0.f;
                        if (jj + 2 < ln) { c0 = cp[jj + 2]; a0 = vp[jj + 2]; }
                        if (jj + 3 < ln) { c1 = cp[jj + 3]; a1 = vp[jj + 3]; }
#pragma unroll
                        for (int v = 0; v < NV; v++) fma4s(acc[v], x0, b0[v]);
#pragma unroll
                        for (int v = 0; v < NV; v++) fma4s(acc[v], x1, b1[v]);
                    }
                    if (row >= 0) store_row(row_begin + r0 + row, acc);
                }
            };
            if (p.dbg & 2) continue;
            if (clean) run_passes(TrueTag()); else run_passes(FalseTag());
        }
        __syncwarp();
        if (lane == 0) {
            if (h1.x & PANEL_END) mbar_arrive(wfree);      // this warp issues no more reads of the panel's window
            mbar_arrive(empty + stage);
        }
        if (++stage == S) { stage = 0; fphase ^= 1; }
    }
}

// ----------------------------------------------------------------------------------------------
// Panel plan: diagonal blocks of the adjacency -> panels.  Row r starts a block iff every earlier row only
// references columns < r and every row from r on only references columns >= r.
// ----------------------------------------------------------------------------------------------
namespace plan {

__global__ void row_extent_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int n,
                                  int* __restrict__ hi, int* __restrict__ lo_rev) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int b = rowptr[r], e = rowptr[r + 1];
    int mx = -1, mn = 0x7fffffff;
    for (int k = b; k < e; k++) { const int c = __ldg(col + k); mx = max(mx, c); mn = min(mn, c); }
    hi[r] = mx;
    lo_rev[n - 1 - r] = mn;
}

__global__ void block_flag_kernel(const int* __restrict__ pmax, const int* __restrict__ smin_rev, int n,
                                  unsigned char* __restrict__ flag) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    flag[r] = (r == 0) || (pmax[r - 1] < r && smin_rev[n - 1 - r] >= r);
}

// One CTA: the block starts are read in shared-memory chunks, thread 0 packs them greedily.  A block of more than
// cap_fit rows cannot have a window: it is cut into no-window panels (win_rows = 0, global gathers).
// info[0] = number of panels, info[1] = rows that sit in windowed panels.
__global__ void pack_kernel(const int* __restrict__ starts, const int* __restrict__ nstarts_p, int n, int cap_fit, int cap_pack,
                            int4* __restrict__ panels, int* __restrict__ info) {
    __shared__ int sh[1025];
    const int nb = *nstarts_p;
    int np = 0, windowed = 0, cur_b = 0, cur_e = 0;
    auto flush = [&]() {
        if (cur_e > cur_b) { panels[np++] = make_int4(cur_b, cur_e, cur_b, cur_e - cur_b); windowed += cur_e - cur_b; }
    };
    for (int base = 0; base < nb; base += 1024) {
        __syncthreads();
        for (int i = threadIdx.x; i <= 1024; i += blockDim.x) {
            const int j = base + i;
            sh[i] = j < nb ? starts[j] : n;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int m = min(1024, nb - base);
            for (int i = 0; i < m; i++) {
                const int s = sh[i], e = sh[i + 1];
                if (e - s > cap_fit) {
                    flush();
                    for (int r = s; r < e; r += cap_pack) panels[np++] = make_int4(r, min(r + cap_pack, e), 0, 0);
                    cur_b = cur_e = e;
                    continue;
                }
                if (e - cur_b > cap_pack && cur_e > cur_b) { flush(); cur_b = s; }
                cur_e = e;
            }
        }
    }
    if (threadIdx.x == 0) { flush(); info[0] = np; info[1] = windowed; }
}

}  // namespace plan
}  // namespace sgrace
