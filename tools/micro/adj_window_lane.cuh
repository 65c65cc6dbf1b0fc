// adj_window_lane.cuh -- EXPERIMENT, not part of the library (measured slower than the shipped gather kernel, see
// DESIGN.md section 3.1b and profiles/r2_adj_window.txt).  ADJ stage for block-diagonal adjacencies (batched graphs) with the gathers served from shared
// memory, float32, P_w = 16 (an XW row is 64 bytes = four 16-byte chunks):
//
//     D[r,:] = act( sum_k val[k] * XW[col[k],:] )
//
// A panel of rows of a block-diagonal adjacency only references a contiguous window of XW rows.  The window is
// bulk-copied (TMA) into shared memory once per panel, the panel's CSR arrays are streamed through a ring of stages,
// and a column outside the window (none for a clean panel) falls back to a global load, so the result never depends
// on the panel plan.  This is the PIPO hand-off of the reference (the XW tile held on chip between loop_fea and
// loop_adj, kernelMatrixmult_all.cpp:3651-3713) at the granularity of a graph, and it turns the latency-bound L2
// gathers of spmm_stream_f32_kernel (13.5 M 64-byte gathers per cora_x1024 launch) into shared-memory reads.
//
// Reference behaviour being replaced (not ported): loop_adj and dsp_kernel_wrapper_adj_* (kernelMatrixmult_all.cpp:
// 3339-3627, 1778-1957): one non-zero per cycle per hardware thread, sblocks of SPMM_BLOCK rows sharing the pipeline.
// Here adjacency rows are short (4.9 non-zeros on Cora), so one LANE owns one row: every warp instruction covers 32
// non-zeros, a lane takes its row of the next chunk the moment it finishes the current one (the sblock idea taken to
// its limit: no lane waits for the longest row of a pass), and a stage goes back to the producer when every lane of
// the warp has left it.  Rows longer than `long_thresh` (hubs) are listed for the long-row kernel.  Accumulation is
// float FMA in CSR order within a row: deterministic, bit-equal to spmm_stream_f32_kernel / spmm_csr_f32_kernel.
//
// Structure: one persistent CTA per SM; warp 0 is the producer (claims panels from a global counter, loads the window,
// streams chunks of 32 x (consumer warps) rows -- rowPtr / columnIndex / values slices -- with cp.async.bulk completing
// on mbarriers, ends every panel with a marker chunk), the other warps consume.
#pragma once
#include "../../sgracex1_b200/csrc/sgrace_spmm_stream.cuh"

#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

namespace sgrace {

// L2 prefetch of a contiguous global range (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

enum { LANE_END = 1 };

struct LaneParams {
    const int* rowptr;
    const int* col;
    const float* val;
    const float4* Bm;          // XW row-major (global)
    float4* out;
    int nrows, relu, streaming_store;
    int long_thresh;           // rows with more non-zeros are deferred to the long-row kernel
    int stage_nnz;             // C: non-zeros a stage can hold (multiple of 4)
    int stages;                // S
    int win_bytes;             // capacity of the window, bytes (multiple of 64)
    const int4* panels;        // {row_begin, row_end, win_base, win_rows}
    const int* npanels;        // number of panels (device memory, written by the planner)
    int* long_rows;
    int* long_count;
    int* span_counter;
    int dbg;                   // measurement only: 1 = skip the arithmetic, 2 = skip the window copies (results are then wrong)
};

struct LaneHeader {            // 32 bytes
    int row_begin;             // first row of the chunk
    int nrows;                 // rows in the chunk; < 0: no more work
    int kbase;                 // global index of the non-zero stored at col_s[0] / val_s[0] (multiple of 4)
    int roff;                  // rp_s[roff + i] is rowptr[row_begin + i]
    int flags;                 // LANE_END: end-of-panel marker
    int pseq;                  // panel sequence number within this CTA
    int win_base;              // first XW row held in the window
    int win_rows;
};

// non-blocking test (try_wait may suspend the warp for a system-dependent time; the callers have other work to do)
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ int4 lds_i4(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

__host__ __device__ inline int lane_stage_bytes(int chunk_rows, int stage_nnz) {
    return (chunk_rows + 8) * 4 + 2 * (stage_nnz + 8) * 4;
}
// total dynamic shared memory of a launch; mirrors the carve-up at the top of the kernel
inline size_t lane_smem_bytes(int stages, int chunk_rows, int stage_nnz, int win_bytes) {
    size_t off = (8 * (2 * (size_t)stages + 2) + 31) & ~(size_t)31;   // mbarriers
    off += 32 * (size_t)stages;                                       // headers
    off = (off + 127) & ~(size_t)127;
    off += ((size_t)win_bytes + 127) & ~(size_t)127;                  // XW window
    off += (size_t)stages * lane_stage_bytes(chunk_rows, stage_nnz);
    return off + 128;
}

__global__ void __launch_bounds__(544, 1)
adj_window_f32_kernel(const LaneParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int S = p.stages, C = p.stage_nnz;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (blockDim.x >> 5) - 1;       // consumer warps
    const int CH = ncw * 32;                     // rows per chunk: one per consumer lane
    // ---- carve shared memory ----
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [S]
    uint64_t* empty = full + S;                                   // [S]
    uint64_t* wfull = full + 2 * S;                               // window of the current panel landed
    uint64_t* wempty = wfull + 1;                                 // every consumer warp has finished the panel
    LaneHeader* hdr = reinterpret_cast<LaneHeader*>(smem + ((8 * (2 * S + 2) + 31) & ~31));
    unsigned char* cur = reinterpret_cast<unsigned char*>(hdr + S);
    cur = smem + (((cur - smem) + 127) & ~127);
    unsigned char* Ws = cur;
    cur += (p.win_bytes + 127) & ~127;
    unsigned char* stage0 = cur;
    const int stage_bytes = lane_stage_bytes(CH, C);
    const int rp_bytes = (CH + 8) * 4, arr_bytes = (C + 8) * 4;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(full + s, 1); mbar_init(empty + s, ncw); }
        mbar_init(wfull, 1);
        mbar_init(wempty, ncw);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // =========================== PRODUCER ===========================
        const int nnz_total = __ldg(p.rowptr + p.nrows);
        const int nspans = __ldg(p.npanels);
        const int long_thresh = min(p.long_thresh, C - 4);
        int stage = 0;
        uint32_t ephase = 1;                      // waiting on parity 1 of a fresh barrier passes at once
        uint32_t wephase = 1;
        int pseq = 0;
        bool win_issued = true;
        int win_base = 0, win_rows = 0;

        auto issue_window = [&]() {
            if (lane == 0) {
                const uint32_t bytes = (p.dbg & 2) ? 0u : (uint32_t)win_rows * 64u;
                if (bytes) mbar_arrive_expect_tx(wfull, bytes); else mbar_arrive(wfull);
                const unsigned char* src = reinterpret_cast<const unsigned char*>(p.Bm) + (size_t)win_base * 64;
                for (uint32_t off = 0; off < bytes; off += 32768)
                    bulk_g2s(Ws + off, src + off, min(bytes - off, 32768u), wfull);
            }
            win_issued = true;
            wephase ^= 1;
        };
        // wait for a free stage; while waiting, load the window as soon as the previous panel has been drained
        auto acquire_stage = [&]() {
            if (!win_issued) {
                for (;;) {
                    if (mbar_try(wempty, wephase)) { issue_window(); break; }
                    if (mbar_try(empty + stage, ephase)) break;
                }
            }
            mbar_wait(empty + stage, ephase);
        };
        auto publish = [&](const LaneHeader& h, uint32_t tx) {
            if (lane == 0) hdr[stage] = h;
            __syncwarp();
            if (lane == 0) { if (tx) mbar_arrive_expect_tx(full + stage, tx); else mbar_arrive(full + stage); }
            if (++stage == S) { stage = 0; ephase ^= 1; }
        };
        // stream rows [rb, re) whose non-zeros [kb, ke) fit a stage
        auto emit = [&](int rb, int re, int kb, int ke) {
            acquire_stage();
            unsigned char* st = stage0 + (size_t)stage * stage_bytes;
            int* rp_s = reinterpret_cast<int*>(st);
            int* col_s = reinterpret_cast<int*>(st + rp_bytes);
            float* val_s = reinterpret_cast<float*>(st + rp_bytes + arr_bytes);
            uint32_t tx = 0;
            const int rb_al = rb & ~3;
            {   // rowPtr slice: elements [rb_al, re]; the trailing partial 16-byte group of the array by hand
                const int tot_safe = (p.nrows + 1) & ~3;
                const int want = ((re + 1 - rb_al) + 3) & ~3;
                const int bulk = max(0, min(want, tot_safe - rb_al));
                const int rem_lo = rb_al + bulk;
                if (lane <= re - rem_lo && lane < 4) rp_s[bulk + lane] = __ldg(p.rowptr + rem_lo + lane);
                if (lane == 0 && bulk > 0) bulk_g2s(rp_s, p.rowptr + rb_al, bulk * 4, full + stage);
                tx += bulk * 4;
            }
            const int kb_al = kb & ~3;
            {   // columnIndex / values slices: elements [kb_al, ke)
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
            LaneHeader h;
            h.row_begin = rb; h.nrows = re - rb; h.kbase = kb_al; h.roff = rb - rb_al;
            h.flags = 0; h.pseq = pseq; h.win_base = win_base; h.win_rows = win_rows;
            publish(h, tx);
        };

        // Everything the NEXT panel of this CTA will read (window rows, CSR slices) is requested into L2 while the
        // current panel is processed: the shared-memory ring is too small to cover HBM latency by itself (the window
        // takes 174 of the 227 KB), so HBM requests are kept in flight without a shared-memory destination and the
        // ring / window copies that follow are L2 hits.
        auto prefetch_panel = [&](int s) {
            if (p.dbg & 4) return;
            const int4 pn = __ldg(p.panels + s);
            const int k0 = __ldg(p.rowptr + pn.x), k1 = __ldg(p.rowptr + pn.y);
            constexpr uint32_t PIECE = 8192;
            const uint32_t wbytes = (uint32_t)max(0, min(pn.w, p.win_bytes / 64)) * 64u;
            const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.Bm) + (size_t)pn.z * 64;
            for (uint32_t off = lane * PIECE; off < wbytes; off += 32 * PIECE) bulk_prefetch_l2(wsrc + off, min(PIECE, wbytes - off));
            const int ka = k0 & ~3, ke = min((k1 + 3) & ~3, nnz_total & ~3);
            if (ke > ka) {
                const uint32_t kbytes = (uint32_t)(ke - ka) * 4u;
                const unsigned char* c0 = reinterpret_cast<const unsigned char*>(p.col + ka);
                const unsigned char* v0 = reinterpret_cast<const unsigned char*>(p.val + ka);
                for (uint32_t off = lane * PIECE; off < kbytes; off += 32 * PIECE) {
                    bulk_prefetch_l2(c0 + off, min(PIECE, kbytes - off));
                    bulk_prefetch_l2(v0 + off, min(PIECE, kbytes - off));
                }
            }
            const int ra = pn.x & ~3, re = min((pn.y + 4) & ~3, (p.nrows + 1) & ~3);
            if (re > ra) {
                const uint32_t rbytes = (uint32_t)(re - ra) * 4u;
                const unsigned char* r0 = reinterpret_cast<const unsigned char*>(p.rowptr + ra);
                for (uint32_t off = lane * PIECE; off < rbytes; off += 32 * PIECE) bulk_prefetch_l2(r0 + off, min(PIECE, rbytes - off));
            }
        };
        auto claim = [&]() -> int { return lane == 0 ? atomicAdd(p.span_counter, 1) : 0; };
        int s_cur = __shfl_sync(0xffffffffu, claim(), 0);
        int s_next_raw = claim();
        while (s_cur < nspans) {
            {
                const int s_nx = __shfl_sync(0xffffffffu, s_next_raw, 0);
                if (s_nx < nspans) prefetch_panel(s_nx);
            }
            const int4 pn = __ldg(p.panels + s_cur);
            const int a = pn.x, b = pn.y;
            win_base = pn.z;
            win_rows = max(0, min(pn.w, p.win_bytes / 64));
            win_issued = false;
            const int nch = (b - a + CH - 1) / CH;
            for (int cb = 0; cb < nch; cb += 32) {
                // chunk boundaries of the next 32 chunks in one round trip
                const int r_lo = min(a + (cb + lane) * CH, b), r_hi = min(r_lo + CH, b);
                const int s_lo = __ldg(p.rowptr + r_lo), s_hi = __ldg(p.rowptr + r_hi);
                const int nb = min(32, nch - cb);
                // Copy descriptors of the 32 chunks are worked out lane-parallel; a chunk whose slices are whole 16-byte
                // groups inside the arrays (all but the last of the matrix) and fit a stage is then issued by its own
                // lane: wait for the stage, three bulk copies, header, arrive -- no shuffles, no serial integer chain.
                const int rb_al = r_lo & ~3, kb_al = s_lo & ~3;
                const int rp_want = ((r_hi + 1 - rb_al) + 3) & ~3, k_want = ((s_hi - kb_al) + 3) & ~3;
                const bool fast = lane < nb && (s_hi - kb_al) <= C && rb_al + rp_want <= ((p.nrows + 1) & ~3) &&
                                  kb_al + k_want <= (nnz_total & ~3);
                const unsigned fastmask = __ballot_sync(0xffffffffu, fast);
                for (int j = 0; j < nb; j++) {
                    if ((fastmask >> j) & 1u) {
                        if (!win_issued) {
                            for (;;) {
                                if (mbar_try(wempty, wephase)) { issue_window(); break; }
                                if (mbar_try(empty + stage, ephase)) break;
                            }
                        }
                        if (lane == j) {
                            mbar_wait(empty + stage, ephase);
                            unsigned char* st = stage0 + (size_t)stage * stage_bytes;
                            bulk_g2s(st, p.rowptr + rb_al, (uint32_t)rp_want * 4u, full + stage);
                            if (k_want > 0) {
                                bulk_g2s(st + rp_bytes, p.col + kb_al, (uint32_t)k_want * 4u, full + stage);
                                bulk_g2s(st + rp_bytes + arr_bytes, p.val + kb_al, (uint32_t)k_want * 4u, full + stage);
                            }
                            LaneHeader h;
                            h.row_begin = r_lo; h.nrows = r_hi - r_lo; h.kbase = kb_al; h.roff = r_lo - rb_al;
                            h.flags = 0; h.pseq = pseq; h.win_base = win_base; h.win_rows = win_rows;
                            hdr[stage] = h;
                            mbar_arrive_expect_tx(full + stage, (uint32_t)rp_want * 4u + (uint32_t)k_want * 8u);
                        }
                        __syncwarp();
                        if (++stage == S) { stage = 0; ephase ^= 1; }
                        continue;
                    }
                    const int rlo = __shfl_sync(0xffffffffu, r_lo, j), rhi = __shfl_sync(0xffffffffu, r_hi, j);
                    const int kb = __shfl_sync(0xffffffffu, s_lo, j), ke = __shfl_sync(0xffffffffu, s_hi, j);
                    if (ke - (kb & ~3) <= C) { emit(rlo, rhi, kb, ke); continue; }
                    // slow path: the chunk overflows a stage -> row by row, over-long rows go to the list
                    int run_rb = rlo, run_kb = kb;
                    for (int r = rlo; r < rhi; r++) {
                        const int k0 = __ldg(p.rowptr + r), k1 = __ldg(p.rowptr + r + 1);
                        const bool is_long = (k1 - k0) > long_thresh;
                        if (is_long || (k1 - (run_kb & ~3)) > C) {
                            if (r > run_rb) emit(run_rb, r, run_kb, k0);
                            if (is_long) {
                                if (lane == 0) p.long_rows[atomicAdd(p.long_count, 1)] = r;
                                run_rb = r + 1; run_kb = k1;
                            } else {
                                run_rb = r; run_kb = k0;
                            }
                        }
                    }
                    if (rhi > run_rb) emit(run_rb, rhi, run_kb, __ldg(p.rowptr + rhi));
                }
            }
            // end-of-panel marker: the consumers drain, release the window and move to the next panel
            acquire_stage();
            if (!win_issued) { mbar_wait(wempty, wephase); issue_window(); }
            {
                LaneHeader h;
                h.row_begin = b; h.nrows = 0; h.kbase = 0; h.roff = 0;
                h.flags = LANE_END; h.pseq = pseq; h.win_base = win_base; h.win_rows = win_rows;
                publish(h, 0);
            }
            pseq++;
            s_cur = __shfl_sync(0xffffffffu, s_next_raw, 0);
            s_next_raw = claim();
        }
        // sentinel: no more work
        mbar_wait(empty + stage, ephase);
        if (lane == 0) {
            LaneHeader h;
            h.row_begin = 0; h.nrows = -1; h.kbase = 0; h.roff = 0; h.flags = 0; h.pseq = pseq; h.win_base = 0; h.win_rows = 0;
            hdr[stage] = h;
            mbar_arrive(full + stage);
        }
        return;
    }

    // =========================== CONSUMERS ===========================
    // Chunk-synchronous: a warp takes the 32 rows of the chunk that belong to it (one per lane), walks them in
    // lock-step for as many steps as its longest row has non-zeros, stores, and gives the stage back.  The per-chunk
    // control is ~100 instructions; the step loop is 4 shared-memory gathers + 16 FMAs per lane.
    const int cw = warp - 1;
    const int li = cw * 32 + lane;                    // this lane's row within every chunk
    // pass pp touches 16-byte chunk (lane + pp) & 3 of the row: the eight lanes of a quarter-warp spread over the
    // four chunk positions (two rows of equal parity still share a bank group)
    int chunk[4];
    uint32_t woff[4];
    const uint32_t ws_u32 = smem_u32(Ws);
#pragma unroll
    for (int pp = 0; pp < 4; pp++) { chunk[pp] = (lane + pp) & 3; woff[pp] = ws_u32 + (uint32_t)chunk[pp] * 16u; }
    const uint32_t hdr_u32 = smem_u32(hdr), stage0_u32 = smem_u32(stage0);
    int slot = 0, cur_pseq = -1;
    uint32_t fphase = 0;
    for (;;) {
        mbar_wait(full + slot, fphase);
        const int4 h0 = lds_i4(hdr_u32 + (uint32_t)slot * 32u);        // row_begin, nrows, kbase, roff
        const int4 h1 = lds_i4(hdr_u32 + (uint32_t)slot * 32u + 16u);  // flags, pseq, win_base, win_rows
        if (h0.y < 0) break;
        if (h1.x & LANE_END) {
            // this warp has finished every row of the panel: the marker's stage and the window go back
            if (lane == 0) { mbar_arrive(empty + slot); mbar_arrive(wempty); }
            if (++slot == S) { slot = 0; fphase ^= 1; }
            continue;
        }
        if (h1.y != cur_pseq) { mbar_wait(wfull, (uint32_t)h1.y & 1u); cur_pseq = h1.y; }
        const uint32_t st = stage0_u32 + (uint32_t)slot * (uint32_t)stage_bytes;
        int k = h0.z, kend = h0.z;                    // lanes without a row point at the first staged non-zero
        bool mine = li < h0.y;
        if (mine) {
            const uint32_t rpa = st + (uint32_t)(h0.w + li) * 4u;
            k = lds_s32(rpa); kend = lds_s32(rpa + 4u);
            if (kend - k > p.long_thresh) {
                p.long_rows[atomicAdd(p.long_count, 1)] = h0.x + li;
                mine = false; k = kend = h0.z;
            }
        }
        const int steps = (p.dbg & 1) ? 0 : __reduce_max_sync(0xffffffffu, kend - k);
        const uint32_t colp = st + (uint32_t)rp_bytes - (uint32_t)h0.z * 4u;
        const uint32_t valp = st + (uint32_t)(rp_bytes + arr_bytes) - (uint32_t)h0.z * 4u;
        const int win_base = h1.z;
        const uint32_t win_rows = (uint32_t)h1.w;
        float4 acc[4];
#pragma unroll
        for (int pp = 0; pp < 4; pp++) acc[pp] = make_float4(0.f, 0.f, 0.f, 0.f);
        // (cn, an) = the non-zero of the coming step; reads one element past a finished row stay inside the stage
        int cn = lds_s32(colp + (uint32_t)k * 4u);
        float an = lds_f32(valp + (uint32_t)k * 4u);
#pragma unroll 4
        for (int t = 0; t < steps; t++) {
            const bool on = k < kend;
            const float a = an;
            const uint32_t rel = (uint32_t)(cn - win_base);
            const bool out = on && rel >= win_rows;
            float4 b[4];
            if (__any_sync(0xffffffffu, out)) {
                // a column outside the window (never on a clean panel): that lane gathers from global memory
                if (out) {
                    const float4* src = p.Bm + (size_t)(unsigned)cn * 4;
#pragma unroll
                    for (int pp = 0; pp < 4; pp++) b[pp] = __ldg(src + chunk[pp]);
                } else {
                    const uint32_t ra = (on ? rel : 0u) * 64u;
#pragma unroll
                    for (int pp = 0; pp < 4; pp++) b[pp] = lds_v4(woff[pp] + ra);
                }
            } else {
                const uint32_t ra = (on ? rel : 0u) * 64u;
#pragma unroll
                for (int pp = 0; pp < 4; pp++) b[pp] = lds_v4(woff[pp] + ra);
            }
            k += on ? 1 : 0;
            cn = lds_s32(colp + (uint32_t)k * 4u);
            an = lds_f32(valp + (uint32_t)k * 4u);
            if (on) {
#pragma unroll
                for (int pp = 0; pp < 4; pp++) fma4s(acc[pp], a, b[pp]);
            }
        }
        if (mine && !(p.dbg & 8)) {
            float4* orow_p = p.out + (size_t)(h0.x + li) * 4;
#pragma unroll
            for (int pp = 0; pp < 4; pp++) {
                float4 r = acc[pp];
                if (p.relu) r = relu4(r);              // val = (acc > 0 || relu == 0) ? acc : 0   (K:2586-2590)
                if (p.streaming_store) __stcs(orow_p + chunk[pp], r); else orow_p[chunk[pp]] = r;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);
        if (++slot == S) { slot = 0; fphase ^= 1; }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Panel planner for LANE_WINDOW: finds the diagonal blocks of the adjacency (row r starts a block iff every earlier
// row only references columns < r) and groups them into panels whose rows -- and therefore, for a block-diagonal
// matrix, whose referenced XW rows -- fit the shared-memory window.  A wrong or coarse plan costs speed, never
// correctness (columns outside the window are gathered from global memory).
// ------------------------------------------------------------------------------------------------------------------

// rowmax[r] = largest column index in row r (-1 for an empty row); one thread per row
__global__ void lane_plan_rowmax_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int* __restrict__ rowmax, int nrows) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int b = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
    int m = -1;
    for (int k = b; k < e; k++) m = max(m, __ldg(col + k));
    rowmax[r] = m;
}

// bpos[r] = r if a block starts at row r, else -1.  pmax = inclusive prefix maximum of rowmax.
__global__ void lane_plan_bounds_kernel(const int* __restrict__ pmax, int* __restrict__ bpos, int nrows) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    bpos[r] = (r == 0 || __ldg(pmax + r - 1) < r) ? r : -1;
}

// cut[r] = 1 iff row r is the first block start at or after a multiple of `quantum` (prevb = exclusive prefix maximum
// of bpos = the last block start before r).  Consecutive cuts are then less than quantum + (largest block) apart.
__global__ void lane_plan_cuts_kernel(const int* __restrict__ bpos, const int* __restrict__ prevb, unsigned char* __restrict__ cut,
                                      int nrows, int quantum) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    cut[r] = (bpos[r] >= 0 && prevb[r] < (r / quantum) * quantum) ? 1 : 0;
}

// panels[i] = {starts[i], starts[i+1] | nrows, starts[i], min(rows, cap_rows)}; stats[0] += rows covered by a window
__global__ void lane_plan_panels_kernel(const int* __restrict__ starts, const int* __restrict__ nstarts, int4* __restrict__ panels,
                                        int nrows, int cap_rows, int* __restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *nstarts;
    if (i >= n) return;
    const int a = starts[i], b = (i + 1 < n) ? starts[i + 1] : nrows;
    panels[i] = make_int4(a, b, a, min(b - a, cap_rows));
    atomicAdd(stats, min(b - a, cap_rows));
    if (i == 0) stats[1] = n;
}

struct LaneMaxOp {
    __host__ __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

// device pointers of a plan, carved out of one scratch allocation by lane_plan_build
struct LanePlan {
    int4* panels = nullptr;    // [max_panels]
    int* npanels = nullptr;    // number of panels
    int* stats = nullptr;      // [0] rows covered by their panel's window, [1] number of panels
    int max_panels = 0;
};

inline size_t lane_plan_align(size_t x) { return (x + 255) & ~(size_t)255; }

// bytes of scratch lane_plan_build needs for a matrix of nrows rows
inline size_t lane_plan_scratch_bytes(int nrows, int quantum) {
    size_t scan_tmp = 0, sel_tmp = 0;
    cub::DeviceScan::InclusiveScan(nullptr, scan_tmp, (const int*)nullptr, (int*)nullptr, LaneMaxOp(), nrows);
    size_t scan2 = 0;
    cub::DeviceScan::ExclusiveScan(nullptr, scan2, (const int*)nullptr, (int*)nullptr, LaneMaxOp(), -1, nrows);
    cub::DeviceSelect::Flagged(nullptr, sel_tmp, cub::CountingInputIterator<int>(0), (const unsigned char*)nullptr, (int*)nullptr,
                               (int*)nullptr, nrows);
    size_t tmp = scan_tmp > scan2 ? scan_tmp : scan2;
    if (sel_tmp > tmp) tmp = sel_tmp;
    const size_t max_panels = (size_t)nrows / (size_t)quantum + 2;
    return lane_plan_align(tmp) + 2 * lane_plan_align(sizeof(int) * (size_t)nrows) + lane_plan_align((size_t)nrows) +
           lane_plan_align(sizeof(int) * max_panels) + lane_plan_align(sizeof(int4) * max_panels) + lane_plan_align(64);
}

// Enqueues the planner on `stream` (no host synchronisation): panels of whole diagonal blocks, consecutive panel starts
// less than quantum + (largest block) rows apart, windows clipped to cap_rows rows.
inline cudaError_t lane_plan_build(void* scratch, const int* rowptr, const int* col, int nrows, int quantum, int cap_rows,
                                   cudaStream_t stream, LanePlan* plan) {
    size_t scan_tmp = 0, sel_tmp = 0, scan2 = 0;
    cub::DeviceScan::InclusiveScan(nullptr, scan_tmp, (const int*)nullptr, (int*)nullptr, LaneMaxOp(), nrows);
    cub::DeviceScan::ExclusiveScan(nullptr, scan2, (const int*)nullptr, (int*)nullptr, LaneMaxOp(), -1, nrows);
    cub::DeviceSelect::Flagged(nullptr, sel_tmp, cub::CountingInputIterator<int>(0), (const unsigned char*)nullptr, (int*)nullptr,
                               (int*)nullptr, nrows);
    size_t tmp = scan_tmp > scan2 ? scan_tmp : scan2;
    if (sel_tmp > tmp) tmp = sel_tmp;
    const size_t max_panels = (size_t)nrows / (size_t)quantum + 2;
    unsigned char* c = (unsigned char*)scratch;
    void* cub_tmp = c; c += lane_plan_align(tmp);
    int* a0 = (int*)c; c += lane_plan_align(sizeof(int) * (size_t)nrows);      // rowmax, then bpos
    int* a1 = (int*)c; c += lane_plan_align(sizeof(int) * (size_t)nrows);      // pmax, then prevb
    unsigned char* cut = c; c += lane_plan_align((size_t)nrows);
    int* starts = (int*)c; c += lane_plan_align(sizeof(int) * max_panels);
    plan->panels = (int4*)c; c += lane_plan_align(sizeof(int4) * max_panels);
    plan->stats = (int*)c;
    plan->npanels = plan->stats + 2;
    plan->max_panels = (int)max_panels;
    const int grid = (nrows + 255) / 256;
    cudaError_t e = cudaMemsetAsync(plan->stats, 0, 16, stream);
    if (e != cudaSuccess) return e;
    lane_plan_rowmax_kernel<<<grid, 256, 0, stream>>>(rowptr, col, a0, nrows);
    size_t t = tmp;
    e = cub::DeviceScan::InclusiveScan(cub_tmp, t, (const int*)a0, a1, LaneMaxOp(), nrows, stream);
    if (e != cudaSuccess) return e;
    lane_plan_bounds_kernel<<<grid, 256, 0, stream>>>(a1, a0, nrows);
    t = tmp;
    e = cub::DeviceScan::ExclusiveScan(cub_tmp, t, (const int*)a0, a1, LaneMaxOp(), -1, nrows, stream);
    if (e != cudaSuccess) return e;
    lane_plan_cuts_kernel<<<grid, 256, 0, stream>>>(a0, a1, cut, nrows, quantum);
    t = tmp;
    e = cub::DeviceSelect::Flagged(cub_tmp, t, cub::CountingInputIterator<int>(0), (const unsigned char*)cut, starts, plan->npanels,
                                   nrows, stream);
    if (e != cudaSuccess) return e;
    lane_plan_panels_kernel<<<(int)((max_panels + 255) / 256), 256, 0, stream>>>(starts, plan->npanels, plan->panels, nrows, cap_rows,
                                                                                 plan->stats);
    return cudaGetLastError();
}

}  // namespace sgrace
