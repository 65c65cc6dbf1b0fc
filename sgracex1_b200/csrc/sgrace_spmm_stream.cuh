// sgrace_spmm_stream.cuh -- the float32 CSR x row-major SpMM of both stages as a persistent,
// warp-specialised streaming kernel for sm_100a:
//
//     out[r,:] = act( sum_k val[k] * Bm[col[k],:] )        FEA: X_csr . W      ADJ: A_csr . XW
//
// Reference behaviour being replaced (not ported): the HLS dataflow stages loop_fea / loop_adj
// (gnn-rfsoc-mt-all-2022/src/kernelMatrixmult_all.cpp:2932-3336, 3339-3627): readptr/readval FIFOs
// feeding dsp_kernel_wrapper_* one non-zero per cycle, sblocks of SPMM_BLOCK rows sharing a pass.
//
// Structure (one CTA per SM, resident for the whole launch):
//   * warp 0 is the PRODUCER.  It claims row tiles from a global counter, samples 33 row
//     pointers per tile in one round trip, packs the tile's row pieces into sub-tiles whose
//     non-zeros fit one shared-memory stage, and streams the sub-tile's three CSR slices
//     (rowPtr, columnIndex, values) into the stage with bulk TMA copies (cp.async.bulk ->
//     UBLKCP) that complete on the stage's "full" mbarrier.  HBM is therefore read in long
//     sequential bursts that run several stages ahead of the arithmetic.
//   * warps 1.. are CONSUMERS.  A row group of LPR lanes owns one CSR row; the 32/LPR groups
//     of a warp take consecutive rows (an sblock: short rows share every pass and their output
//     is one contiguous store).  A group walks its row in 16-byte aligned blocks of four
//     non-zeros: one 128-bit shared load brings four column indices, one brings four values,
//     four 128-bit gathers of Bm rows are issued back to back, then the FMAs.  Lane l holds
//     float4 chunk (v*LPR + l) of the output row, so a gathered Bm row is one coalesced
//     16*LPR-byte request.  When the stage has been read the warp arrives on its "empty"
//     mbarrier and the producer refills it.
//   * Bm is either gathered from global memory through L1 (ADJ: Bm = XW; FEA with a weight
//     matrix too large for shared memory) or staged once per CTA in shared memory (FEA: the
//     Cora-shape W is 1433 x 16 floats = 92 KB).  With 64-byte rows two row groups of a quarter-warp
//     share one shared-memory wavefront when their rows have opposite parity and take two otherwise
//     (a duplicated image that removes the conflict leaves too little room for stages: measured
//     0.32 ms against 0.25 ms).
//   * rows longer than `long_thresh` (or than a stage) are appended to a list and handled by
//     spmm_long_rows_f32_kernel afterwards (row-bucket scheduling for power-law graphs).
// Accumulation is float FMA in CSR order within a row: deterministic, no atomics on values.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgrace {

enum { BSRC_GLOBAL = 0, BSRC_SMEM = 1 };
constexpr int MAX_PEERS = 8;

// quantised full-design ADJ (demo/sgrace_lib/sgrace.py:626-667): adjacency codes formed on the fly, zero codes are
// pruned edges, float multiply THEN add in CSR order (the emulation's arithmetic), ReLU, dequantise
struct StreamQ {
    float inv_as;            // 1 / a_s
    int a_z, qbits;
    float den;               // 2^(q-1) (2 for q = 1): a power of two, so the division below is an exact scaling
    float deq_o;
    int quant;               // 0: full-design arithmetic without quantisation
};
enum { QM_NONE = 0, QM_GCN = 1 };

struct StreamParams {
    const int* rowptr;
    const int* col;
    const float* val;
    const float4* Bm;        // row-major, row stride P4 float4 (BSRC_GLOBAL) or the pre-laid-out image (SMEM*)
    float4* out;
    int nrows, P4, relu;
    int long_thresh;         // rows with more non-zeros are deferred
    int tile_rows;           // TR: rows per claimed tile, multiple of 32, <= 1024
    int stage_nnz;           // C: non-zeros a stage can hold (multiple of 4)
    int stages;              // S: stages per group
    int groups;              // G: independent producer/consumer pipelines per CTA
    int b_bytes;             // bytes of the Bm image staged in shared memory (SMEM*), multiple of 16
    int streaming_store;     // 1: out is not re-read soon (D) -> st.global.cs
    int accumulate;          // 1: out = act(out + A.Bm)  (second pass over a split adjacency)
    int* long_rows;
    int* long_count;
    int* tile_counter;
    // PEER gathers: Bm is row-partitioned over `peer_count` GPUs, rank r holds rows [r*peer_block, (r+1)*peer_block)
    // at peer_base[r] (its own buffer or a peer-mapped one): the gather goes straight over NVLink
    const char* peer_base[MAX_PEERS];
    int peer_block;
    int peer_count;
    StreamQ q;               // QM_GCN kernels only
};

// ----------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (TMA, 1-D); dst/src 16-byte aligned, bytes a positive multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct StageHeader {
    int row_begin;   // first row of the sub-tile
    int nrows;       // rows in the sub-tile; < 0: no more work
    int kbase;       // global index of the non-zero stored at col_s[0] / val_s[0] (multiple of 4)
    int roff;        // rp_s[roff + i] is rowptr[row_begin + i]
};

__device__ __forceinline__ void fma4s(float4& a, float s, const float4& b) {
    a.x = fmaf(s, b.x, a.x); a.y = fmaf(s, b.y, a.y);
    a.z = fmaf(s, b.z, a.z); a.w = fmaf(s, b.w, a.w);
}

__device__ __forceinline__ void mul_add4s(float4& a, float s, const float4& b) {     // rounded product, rounded sum
    a.x = __fadd_rn(a.x, __fmul_rn(s, b.x)); a.y = __fadd_rn(a.y, __fmul_rn(s, b.y));
    a.z = __fadd_rn(a.z, __fmul_rn(s, b.z)); a.w = __fadd_rn(a.w, __fmul_rn(s, b.w));
}
// quantization_ufbits (sgrace.py:253-265) / 2^(q-1): torch.round = round-half-even
__device__ __forceinline__ float stream_adj_code(float x, const StreamQ& q) {
    float r = rintf(__fadd_rn(__fmul_rn(q.inv_as, x), (float)q.a_z));
    const float hi = (float)((1 << q.qbits) - 1);
    r = r < 0.f ? 0.f : r;
    r = r > hi ? hi : r;
    return __fmul_rn(r, __frcp_rn(q.den));       // == r / den bit for bit (den is a power of two)
}

__device__ __forceinline__ float4 relu4(float4 r) {
    r.x = r.x > 0.f ? r.x : 0.f; r.y = r.y > 0.f ? r.y : 0.f;
    r.z = r.z > 0.f ? r.z : 0.f; r.w = r.w > 0.f ? r.w : 0.f;
    return r;
}

// shared-memory footprint of one stage, bytes (each array padded so that aligned-down starts and
// the 16-byte rounding of the ends stay inside)
__host__ __device__ inline int stream_stage_bytes(int tile_rows, int stage_nnz) {
    return (tile_rows + 8) * 4 + 2 * (stage_nnz + 8) * 4;
}
// total dynamic shared memory of a launch; mirrors the carve-up at the top of the kernel
inline size_t stream_smem_bytes(int groups, int stages, int tile_rows, int stage_nnz, int b_bytes) {
    const int gs = groups * stages;
    size_t off = (8 * (2 * gs + 1) + 15) & ~15;           // mbarriers
    off += 16 * (size_t)gs;                                // stage headers
    off = (off + 127) & ~(size_t)127;
    off += ((size_t)b_bytes + 127) & ~(size_t)127;         // Bm image
    off += (size_t)gs * stream_stage_bytes(tile_rows, stage_nnz);
    return off + 128;                                      // slack for the 128-byte alignment of the base
}

// EXACT: P4 == LPR*NV, so every lane owns live columns and the row stride of Bm is a constant.
template <int LPR, int NV, int BSRC, int MAXT, int MINB, bool EXACT, bool PEER = false, int QM = QM_NONE>
__global__ void __launch_bounds__(MAXT, MINB)
spmm_stream_f32_kernel(const StreamParams p) {
    constexpr int RPW = 32 / LPR;                 // row groups (= rows in flight) per warp
    extern __shared__ __align__(128) unsigned char smem[];
    const int S = p.stages, TR = p.tile_rows, C = p.stage_nnz, G = p.groups;
    // A CTA is G independent pipelines ("groups"): each has its own producer warp, consumer warps,
    // mbarriers and stage ring; all share the Bm image.  One producer warp issues ~250 dependent
    // instructions per stage (~0.9 us), so a single one cannot feed an SM (measured: 16 KB/us).
    const int warp_all = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpg = (blockDim.x >> 5) / G;        // warps per group
    const int grp = warp_all / wpg, warp = warp_all % wpg;
    const int ncw = wpg - 1;                      // consumer warps per group
    // ---- carve shared memory ----
    uint64_t* full = reinterpret_cast<uint64_t*>(smem) + grp * S;         // [G][S]
    uint64_t* empty = reinterpret_cast<uint64_t*>(smem) + (G + grp) * S;  // [G][S]
    uint64_t* bfull = reinterpret_cast<uint64_t*>(smem) + 2 * G * S;      // [1]
    StageHeader* hdr0 = reinterpret_cast<StageHeader*>(smem + ((8 * (2 * G * S + 1) + 15) & ~15));   // [G][S]
    StageHeader* hdr = hdr0 + grp * S;
    unsigned char* cur = reinterpret_cast<unsigned char*>(hdr0 + G * S);
    cur = smem + (((cur - smem) + 127) & ~127);
    const unsigned char* Bs = cur;                                // Bm image (SMEM variants)
    if (BSRC != BSRC_GLOBAL) cur += (p.b_bytes + 127) & ~127;
    const int stage_bytes = stream_stage_bytes(TR, C);
    unsigned char* stage0 = cur + (size_t)grp * S * stage_bytes;
    const int rp_bytes = (TR + 8) * 4, arr_bytes = (C + 8) * 4;

    if (threadIdx.x == 0) {
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
        for (int s = 0; s < G * S; s++) { mbar_init(bars + s, 1); mbar_init(bars + G * S + s, ncw); }
        mbar_init(bfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (grp >= G) return;                         // warps beyond G whole groups (none when blockDim = 32*G*wpg)

    if (warp == 0) {
        // =========================== PRODUCER ===========================
        if (BSRC != BSRC_GLOBAL && lane == 0 && grp == 0) {
            mbar_arrive_expect_tx(bfull, (uint32_t)p.b_bytes);
            // a few large copies; the tx-count of an mbarrier holds up to 2^20-1 bytes
            int off = 0;
            while (off < p.b_bytes) {
                int n = min(p.b_bytes - off, 65536);
                bulk_g2s(const_cast<unsigned char*>(Bs) + off, reinterpret_cast<const unsigned char*>(p.Bm) + off, n, bfull);
                off += n;
            }
        }
        const int nnz_total = __ldg(p.rowptr + p.nrows);
        const int SUB = TR / 32;                  // rows per piece
        const int long_thresh = min(p.long_thresh, C - 4);
        int stage = 0;
        uint32_t ephase = 1;                      // waiting on parity 1 of a fresh barrier passes at once

        // stream rows [rb, re) whose non-zeros [kb, ke) fit a stage
        auto emit = [&](int rb, int re, int kb, int ke) {
            mbar_wait(empty + stage, ephase);
            unsigned char* st = stage0 + (size_t)stage * stage_bytes;
            int* rp_s = reinterpret_cast<int*>(st);
            int* col_s = reinterpret_cast<int*>(st + rp_bytes);
            float* val_s = reinterpret_cast<float*>(st + rp_bytes + arr_bytes);
            uint32_t tx = 0;
            // --- rowPtr slice: elements [rb_al, re] ---
            const int rb_al = rb & ~3;
            {
                const int tot_safe = (p.nrows + 1) & ~3;                 // 16-byte groups fully inside the array
                const int want = ((re + 1 - rb_al) + 3) & ~3;
                const int bulk = max(0, min(want, tot_safe - rb_al));
                const int rem_lo = rb_al + bulk;                         // elements [rem_lo, re] by hand
                if (lane <= re - rem_lo && lane < 4) rp_s[bulk + lane] = __ldg(p.rowptr + rem_lo + lane);
                if (lane == 0 && bulk > 0) bulk_g2s(rp_s, p.rowptr + rb_al, bulk * 4, full + stage);
                tx += bulk * 4;
            }
            // --- columnIndex / values slices: elements [kb_al, ke) ---
            const int kb_al = kb & ~3;
            {
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
            if (lane == 0) {
                StageHeader h;
                h.row_begin = rb; h.nrows = re - rb; h.kbase = kb_al; h.roff = rb - rb_al;
                hdr[stage] = h;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(full + stage, tx);
            if (++stage == S) { stage = 0; ephase ^= 1; }
        };

        // Tiles are claimed two ahead and their row-pointer samples loaded one ahead, so the atomic
        // and the sample loads of the coming tiles are in flight while this tile is being streamed.
        // Tiles come from a global counter (load balance; a static split was measured slower).
        auto claim = [&]() -> int {
            return lane == 0 ? atomicAdd(p.tile_counter, 1) : 0;    // valid in lane 0 only until shuffled
        };
        const int ntiles = (p.nrows + TR - 1) / TR;
        // piece j of a tile covers rows [a + j*SUB, a + (j+1)*SUB) clipped to the tile
        auto sample = [&](int t, int& r_lo, int& r_hi, int& s_lo, int& s_hi) {
            const long long a_ll = t < ntiles ? (long long)t * TR : (long long)p.nrows;
            const int a = a_ll < p.nrows ? (int)a_ll : p.nrows;
            const int tile_end = min(a + TR, p.nrows);
            r_lo = min(a + lane * SUB, tile_end);
            r_hi = min(r_lo + SUB, tile_end);
            s_lo = __ldg(p.rowptr + r_lo);
            s_hi = __ldg(p.rowptr + r_hi);
        };
        int t_cur = __shfl_sync(0xffffffffu, claim(), 0);
        int n_rlo, n_rhi, n_slo, n_shi;
        sample(t_cur, n_rlo, n_rhi, n_slo, n_shi);
        int t_next_raw = claim();
        for (;;) {
            if (t_cur >= ntiles) break;
            const int a = t_cur * TR;
            const int tile_end = min(a + TR, p.nrows);
            const int r_lo = n_rlo, r_hi = n_rhi, s_lo = n_slo, s_hi = n_shi;
            // next tile: its claim was issued an iteration ago; issue its samples and the claim after it
            t_cur = __shfl_sync(0xffffffffu, t_next_raw, 0);
            sample(t_cur, n_rlo, n_rhi, n_slo, n_shi);
            t_next_raw = claim();
            int piece = 0;
            while (piece < 32) {
                const int pr_lo = __shfl_sync(0xffffffffu, r_lo, piece);
                if (pr_lo >= tile_end) break;
                const int kb = __shfl_sync(0xffffffffu, s_lo, piece);
                // pieces piece..e fit one stage if their span (from the aligned-down start) does,
                // and a piece can only hide a long row if its own span exceeds the threshold
                const bool fits = lane >= piece && (s_hi - (kb & ~3)) <= C && (s_hi - s_lo) <= long_thresh;
                const unsigned nofit = ~__ballot_sync(0xffffffffu, fits) & (0xffffffffu << piece);
                const int e = nofit ? (__ffs(nofit) - 1) : 32;      // first piece that does not fit
                if (e > piece) {
                    const int re = __shfl_sync(0xffffffffu, r_hi, e - 1);
                    const int ke = __shfl_sync(0xffffffffu, s_hi, e - 1);
                    emit(pr_lo, re, kb, ke);
                    piece = e;
                    continue;
                }
                // slow path: this piece alone overflows a stage or may hold a long row -> row by row
                const int pr_hi = __shfl_sync(0xffffffffu, r_hi, piece);
                int run_rb = pr_lo, run_kb = kb;
                for (int r = pr_lo; r < pr_hi; r++) {
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
                if (pr_hi > run_rb) emit(run_rb, pr_hi, run_kb, __ldg(p.rowptr + pr_hi));
                piece++;
            }
        }
        // sentinel: no more work
        mbar_wait(empty + stage, ephase);
        if (lane == 0) {
            StageHeader h;
            h.row_begin = 0; h.nrows = -1; h.kbase = 0; h.roff = 0;
            hdr[stage] = h;
            mbar_arrive(full + stage);
        }
        return;
    }

    // =========================== CONSUMERS ===========================
    const int cw = warp - 1;
    const int g = lane / LPR, l = lane % LPR;
    const int P4 = p.P4;
    const uint32_t rowbytes = EXACT ? (uint32_t)(LPR * NV * 16) : (uint32_t)P4 * 16u;
    // byte address of this lane's first chunk inside a Bm row
    const unsigned char* bs_lane = Bs;
    const char* bg_lane = nullptr;
    uint32_t bs_rowbytes = 0;
    if (BSRC == BSRC_GLOBAL) {
        bg_lane = reinterpret_cast<const char*>(p.Bm + l);
    } else {
        bs_lane = Bs + l * 16;
        bs_rowbytes = rowbytes;
    }
    if (BSRC != BSRC_GLOBAL) mbar_wait(bfull, 0);

    const float inv_block = PEER ? 1.0f / (float)p.peer_block : 0.f;
    auto gather = [&](int c, int v) -> float4 {
        if (BSRC == BSRC_GLOBAL && PEER) {
            // owner rank of row c: float estimate of c / block, corrected to the exact quotient
            int r = (int)((float)c * inv_block);
            int local = c - r * p.peer_block;
            if (local < 0) { r--; local += p.peer_block; }
            if (local >= p.peer_block) { r++; local -= p.peer_block; }
            const char* base = p.peer_base[r] + (size_t)l * 16;
            float4 out4;      // plain (coherent) load: peer memory is not read-only for the kernel's lifetime
            const float4* ptr = reinterpret_cast<const float4*>(base + (size_t)(unsigned)local * rowbytes) + v * LPR;
            asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(out4.x), "=f"(out4.y), "=f"(out4.z), "=f"(out4.w) : "l"(ptr));
            return out4;
        } else if (BSRC == BSRC_GLOBAL) {
            return __ldg(reinterpret_cast<const float4*>(bg_lane + (size_t)(unsigned)c * rowbytes) + v * LPR);
        } else {
            return *reinterpret_cast<const float4*>(bs_lane + (uint32_t)c * bs_rowbytes + v * (LPR * 16));
        }
    };

    int stage = 0;
    uint32_t fphase = 0;
    for (;;) {
        mbar_wait(full + stage, fphase);
        const StageHeader h = hdr[stage];
        if (h.nrows < 0) break;
        const unsigned char* st = stage0 + (size_t)stage * stage_bytes;
        const int* rp_s = reinterpret_cast<const int*>(st) + h.roff;
        // col_s[k] / val_s[k] for a GLOBAL non-zero index k
        const int* col_k = reinterpret_cast<const int*>(st + rp_bytes) - h.kbase;
        const float* val_k = reinterpret_cast<const float*>(st + rp_bytes + arr_bytes) - h.kbase;

        for (int i = cw * RPW + g; i < h.nrows; i += ncw * RPW) {
            const int beg = rp_s[i], end = rp_s[i + 1];
            float4 acc[NV];
#pragma unroll
            for (int v = 0; v < NV; v++) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

            for (int kb = beg & ~3; kb < end; kb += 4) {
                const int4 c4 = *reinterpret_cast<const int4*>(col_k + kb);
                const float4 a4 = *reinterpret_cast<const float4*>(val_k + kb);
                const int cs[4] = {c4.x, c4.y, c4.z, c4.w};
                float as[4] = {a4.x, a4.y, a4.z, a4.w};
                const int lo = beg - kb, hi = end - kb;      // slot s is this row's iff lo <= s < hi
                bool live[4];
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    live[s] = s >= lo && s < hi;
                    if (QM == QM_GCN && p.q.quant) {
                        as[s] = stream_adj_code(as[s], p.q);
                        live[s] = live[s] && as[s] != 0.f;       // a zero code is a pruned edge: no gather, no term
                    }
                }
                if (NV <= 2) {
                    // all gathers of the block are issued before any of them is consumed
                    float4 b[4][NV];
#pragma unroll
                    for (int s = 0; s < 4; s++) {
#pragma unroll
                        for (int v = 0; v < NV; v++) {
                            b[s][v] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (live[s] && (EXACT || v * LPR + l < P4)) b[s][v] = gather(cs[s], v);
                        }
                    }
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        if (QM == QM_GCN) {
                            if (live[s]) {
#pragma unroll
                                for (int v = 0; v < NV; v++) mul_add4s(acc[v], as[s], b[s][v]);
                            }
                        } else {
                            const float a = live[s] ? as[s] : 0.f;
#pragma unroll
                            for (int v = 0; v < NV; v++) fma4s(acc[v], a, b[s][v]);
                        }
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        if (live[s]) {
                            float4 b[NV];
#pragma unroll
                            for (int v = 0; v < NV; v++) {
                                b[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (EXACT || v * LPR + l < P4) b[v] = gather(cs[s], v);
                            }
#pragma unroll
                            for (int v = 0; v < NV; v++) { if (QM == QM_GCN) mul_add4s(acc[v], as[s], b[v]); else fma4s(acc[v], as[s], b[v]); }
                        }
                    }
                }
            }
            float4* orow = p.out + (size_t)(h.row_begin + i) * (EXACT ? LPR * NV : P4);
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const int q = v * LPR + l;
                if (EXACT || q < P4) {
                    float4 r = acc[v];
                    if (p.accumulate) { const float4 o = orow[q]; r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
                    if (p.relu) r = relu4(r);     // val = (acc > 0 || relu == 0) ? acc : 0   (K:2586-2590)
                    if (QM == QM_GCN && p.q.quant) {
                        r.x = __fmul_rn(r.x, p.q.deq_o); r.y = __fmul_rn(r.y, p.q.deq_o);
                        r.z = __fmul_rn(r.z, p.q.deq_o); r.w = __fmul_rn(r.w, p.q.deq_o);
                    }
                    if (p.streaming_store) __stcs(orow + q, r); else orow[q] = r;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);
        if (++stage == S) { stage = 0; fphase ^= 1; }
    }
}

}  // namespace sgrace
