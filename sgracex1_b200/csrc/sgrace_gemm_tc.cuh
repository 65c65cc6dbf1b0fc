// sgrace_gemm_tc.cuh -- tensor-core path for the one real contraction on the hot path: dense FEA
// (gemm_mode = 1, kernelMatrixmult_all.cpp:847-865 / 985-1013) with a wide hidden layer, e.g. the
// ogbn-products shape XW[N x 256] = X[N x 100] . W[100 x 256].
//
// float32 in, float32 out, <= 1e-5 relative: plain TF32 (10-bit mantissa) is not enough, so every
// operand is split on the fly into hi = its TF32 truncation and lo = the exact remainder and the
// product is formed as  hi.hi + lo.hi + hi.lo  (3xTF32, error ~2^-21 per product), accumulated in
// float32 in tensor memory.
//
// Layouts need no transposition: X is [rows][K] K-major and the B buffer holds W transposed,
// [P][K] K-major (kernelMatrixmult_all.cpp:3038-3051) -- exactly the two K-major operands of
// tcgen05.mma.  One CTA owns a 128-column (or 64-column) slice of W for the whole launch (hi and lo images resident
// in shared memory) and walks 128-row tiles of X:
//   warp 0      one lane: tcgen05.mma issuer
//   warp 1      TMEM allocation
//   warp 2      one lane: TMA producer (cp.async.bulk.tensor 2-D, 128-byte swizzle, zero fill past
//               K and past the last row)
//   warp 3      spare (keeps the epilogue warps at warp%4 == TMEM lane quarter)
//   warps 4-7   epilogue: tcgen05.ld the 128 x 128 accumulator, 128-bit conflict-free stores into a
//               swizzled staging tile, one TMA store (cp.async.bulk.tensor) per 32 x 32 block
//   warps 8-11  splitter: rewrite each landed K-chunk in place as hi, write lo beside it
// Ring of K-chunks x {hi, lo} (as many as fit beside the resident W slice; the chunk is 32 floats =
// 128-byte swizzle, or 16 floats = 64-byte swizzle when that buys a deeper ring); accumulator
// double-buffered in TMEM so the epilogue of tile t overlaps the MMAs of tile t+1.  The CTAs that
// share a row tile (column slices 0..P/128-1) walk the tiles in the same order, so X is fetched from
// HBM once and re-read from L2.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgrace {
namespace tc {

constexpr int BM = 128;          // rows per tile (UMMA M)
constexpr int MAX_RING = 8;      // K-chunks in flight (as many as fit beside the resident W slice)
constexpr int THREADS = 384;

// small fixed part of shared memory; the 1024-byte aligned tiles follow it:
//   stage[4] (4 KB each, epilogue)   b_hi[kc], b_lo[kc]   a_hi[ring], a_lo[ring]
struct Ctrl {
    uint64_t a_full[MAX_RING], a_split[MAX_RING], a_empty[MAX_RING];
    uint64_t acc_full[2], acc_empty[2];
    uint64_t b_full;
    uint32_t tmem_base;
};
constexpr int CTRL_BYTES = 1024;
constexpr int STAGE_BYTES = 4 * 4096;

inline size_t smem_bytes(int bn, int bk, int kc, int ring) {
    return 1024 + CTRL_BYTES + STAGE_BYTES + (size_t)2 * kc * (bn * bk * 4) + (size_t)2 * ring * (BM * bk * 4);
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void bar_arrive(uint64_t* b) { asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void bar_expect(uint64_t* b, uint32_t n) { asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)),
                 "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(s32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
// K-major operand tile whose rows are one swizzle span (128 or 64 bytes): 8-row atoms, SBO = 8 rows
template <int BK>
__device__ __forceinline__ uint64_t umma_desc(const void* tile, int k_byte_off) {
    constexpr int ROWB = BK * 4;
    const uint32_t addr = s32(tile) + k_byte_off;
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8 * ROWB) >> 4) << 32;          // stride byte offset
    d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
    d |= (uint64_t)(ROWB == 128 ? 2 : 4) << 61;      // SWIZZLE_128B / SWIZZLE_64B
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(da),
                 "l"(db), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
        "%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// hi = value rounded to the nearest TF32 number (13 low mantissa bits zero), lo = value - hi (exact,
// |lo| <= 2^-11 |value|); Inf/NaN keep their exponent and propagate through hi
__device__ __forceinline__ float tf32_hi(float v) {
    const uint32_t b = __float_as_uint(v);
    return __uint_as_float(((b & 0x7F800000u) == 0x7F800000u ? b : b + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split4(float4 v, float4& hi, float4& lo) {
    hi.x = tf32_hi(v.x); lo.x = v.x - hi.x;
    hi.y = tf32_hi(v.y); lo.y = v.y - hi.y;
    hi.z = tf32_hi(v.z); lo.z = v.z - hi.z;
    hi.w = tf32_hi(v.w); lo.w = v.w - hi.w;
}

// BN = columns of W per CTA (UMMA N): 128, or 64 for hidden widths that are only a multiple of 64
template <int BK, int BN>
__global__ void __launch_bounds__(THREADS, 1)
fea_dense_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_o, int K, int P, int tiles, int ctas_per_slice, int kc, int ring,
                    int relu) {
    constexpr int A_TILE = BM * BK * 4, B_TILE = BN * BK * 4;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Ctrl& s = *reinterpret_cast<Ctrl*>(base);
    unsigned char* stage0 = base + CTRL_BYTES;
    unsigned char* b_hi0 = stage0 + STAGE_BYTES;
    unsigned char* b_lo0 = b_hi0 + (size_t)kc * B_TILE;
    unsigned char* a_hi0 = b_lo0 + (size_t)kc * B_TILE;
    unsigned char* a_lo0 = a_hi0 + (size_t)ring * A_TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nslice = P / BN;
    const int slice = blockIdx.x % nslice;          // which BN columns of W this CTA owns
    const int first_tile = blockIdx.x / nslice;     // CTAs sharing a tile are neighbours: same tile order
    const int RING = ring;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; i++) { bar_init(&s.a_full[i], 1); bar_init(&s.a_split[i], 128); bar_init(&s.a_empty[i], 1); }
        for (int i = 0; i < 2; i++) { bar_init(&s.acc_full[i], 1); bar_init(&s.acc_empty[i], 128); }
        bar_init(&s.b_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {    // TMEM: 2 accumulators x BN columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&s.tmem_base)), "r"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;

    if (warp == 2) {
        if (lane == 0) {
            // ---- TMA producer: the W slice once (kc chunks of BN rows), then the X chunks ----
            bar_expect(&s.b_full, (uint32_t)(kc * B_TILE));
            for (int c = 0; c < kc; c++) tma_2d(b_hi0 + (size_t)c * B_TILE, &map_b, c * BK, slice * BN, &s.b_full);
            int slot = 0;
            uint32_t phase = 1;                                       // a fresh "empty" barrier passes at parity 1
            for (int t = first_tile; t < tiles; t += ctas_per_slice) {
                for (int c = 0; c < kc; c++) {
                    bar_wait(&s.a_empty[slot], phase);
                    bar_expect(&s.a_full[slot], A_TILE);
                    tma_2d(a_hi0 + (size_t)slot * A_TILE, &map_x, c * BK, t * BM, &s.a_full[slot]);
                    if (++slot == RING) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 0) {
        if (lane == 0) {
            // ---- MMA issuer ----
            int slot = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = first_tile; t < tiles; t += ctas_per_slice, it++) {
                const int ab = it & 1;
                bar_wait(&s.acc_empty[ab], ((it >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem + ab * BN;
                for (int c = 0; c < kc; c++) {
                    bar_wait(&s.a_split[slot], phase);                 // hi/lo of this chunk are in place
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int ksteps = min(BK, K - c * BK + 7) / 8;    // 8 floats of K per MMA; past K is zero-filled
                    for (int k = 0; k < ksteps; k++) {
                        const uint64_t ah = umma_desc<BK>(a_hi0 + (size_t)slot * A_TILE, k * 32);
                        const uint64_t al = umma_desc<BK>(a_lo0 + (size_t)slot * A_TILE, k * 32);
                        const uint64_t bh = umma_desc<BK>(b_hi0 + (size_t)c * B_TILE, k * 32);
                        const uint64_t bl = umma_desc<BK>(b_lo0 + (size_t)c * B_TILE, k * 32);
                        umma_tf32(d, ah, bh, idesc, (c | k) != 0);
                        umma_tf32(d, al, bh, idesc, 1);
                        umma_tf32(d, ah, bl, idesc, 1);
                    }
                    umma_commit(&s.a_empty[slot]);                     // frees the ring slot when these MMAs retire
                    if (c == kc - 1) umma_commit(&s.acc_full[ab]);
                    if (++slot == RING) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 8) {
        // ---- splitter: 128 threads, chunk by chunk in ring order ----
        const int tid = threadIdx.x - 256;
        {   // W slice first
            bar_wait(&s.b_full, 0);
            float4* hi = reinterpret_cast<float4*>(b_hi0);
            float4* lo = reinterpret_cast<float4*>(b_lo0);
            for (int i = tid; i < kc * B_TILE / 16; i += 128) { float4 h, l; split4(hi[i], h, l); hi[i] = h; lo[i] = l; }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        int slot = 0;
        uint32_t phase = 0;
        for (int t = first_tile; t < tiles; t += ctas_per_slice) {
            for (int c = 0; c < kc; c++) {
                bar_wait(&s.a_full[slot], phase);
                float4* hi = reinterpret_cast<float4*>(a_hi0 + (size_t)slot * A_TILE);
                float4* lo = reinterpret_cast<float4*>(a_lo0 + (size_t)slot * A_TILE);
#pragma unroll
                for (int i = 0; i < A_TILE / 16 / 128; i++) {
                    float4 h, l;
                    split4(hi[tid + i * 128], h, l);
                    hi[tid + i * 128] = h;
                    lo[tid + i * 128] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (MMA) reads
                bar_arrive(&s.a_split[slot]);
                if (++slot == RING) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ---- epilogue: warp w reads TMEM lanes 32*(w%4).. ; thread = one output row ----
        const int q = warp & 3;
        unsigned char* st = stage0 + q * 4096;          // 32 rows x 128 bytes, 128-byte swizzled like the TMA box
        int it = 0;
        for (int t = first_tile; t < tiles; t += ctas_per_slice, it++) {
            const int ab = it & 1;
            bar_wait(&s.acc_full[ab], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int blk = 0; blk < BN / 32; blk++) {
                uint32_t r[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * BN + blk * 32, r);
                if (relu) {     // val = (acc > 0 || relu == 0) ? acc : 0  (kernelMatrixmult_all.cpp:2586-2590), aggregate-first order
#pragma unroll
                    for (int j = 0; j < 32; j++) r[j] = __uint_as_float(r[j]) > 0.f ? r[j] : 0u;
                }
                // the previous TMA store must have finished reading the staging tile
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int c16 = 0; c16 < 8; c16++) {     // row = lane; 16-byte chunk c16 lands at c16 ^ (row & 7): conflict-free
                    const float4 v = make_float4(__uint_as_float(r[c16 * 4]), __uint_as_float(r[c16 * 4 + 1]),
                                                 __uint_as_float(r[c16 * 4 + 2]), __uint_as_float(r[c16 * 4 + 3]));
                    *reinterpret_cast<float4*>(st + lane * 128 + ((c16 ^ (lane & 7)) << 4)) = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&map_o, st, slice * BN + blk * 32, t * BM + q * 32);   // rows past N are clipped by the map
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(&s.acc_empty[ab]);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // stores complete before exit
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * BN));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows][cols] float32 row-major, boxes of box_rows x box_cols, swizzle span = box_cols * 4 bytes
// (128 or 64), zero fill out of bounds on loads, clipping on stores
inline int make_map(CUtensorMap* map, const float* base, int rows, int cols, int box_rows, int box_cols) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return -1;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

}  // namespace tc

// X: N x M row-major; B: W transposed, P x M row-major; out: N x P row-major
inline bool fea_dense_tc_supported(int N, int M, int P) {
    return N >= tc::BM && M % 4 == 0 && M >= 32 && M <= 128 && P % 64 == 0 && P >= 64;
}

// returns 0 on success, -100 when the shape / pointers are not eligible (caller falls back), other
// negatives on CUDA errors
inline int fea_dense_tc_launch(const float* X, const float* B, float* out, int N, int M, int P, int num_sms, cudaStream_t stream,
                               int force_bk = 0, int relu = 0) {
    if (!fea_dense_tc_supported(N, M, P)) return -100;
    if ((((uintptr_t)X) | ((uintptr_t)B) | ((uintptr_t)out)) & 15) return -100;
    int max_optin = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int bn = P % 128 == 0 ? 128 : 64;
    // chunk width: 32 floats (128-byte swizzle) unless 16 floats lets at least a whole tile of X be in flight
    auto ring_for = [&](int bk) {
        const int kc = (M + bk - 1) / bk;
        int ring = tc::MAX_RING;
        while (ring > 1 && tc::smem_bytes(bn, bk, kc, ring) > (size_t)max_optin) ring--;
        return ring;
    };
    int bk = 32;
    if (ring_for(32) * 32 < M && ring_for(16) * 16 > ring_for(32) * 32) bk = 16;
    if (force_bk == 16 || force_bk == 32) bk = force_bk;
    const int kc = (M + bk - 1) / bk, ring = ring_for(bk);
    if (ring < 2) return -100;
    CUtensorMap mx, mb, mo;
    if (tc::make_map(&mx, X, N, M, tc::BM, bk) != 0 || tc::make_map(&mb, B, P, M, bn, bk) != 0 ||
        tc::make_map(&mo, out, N, P, 32, 32) != 0)
        return -100;
    const int tiles = (N + tc::BM - 1) / tc::BM;
    const int nslice = P / bn;
    int per_slice = num_sms / nslice;
    if (per_slice < 1) return -100;
    if (per_slice > tiles) per_slice = tiles;
    const size_t smem = tc::smem_bytes(bn, bk, kc, ring);
#define SGRACE_TC_LAUNCH(BKV, BNV)                                                                                              \
    do {                                                                                                                        \
        if (cudaFuncSetAttribute(tc::fea_dense_tc_kernel<BKV, BNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=   \
            cudaSuccess)                                                                                                        \
            return -2;                                                                                                          \
        tc::fea_dense_tc_kernel<BKV, BNV><<<per_slice * nslice, tc::THREADS, smem, stream>>>(mx, mb, mo, M, P, tiles, per_slice, \
                                                                                             kc, ring, relu);                   \
    } while (0)
    if (bk == 32 && bn == 128) SGRACE_TC_LAUNCH(32, 128);
    else if (bk == 16 && bn == 128) SGRACE_TC_LAUNCH(16, 128);
    else if (bk == 32) SGRACE_TC_LAUNCH(32, 64);
    else SGRACE_TC_LAUNCH(16, 64);
#undef SGRACE_TC_LAUNCH
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

}  // namespace sgrace
