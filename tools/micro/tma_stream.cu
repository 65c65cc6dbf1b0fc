// Micro-benchmark: how fast can one producer warp per CTA stream a large buffer from HBM into a
// shared-memory ring with cp.async.bulk + mbarriers?  (sizing evidence for sgrace_spmm_stream.cuh)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(su32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(su32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(su32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(d)), "l"(s), "r"(n), "r"(su32(b)) : "memory");
}

// each CTA streams chunks blockIdx.x, blockIdx.x + grid, ... of `stage_bytes`; `ncopy` copies per stage
__global__ void __launch_bounds__(1024, 1) stream_kernel(const char* src, size_t total, int stage_bytes, int S, int ncopy, int ncw, unsigned* sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = (uint64_t*)sm;
    uint64_t* empty = full + S;
    unsigned char* st0 = sm + 256;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(full + s, 1); mbar_init(empty + s, ncw); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nchunks = total / stage_bytes;
    if (warp == 0) {
        int stage = 0; uint32_t ph = 1;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            mbar_wait(empty + stage, ph);
            if (lane == 0) {
                mbar_expect(full + stage, stage_bytes);
                const int per = stage_bytes / ncopy;
                for (int i = 0; i < ncopy; i++)
                    bulk(st0 + (size_t)stage * stage_bytes + i * per, src + c * stage_bytes + i * per, per, full + stage);
            }
            __syncwarp();
            if (++stage == S) { stage = 0; ph ^= 1; }
        }
    } else if (warp <= ncw) {
        int stage = 0; uint32_t ph = 0; unsigned acc = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            mbar_wait(full + stage, ph);
            acc += *(volatile unsigned*)(st0 + (size_t)stage * stage_bytes + lane * 4);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + stage);
            if (++stage == S) { stage = 0; ph ^= 1; }
        }
        if (acc == 0x12345678u) *sink = acc;
    }
}

int main() {
    const size_t total = (size_t)1 << 30;
    char* src; unsigned* sink;
    cudaMalloc(&src, total); cudaMalloc(&sink, 4);
    cudaMemset(src, 1, total);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int cfgs[][5] = {  // stage_bytes, S, ncopy, consumer warps, ctas per sm
        {8192, 4, 1, 1, 1}, {16384, 4, 1, 1, 1}, {16384, 6, 1, 1, 1}, {16384, 6, 3, 1, 1}, {16384, 6, 3, 31, 1},
        {16384, 12, 1, 1, 1}, {32768, 6, 1, 1, 1}, {32768, 6, 3, 31, 1}, {65536, 3, 1, 1, 1}, {8192, 16, 1, 1, 1},
        {8192, 6, 1, 1, 2}, {16384, 4, 1, 1, 3}, {4096, 32, 1, 1, 1}, {16384, 12, 3, 31, 1}};
    for (auto& c : cfgs) {
        const int sb = c[0], S = c[1], nc = c[2], ncw = c[3], per_sm = c[4];
        const size_t smem = 256 + (size_t)S * sb;
        const int threads = 32 * (1 + ncw);
        const int grid = 148 * per_sm;
        for (int it = 0; it < 2; it++) {
            cudaEventRecord(e0);
            stream_kernel<<<grid, threads, smem>>>(src, total, sb, S, nc, ncw, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        printf("stage %6d B x S=%2d copies/stage %d consumers %2d ctas/sm %d : %.3f ms  %.0f GB/s  (%s)\n", sb, S, nc, ncw, per_sm, ms, total / ms / 1e6, cudaGetErrorString(err));
    }
    return 0;
}
