// Prototype for DESIGN.md section 9 item 1 (NOT part of the library): ADJ on a block-diagonal batch with the
// panel's window of XW rows held in shared memory.
//
//   ./adj_window [graphs=1024] [nodes=2708] [P=16] [reps=20]
//
// Builds `graphs` random blocks (degree ~ Cora: mean 4.9, one hub of 168 per block, self-loops), runs
//   (a) gather_kernel : LPR lanes x float4 per row, XW rows gathered from global memory (the row-strided baseline)
//   (b) window_kernel : one CTA per panel (= block); the panel's XW rows are bulk-copied (TMA 1-D) into shared
//                       memory, then the same row loop gathers from shared memory, column indices rebased
// checks (b) == (a) bit for bit (same k-ascending FMA order) and prints the time of each and the algorithmic GB/s
// (SURVEY 8d ADJ bytes).  Panels come from the generator here; the library version would find them with the
// prefix-max scan described in DESIGN.md.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(su32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(su32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(d)), "l"(s), "r"(n), "r"(su32(b)) : "memory");
}
__device__ __forceinline__ void fma4(float4& a, float s, const float4& b) {
    a.x = fmaf(s, b.x, a.x); a.y = fmaf(s, b.y, a.y); a.z = fmaf(s, b.z, a.z); a.w = fmaf(s, b.w, a.w);
}

// rows [r0, r1) of the CSR matrix, LPR lanes per row; `Bsrc` is either global XW or the shared window
// (then `cbase` = first row of the window).  Four gathers in flight, k ascending.
template <int LPR, bool SMEM>
__device__ __forceinline__ void rows_loop(const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ val,
                                          const float4* Bsrc, int cbase, float4* __restrict__ out, int r0, int r1, int P4, int relu,
                                          int warp, int nwarps) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, g = lane / LPR, l = lane % LPR;
    for (int row = r0 + warp * RPW + g; row < r1; row += nwarps * RPW) {
        const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = beg; k < end; k += 4) {
            int c[4]; float a[4]; float4 b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const bool in = k + i < end;
                c[i] = in ? __ldg(col + k + i) - cbase : 0;
                a[i] = in ? __ldg(val + k + i) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k + i < end && l < P4) b[i] = SMEM ? Bsrc[(size_t)c[i] * P4 + l] : __ldg(Bsrc + (size_t)c[i] * P4 + l);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) fma4(acc, a[i], b[i]);
        }
        if (l < P4) {
            if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
            __stcs(out + (size_t)row * P4 + l, acc);
        }
    }
}

template <int LPR>
__global__ void __launch_bounds__(256) gather_kernel(const int* rowptr, const int* col, const float* val, const float4* XW, float4* out,
                                                     int nrows, int P4, int relu) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    rows_loop<LPR, false>(rowptr, col, val, XW, 0, out, 0, nrows, P4, relu, warp, nwarps);
}

struct Panel { int r0, r1, w0, wrows; };

// persistent: CTAs claim panels from a counter.  Window single-buffered: load | barrier | compute | barrier.
template <int LPR>
__global__ void __launch_bounds__(1024, 1) window_kernel(const int* rowptr, const int* col, const float* val, const float4* XW, float4* out,
                                                         const Panel* panels, int npanels, int P4, int relu, int* counter) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = (uint64_t*)sm;
    int* claim = (int*)(sm + 16);
    float4* win = (float4*)(sm + 128);
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;
    for (;;) {
        if (threadIdx.x == 0) *claim = atomicAdd(counter, 1);
        __syncthreads();                         // also: everybody has finished reading the previous window
        const int pi = *claim;
        if (pi >= npanels) break;
        const Panel p = panels[pi];
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)p.wrows * P4 * 16;
            mbar_expect(full, bytes);
            const char* src = (const char*)(XW + (size_t)p.w0 * P4);
            for (uint32_t off = 0; off < bytes; off += 32768)
                bulk((char*)win + off, src + off, bytes - off < 32768 ? bytes - off : 32768, full);
        }
        mbar_wait(full, phase);
        phase ^= 1;
        rows_loop<LPR, true>(rowptr, col, val, win, p.w0, out, p.r0, p.r1, P4, relu, warp, nwarps);
        __syncthreads();                         // the claim word is rewritten after this point
    }
}

int main(int argc, char** argv) {
    const int graphs = argc > 1 ? atoi(argv[1]) : 1024, n1 = argc > 2 ? atoi(argv[2]) : 2708, P = argc > 3 ? atoi(argv[3]) : 16;
    const int reps = argc > 4 ? atoi(argv[4]) : 20;
    const int P4 = P / 4;
    if (P != 16) { printf("this prototype instantiates P = 16 only\n"); return 1; }
    const long long N = (long long)graphs * n1;
    std::mt19937 rng(1);
    // one block: symmetric pattern, self-loops, degrees ~ geometric with mean ~3.9 + self = 4.9, one hub with 168 neighbours
    std::vector<std::vector<int>> nb(n1);
    for (int i = 0; i < n1; i++) nb[i].push_back(i);
    std::geometric_distribution<int> gd(0.34);
    for (int i = 0; i < n1; i++) {
        const int d = std::min(gd(rng), 30);
        for (int j = 0; j < d; j++) { const int t = rng() % n1; if (t != i) { nb[i].push_back(t); nb[t].push_back(i); } }
    }
    for (int j = 0; j < 168; j++) { const int t = 1 + rng() % (n1 - 1); nb[0].push_back(t); nb[t].push_back(0); }
    std::vector<int> rp1(n1 + 1, 0), ci1;
    for (int i = 0; i < n1; i++) {
        std::sort(nb[i].begin(), nb[i].end());
        nb[i].erase(std::unique(nb[i].begin(), nb[i].end()), nb[i].end());
        rp1[i + 1] = rp1[i] + (int)nb[i].size();
        ci1.insert(ci1.end(), nb[i].begin(), nb[i].end());
    }
    const long long nnz1 = ci1.size(), nnz = nnz1 * graphs;
    if (nnz >= (1ll << 31)) { printf("too large\n"); return 1; }
    std::vector<int> rp(N + 1), ci(nnz);
    std::vector<float> va(nnz), xw((size_t)N * P);
    std::vector<Panel> panels(graphs);
    for (int gi = 0; gi < graphs; gi++) {
        for (int i = 0; i < n1; i++) rp[(size_t)gi * n1 + i] = (int)(gi * nnz1 + rp1[i]);
        for (long long k = 0; k < nnz1; k++) { ci[gi * nnz1 + k] = ci1[k] + gi * n1; va[gi * nnz1 + k] = 1.0f / (1 + (rng() % 7)); }
        panels[gi] = Panel{gi * n1, (gi + 1) * n1, gi * n1, n1};
    }
    rp[N] = (int)nnz;
    for (auto& v : xw) v = (float)((int)(rng() % 2001) - 1000) * 1e-3f;
    printf("graphs %d x %d nodes, nnz %lld (%.2f per row), P %d, window %.1f KB\n", graphs, n1, nnz, (double)nnz / N, P, n1 * P * 4 / 1024.0);

    int *d_rp, *d_ci, *d_counter; float *d_va, *d_xw, *d_o1, *d_o2; Panel* d_panels;
    CK(cudaMalloc(&d_rp, (N + 1) * 4)); CK(cudaMalloc(&d_ci, nnz * 4)); CK(cudaMalloc(&d_va, nnz * 4));
    CK(cudaMalloc(&d_xw, (size_t)N * P * 4)); CK(cudaMalloc(&d_o1, (size_t)N * P * 4)); CK(cudaMalloc(&d_o2, (size_t)N * P * 4));
    CK(cudaMalloc(&d_panels, graphs * sizeof(Panel))); CK(cudaMalloc(&d_counter, 4));
    CK(cudaMemcpy(d_rp, rp.data(), (N + 1) * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_ci, ci.data(), nnz * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_va, va.data(), nnz * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_xw, xw.data(), (size_t)N * P * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_panels, panels.data(), graphs * sizeof(Panel), cudaMemcpyHostToDevice));

    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t smem = 128 + (size_t)n1 * P * 4;
    int optin = 0; CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0));
    if ((int)smem > optin) { printf("window of %zu bytes does not fit %d bytes of shared memory\n", smem, optin); return 1; }
    CK(cudaFuncSetAttribute(window_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double bytes = (N + 1) * 4.0 + nnz * 8.0 + 2.0 * N * P * 4;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    const int ggrid = sms * 8;
    for (int w = 0; w < 3; w++) gather_kernel<4><<<ggrid, 256>>>(d_rp, d_ci, d_va, (const float4*)d_xw, (float4*)d_o1, (int)N, P4, 1);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) gather_kernel<4><<<ggrid, 256>>>(d_rp, d_ci, d_va, (const float4*)d_xw, (float4*)d_o1, (int)N, P4, 1);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("gather (global)  %.4f ms  %.0f GB/s algorithmic  %.2f GTEPS\n", ms / reps, bytes / (ms / reps) / 1e6, nnz / (ms / reps) / 1e6);
    for (int w = 0; w < 3; w++) {
        CK(cudaMemsetAsync(d_counter, 0, 4));
        window_kernel<4><<<sms, 1024, smem>>>(d_rp, d_ci, d_va, (const float4*)d_xw, (float4*)d_o2, d_panels, graphs, P4, 1, d_counter);
    }
    CK(cudaDeviceSynchronize());
    float tot = 0.f;
    for (int r = 0; r < reps; r++) {
        CK(cudaMemsetAsync(d_counter, 0, 4));
        CK(cudaEventRecord(e0));
        window_kernel<4><<<sms, 1024, smem>>>(d_rp, d_ci, d_va, (const float4*)d_xw, (float4*)d_o2, d_panels, graphs, P4, 1, d_counter);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        tot += ms;
    }
    printf("window (smem)    %.4f ms  %.0f GB/s algorithmic  %.2f GTEPS\n", tot / reps, bytes / (tot / reps) / 1e6, nnz / (tot / reps) / 1e6);
    std::vector<float> o1((size_t)N * P), o2((size_t)N * P);
    CK(cudaMemcpy(o1.data(), d_o1, o1.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o2.data(), d_o2, o2.size() * 4, cudaMemcpyDeviceToHost));
    printf("results %s\n", memcmp(o1.data(), o2.data(), o1.size() * 4) == 0 ? "bit-equal" : "DIFFER");
    return 0;
}
