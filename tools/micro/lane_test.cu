// Stand-alone check + timing of adj_window_lane.cuh (experiment: ADJ with the XW window of a panel in shared memory):
//
//   ./lane_test [graphs=1024] [nodes=2708] [M=1433] [reps=20]
//   env: NCW_ADJ (consumer warps), C_ADJ, S_ADJ, QUANTUM, CAP (window rows), LONG (deferred-row threshold),
//        SMALL=1 (ragged edge-case matrices)
//
// Builds `graphs` Cora-shape blocks (block-diagonal adjacency), runs the windowed ADJ kernel (XW window in shared
// memory, panels from the planner) and compares it bit for bit with a row-per-thread kernel that does the same
// k-ascending FMAs.
#include "adj_window_lane.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <vector>

using namespace sgrace;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static int env_int(const char* n, int d) { const char* v = getenv(n); return (v && *v) ? atoi(v) : d; }

__global__ void ref_kernel(const int* rp, const int* ci, const float* va, const float4* Bm, float4* out, int nrows, int relu) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = t >> 2, q = t & 3;
    if (r >= nrows) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = rp[r]; k < rp[r + 1]; k++) fma4s(acc, va[k], Bm[(size_t)ci[k] * 4 + q]);
    if (relu) acc = relu4(acc);
    out[(size_t)r * 4 + q] = acc;
}
// rows the lane kernel deferred
__global__ void ref_list_kernel(const int* rp, const int* ci, const float* va, const float4* Bm, float4* out, const int* list, const int* count, int relu) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 2, q = t & 3;
    if (i >= *count) return;
    const int r = list[i];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = rp[r]; k < rp[r + 1]; k++) fma4s(acc, va[k], Bm[(size_t)ci[k] * 4 + q]);
    if (relu) acc = relu4(acc);
    out[(size_t)r * 4 + q] = acc;
}

struct Csr { std::vector<int> rp, ci; std::vector<float> va; };

static float run_lane(LaneParams lp, int threads, size_t smem, int sms, int reps, int* d_counters, const float4* Bref, float4* d_out,
                      const int* d_rp, const int* d_ci, const float* d_va) {
    auto kern = adj_window_f32_kernel;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float tot = 0.f, ms;
    for (int r = 0; r < reps + 2; r++) {
        CK(cudaMemsetAsync(d_counters, 0, 64));
        CK(cudaEventRecord(e0));
        kern<<<sms, threads, smem>>>(lp);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) tot += ms;
    }
    // deferred rows
    ref_list_kernel<<<1024, 256>>>(d_rp, d_ci, d_va, Bref, d_out, lp.long_rows, lp.long_count, lp.relu);
    CK(cudaDeviceSynchronize());
    return tot / reps;
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const int small = env_int("SMALL", 0);
    int graphs = argc > 1 ? atoi(argv[1]) : 1024, n1 = argc > 2 ? atoi(argv[2]) : 2708, M = argc > 3 ? atoi(argv[3]) : 1433;
    const int reps = argc > 4 ? atoi(argv[4]) : 20;
    std::mt19937 rng(1 + small);
    // ---- one block ----
    std::vector<std::vector<int>> nb(n1);
    for (int i = 0; i < n1; i++) nb[i].push_back(i);
    std::geometric_distribution<int> gd(0.34);
    for (int i = 0; i < n1; i++) {
        const int d = std::min(gd(rng), 30);
        for (int j = 0; j < d; j++) { const int t = rng() % n1; if (t != i) { nb[i].push_back(t); nb[t].push_back(i); } }
    }
    const int hub = std::min(168, n1 - 1);
    for (int j = 0; j < hub; j++) { const int t = 1 + rng() % (n1 - 1); nb[0].push_back(t); nb[t].push_back(0); }
    if (small) { nb[n1 / 2].clear(); nb[n1 - 1].clear(); }       // empty rows (pattern no longer symmetric: fine)
    std::vector<int> rp1(n1 + 1, 0), ci1;
    for (int i = 0; i < n1; i++) {
        std::sort(nb[i].begin(), nb[i].end());
        nb[i].erase(std::unique(nb[i].begin(), nb[i].end()), nb[i].end());
        rp1[i + 1] = rp1[i] + (int)nb[i].size();
        ci1.insert(ci1.end(), nb[i].begin(), nb[i].end());
    }
    // features: 1..30 non-zeros per row (mean ~18), sorted distinct columns
    std::vector<int> frp1(n1 + 1, 0), fci1;
    std::binomial_distribution<int> bd(36, 0.5);
    for (int i = 0; i < n1; i++) {
        int d = std::min(std::max(bd(rng), 1), std::min(30, M));
        if (small && i % 97 == 0) d = 0;
        if (small && i % 211 == 5) d = std::min(M, 700);            // rows longer than a stage -> slow path + list
        std::vector<int> cs;
        while ((int)cs.size() < d) { cs.push_back(rng() % M); std::sort(cs.begin(), cs.end()); cs.erase(std::unique(cs.begin(), cs.end()), cs.end()); }
        frp1[i + 1] = frp1[i] + d;
        fci1.insert(fci1.end(), cs.begin(), cs.end());
    }
    const long long N = (long long)graphs * n1, nnzA1 = ci1.size(), nnzF1 = fci1.size();
    const long long nnzA = nnzA1 * graphs, nnzF = nnzF1 * graphs;
    if (nnzF >= (1ll << 31)) { printf("too large\n"); return 1; }
    Csr A, F;
    A.rp.resize(N + 1); A.ci.resize(nnzA); A.va.resize(nnzA);
    F.rp.resize(N + 1); F.ci.resize(nnzF); F.va.resize(nnzF);
    for (int gi = 0; gi < graphs; gi++) {
        for (int i = 0; i < n1; i++) { A.rp[(size_t)gi * n1 + i] = (int)(gi * nnzA1 + rp1[i]); F.rp[(size_t)gi * n1 + i] = (int)(gi * nnzF1 + frp1[i]); }
        for (long long k = 0; k < nnzA1; k++) { A.ci[gi * nnzA1 + k] = ci1[k] + gi * n1; A.va[gi * nnzA1 + k] = 1.0f / (1 + (rng() % 7)); }
        for (long long k = 0; k < nnzF1; k++) { F.ci[gi * nnzF1 + k] = fci1[k]; F.va[gi * nnzF1 + k] = 0.25f * (1 + (rng() % 8)); }
    }
    A.rp[N] = (int)nnzA; F.rp[N] = (int)nnzF;
    std::vector<float> W((size_t)M * 16), XW((size_t)N * 16);
    for (auto& v : W) v = (float)((int)(rng() % 2001) - 1000) * 2.5e-4f;
    for (auto& v : XW) v = (float)((int)(rng() % 2001) - 1000) * 1e-3f;
    printf("graphs %d x %d nodes: N %lld, nnz_adj %lld (%.2f/row), nnz_fea %lld (%.2f/row), M %d\n", graphs, n1, N, nnzA, (double)nnzA / N, nnzF,
           (double)nnzF / N, M);

    int *d_arp, *d_aci, *d_frp, *d_fci, *d_counters, *d_list; float *d_ava, *d_fva, *d_W, *d_Wdup, *d_XW, *d_ref, *d_out;
    CK(cudaMalloc(&d_arp, (N + 1) * 4)); CK(cudaMalloc(&d_aci, nnzA * 4 + 16)); CK(cudaMalloc(&d_ava, nnzA * 4 + 16));
    CK(cudaMalloc(&d_frp, (N + 1) * 4)); CK(cudaMalloc(&d_fci, nnzF * 4 + 16)); CK(cudaMalloc(&d_fva, nnzF * 4 + 16));
    CK(cudaMalloc(&d_W, W.size() * 4)); CK(cudaMalloc(&d_Wdup, W.size() * 8)); CK(cudaMalloc(&d_XW, XW.size() * 4));
    CK(cudaMalloc(&d_ref, XW.size() * 4)); CK(cudaMalloc(&d_out, XW.size() * 4)); CK(cudaMalloc(&d_counters, 64)); CK(cudaMalloc(&d_list, (N + 1) * 4));
    CK(cudaMemcpy(d_arp, A.rp.data(), (N + 1) * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_aci, A.ci.data(), nnzA * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ava, A.va.data(), nnzA * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_frp, F.rp.data(), (N + 1) * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_fci, F.ci.data(), nnzF * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_fva, F.va.data(), nnzF * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_W, W.data(), W.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_XW, XW.data(), XW.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaDeviceSynchronize());

    int sms = 0, optin = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0));
    std::vector<float> href(XW.size()), hout(XW.size());
    const int rgrid = (int)((N * 4 + 255) / 256);
    int rc = 0;

    // =============================== ADJ ===============================
    if (!env_int("SKIP_ADJ", 0)) {
        ref_kernel<<<rgrid, 256>>>(d_arp, d_aci, d_ava, (const float4*)d_XW, (float4*)d_ref, (int)N, 1);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(href.data(), d_ref, href.size() * 4, cudaMemcpyDeviceToHost));
        const int ncw = env_int("NCW_ADJ", 8);
        const int CH = ncw * 32;
        int C = env_int("C_ADJ", 0), S = env_int("S_ADJ", 0);
        const double avg = (double)nnzA / N;
        if (!C) C = ((int)(CH * avg * 1.25 + 256) + 3) & ~3;
        int cap = env_int("CAP", 0);
        if (!cap) { cap = 2720; }
        const int quantum = env_int("QUANTUM", cap / 2);
        const int b_bytes = cap * 64;
        if (!S) { S = 2; while (S < 12 && lane_smem_bytes(S + 1, CH, C, b_bytes) <= (size_t)optin) S++; }
        const size_t smem = lane_smem_bytes(S, CH, C, b_bytes);
        // plan
        void* scratch; const size_t sbytes = lane_plan_scratch_bytes((int)N, quantum);
        CK(cudaMalloc(&scratch, sbytes));
        LanePlan plan;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float pms = 0;
        for (int r = 0; r < 3; r++) {
            CK(cudaEventRecord(e0));
            CK(lane_plan_build(scratch, d_arp, d_aci, (int)N, quantum, cap, 0, &plan));
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&pms, e0, e1));
        }
        int st[4]; CK(cudaMemcpy(st, plan.stats, 16, cudaMemcpyDeviceToHost));
        printf("plan %.4f ms: %d panels, %d of %lld rows inside a window (scratch %.1f MB)\n", pms, st[2], st[0], N, sbytes / 1e6);
        LaneParams lp; memset(&lp, 0, sizeof(lp));
        lp.rowptr = d_arp; lp.col = d_aci; lp.val = d_ava; lp.Bm = (const float4*)d_XW; lp.out = (float4*)d_out;
        lp.nrows = (int)N; lp.relu = 1; lp.streaming_store = 1; lp.long_thresh = env_int("LONG", 48);
        lp.stage_nnz = C; lp.stages = S; lp.win_bytes = b_bytes; lp.dbg = env_int("DBG", 0);
        lp.panels = plan.panels; lp.npanels = plan.npanels;
        lp.long_rows = d_list; lp.long_count = d_counters; lp.span_counter = d_counters + 2;
        printf("ADJ  ncw %d CH %d C %d S %d smem %zu B (window %d rows) long %d\n", ncw, CH, C, S, smem, cap, lp.long_thresh);
        if (smem > (size_t)optin) { printf("does not fit\n"); return 1; }
        CK(cudaMemset(d_out, 0xff, XW.size() * 4));
        const int threads = 32 * (ncw + 1);
        const float ms = run_lane(lp, threads, smem, sms, reps, d_counters, (const float4*)d_XW, (float4*)d_out, d_arp, d_aci, d_ava);
        const double bytes = (N + 1) * 4.0 + nnzA * 8.0 + 2.0 * N * 64.0;
        int hc[16]; CK(cudaMemcpy(hc, d_counters, 64, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hout.data(), d_out, hout.size() * 4, cudaMemcpyDeviceToHost));
        const bool ok = memcmp(href.data(), hout.data(), hout.size() * 4) == 0;
        printf("ADJ  %.4f ms  %.0f GB/s algorithmic  %.2f GTEPS  deferred rows %d  %s\n", ms, bytes / ms / 1e6, nnzA / ms / 1e6, hc[0], ok ? "bit-equal" : "DIFFER");
        if (!ok) {
            rc = 1; long long bad = 0, first = -1;
            for (size_t i = 0; i < hout.size(); i++) if (memcmp(&href[i], &hout[i], 4)) { if (first < 0) first = i; bad++; }
            printf("     %lld values differ, first at row %lld col %lld: got %g want %g\n", bad, first / 16, first % 16, hout[first], href[first]);
        }
    }
    return rc;
}
