// Standalone check + timing of the tcgen05 dense-FEA kernel (csrc/sgrace_gemm_tc.cuh).
#include "../../sgracex1_b200/csrc/sgrace_gemm_tc.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 1000, M = argc > 2 ? atoi(argv[2]) : 100, P = argc > 3 ? atoi(argv[3]) : 256;
    const int BKF = argc > 4 ? atoi(argv[4]) : 0;
    printf("N=%d M=%d P=%d supported=%d\n", N, M, P, (int)sgrace::fea_dense_tc_supported(N, M, P));
    std::vector<float> X((size_t)N * M), B((size_t)P * M), out((size_t)N * P);
    srand(1);
    for (auto& v : X) v = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
    for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 0.2f;
    float *dX, *dB, *dO;
    cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, out.size() * 4);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0xff, out.size() * 4);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int rc = sgrace::fea_dense_tc_launch(dX, dB, dO, N, M, P, sms, 0, BKF);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch rc=%d sync=%s\n", rc, cudaGetErrorString(e));
    if (rc || e != cudaSuccess) return 1;
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; long bad = 0;
    const int check_rows = N < 4096 ? N : 4096;
    for (int rr = 0; rr < check_rows; rr++) {
        const int r = (int)((long long)rr * N / check_rows);
        for (int p = 0; p < P; p++) {
            double acc = 0;
            for (int k = 0; k < M; k++) acc += (double)X[(size_t)r * M + k] * (double)B[(size_t)p * M + k];
            const double err = fabs(acc - (double)out[(size_t)r * P + p]);
            if (!(err <= 1e-5 * (fabs(acc) + 1.0))) bad++;
            if (err > maxerr) maxerr = err;
            if (fabs(acc) > maxref) maxref = fabs(acc);
        }
    }
    printf("max abs err %.3e (max |ref| %.3f), rel-to-max %.3e, elements beyond 1e-5: %ld\n", maxerr, maxref, maxerr / maxref, bad);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) sgrace::fea_dense_tc_launch(dX, dB, dO, N, M, P, sms, 0, BKF);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; i++) sgrace::fea_dense_tc_launch(dX, dB, dO, N, M, P, sms, 0, BKF);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double bytes = (double)N * M * 4 + (double)M * P * 4 + (double)N * P * 4;
    printf("%.3f ms  %.0f GB/s  %.1f TFLOP/s (useful 2NMP)\n", ms, bytes / ms / 1e6, 2.0 * N * M * P / ms / 1e9);
    return bad ? 2 : 0;
}
