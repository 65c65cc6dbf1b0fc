// sgrace_gemm_tc.cuh -- tensor-core path for the one real contraction on the hot path:
// dense FEA (gemm_mode = 1) with a wide hidden layer (e.g. ogbn-products shape 100 -> 256).
// Placeholder until the tcgen05 kernel lands: reports "unsupported" so the CUDA-core kernel runs.
#pragma once
#include <cuda_runtime.h>
namespace sgrace {
inline bool fea_dense_tc_supported(int, int, int) { return false; }
inline int fea_dense_tc_launch(const float*, const float*, float*, int, int, int, int, cudaStream_t) { return -100; }
}  // namespace sgrace
