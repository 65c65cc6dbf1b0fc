// Graph preparation on the GPU: the per-forward host work of the reference
// that dominates once a layer takes microseconds.
//
//   sym_norm      demo/sgrace_lib/sgrace.py:18-51 (sym_norm2): add the missing self-loops, sort the edges by
//                 (row, col), deg = row sums in that order, value = deg^-1/2[row] * w * deg^-1/2[col]
//   dense_to_csr  sgrace.py:1218-1227 / Graph_Classification.ipynb cell 18:53-58 (`to_sparse` of X)
//
// Same arithmetic and the same order of additions as the torch code, so the values are bit-equal to the host
// mirror (sgracex1_b200/sgrace.py), which is pinned against the reference's own output.  The sort and the
// scans are CUB device primitives (library code, like cuBLAS for a plain GEMM); everything else is here.
#pragma once

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgrace {
namespace prep {

// keys: (row << 32) | col for kept edges, all-ones for the original self-loops (they sort last and are
// dropped; their weight moves to the appended loop), then one (i, i) entry per node.
__global__ void sym_keys_kernel(const int* __restrict__ row, const int* __restrict__ col, long long nnz, int n,
                                unsigned long long* __restrict__ keys, int* __restrict__ ids,
                                int* __restrict__ loop_edge, int* __restrict__ counters /* [0]: self loops, [1]: bad index */) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) {
        const int r = row[i], c = col[i];
        ids[i] = (int)i;
        if ((unsigned)r >= (unsigned)n || (unsigned)c >= (unsigned)n) {
            atomicAdd(counters + 1, 1);
            keys[i] = ~0ull;
        } else if (r == c) {
            atomicAdd(counters, 1);
            atomicMax(loop_edge + r, (int)i);          // the last self-loop of a node wins, as torch's indexed assignment does
            keys[i] = ~0ull;
        } else {
            keys[i] = ((unsigned long long)(unsigned)r << 32) | (unsigned)c;
        }
    } else if (i < nnz + n) {
        const unsigned v = (unsigned)(i - nnz);
        keys[i] = ((unsigned long long)v << 32) | v;
        ids[i] = (int)i;
    }
}

__global__ void sym_emit_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ ids,
                                const float* __restrict__ weight, const int* __restrict__ loop_edge, float fill,
                                long long nnz, long long out_nnz, int* __restrict__ out_row, int* __restrict__ out_col,
                                float* __restrict__ out_w) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= out_nnz) return;
    const unsigned long long key = keys[k];
    const int id = ids[k];
    out_row[k] = (int)(key >> 32);
    out_col[k] = (int)(key & 0xffffffffu);
    float w;
    if (id < nnz) {
        w = weight ? weight[id] : 1.f;
    } else {
        const int e = loop_edge[id - nnz];
        w = e >= 0 ? (weight ? weight[e] : 1.f) : fill;
    }
    out_w[k] = w;
}

// deg[r] = sum of the row's weights in sorted order (what scatter_add / index_add_ do on the CPU);
// dis = deg^-1/2 with inf -> 0.  torch's pow(-0.5) is 1 / sqrt(x) with IEEE rounding of both steps.
__global__ void sym_deg_kernel(const int* __restrict__ rowptr, const float* __restrict__ w, int n, float* __restrict__ dis) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float deg = 0.f;
    for (int k = rowptr[r]; k < rowptr[r + 1]; k++) deg = __fadd_rn(deg, w[k]);
    float d = __frcp_rn(__fsqrt_rn(deg));
    if (isinf(d)) d = 0.f;
    dis[r] = d;
}

__global__ void sym_norm_kernel(const int* __restrict__ row, const int* __restrict__ col, const float* __restrict__ w,
                                const float* __restrict__ dis, long long nnz, float* __restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    out[k] = __fmul_rn(__fmul_rn(dis[row[k]], w[k]), dis[col[k]]);
}

// ---- dense X -> CSR: one warp per row, ballot compaction keeps the column order ----
__global__ void dense_count_kernel(const float* __restrict__ X, int n, int m, int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    int c = 0;
    for (int j = lane; j < m; j += 32) c += X[(size_t)row * m + j] != 0.f;
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) counts[row] = c;
}

__global__ void dense_fill_kernel(const float* __restrict__ X, int n, int m, const int* __restrict__ rowptr,
                                  long long capacity, int* __restrict__ col, float* __restrict__ val) {
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    long long pos = rowptr[row];
    for (int j0 = 0; j0 < m; j0 += 32) {
        const int j = j0 + lane;
        const float x = j < m ? X[(size_t)row * m + j] : 0.f;
        const unsigned live = __ballot_sync(0xffffffffu, x != 0.f);
        if (x != 0.f) {
            const long long k = pos + __popc(live & ((1u << lane) - 1u));
            if (k < capacity) { col[k] = j; val[k] = x; }
        }
        pos += __popc(live);
    }
}

}  // namespace prep

// ---- adaptive pruning (sgrace.py:626-629): adjacency entries whose quantised code is 0 are dropped -------------
// flag[k] = 1 iff clip(round(a_k / a_s + z), 0, 2^q - 1) != 0   (quantization_ufbits, round-half-even); flag[nnz] = 0
__global__ void prune_flag_kernel(const float* __restrict__ val, long long nnz, float inv_as, int a_z, int qbits, int* __restrict__ flag) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nnz) return;
    int f = 0;
    if (k < nnz) {
        float r = rintf(__fadd_rn(__fmul_rn(inv_as, val[k]), (float)a_z));
        const float hi = (float)((1 << qbits) - 1);
        r = r < 0.f ? 0.f : (r > hi ? hi : r);
        f = r != 0.f;
    }
    flag[k] = f;
}
// pos = exclusive prefix sum of flag (nnz + 1 entries): survivors keep their order
__global__ void prune_scatter_kernel(const int* __restrict__ col, const float* __restrict__ val, const int* __restrict__ pos, long long nnz,
                                     int* __restrict__ out_col, float* __restrict__ out_val, int* __restrict__ kept) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const int p0 = pos[k];
    if (pos[k + 1] != p0) {
        out_col[p0] = col[k];
        out_val[p0] = val[k];
        if (kept) kept[p0] = (int)k;
    }
}
__global__ void prune_rowptr_kernel(const int* __restrict__ rowptr, const int* __restrict__ pos, int n, int* __restrict__ out_rowptr) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= n) out_rowptr[r] = pos[rowptr[r]];
}

}  // namespace sgrace
