/*
 * sgrace_oracle.c -- CPU restatement of the SGRACE fused graph layer.
 *
 * TEST INFRASTRUCTURE ONLY (see sgrace_oracle.h).  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC
 * (-ffp-contract=off matters: the reference multiplies, rounds, then adds; a
 *  fused multiply-add would skip one rounding.)
 *
 * Citations: K: = gnn-rfsoc-mt-all-2022/src/kernelMatrixmult_all.cpp,
 *            H: = .../src/matrix_mult.h,  S: = demo/sgrace_lib/sgrace.py.
 */
#include "sgrace_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* binary16 in software.  The reference's `half` (H:131-136) is Xilinx        */
/* hls_half.h (Vitis HLS 2022.1, third-party, not vendored): operands are     */
/* widened to float, the operation is done in float and the result is rounded */
/* to nearest-even binary16.  float has 24 >= 2*11+2 significand bits, so the */
/* double rounding is innocuous for + and * and this equals a correctly       */
/* rounded binary16 operation.                                                */
/* ------------------------------------------------------------------------- */
float sgo_f16_to_f32(uint16_t h)
{
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp  = (h >> 10) & 0x1fu;
    uint32_t man  = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {                      /* subnormal: normalise */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

uint16_t sgo_f32_to_f16(float f)
{
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) {                         /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0));
    }
    if (ax >= 0x477ff000u) {                         /* >= 65520 -> inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x38800000u) {                          /* < 2^-14: subnormal or zero */
        if (ax < 0x33000000u) {                      /* < 2^-25 -> 0 (2^-25 ties to even -> 0) */
            return (uint16_t)sign;
        }
        int e = (int)(ax >> 23);                     /* biased float exponent */
        uint32_t man = (ax & 0x7fffffu) | 0x800000u; /* 24-bit significand */
        int shift = 126 - e;                         /* result = man >> shift, units 2^-24 */
        uint32_t q = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1u);
        uint32_t half = 1u << (shift - 1);
        if (rem > half || (rem == half && (q & 1u))) q++;
        return (uint16_t)(sign | q);
    }
    uint32_t e = (ax >> 23) - 127 + 15;
    uint32_t man = ax & 0x7fffffu;
    uint32_t q = (e << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) q++;   /* carries into exponent correctly */
    return (uint16_t)(sign | q);
}

/* ------------------------------------------------------------------------- */
/* Per-dtype scalar arithmetic                                                */
/* ------------------------------------------------------------------------- */
/* FLOAT (H:141-151): float multiply, float add, each rounded (no FMA).       */
static inline float f32_mul(float a, float b) { return a * b; }
static inline float f32_add(float a, float b) { return a + b; }
static inline int   f32_gt0(float a) { return a > 0.0f; }

/* HALF (H:129-139).  Both the FPGA and the recorded C-simulation evaluate half
 * + and * with the Xilinx Floating-Point Operator v7.0 (csim.exe links its
 * bit-accurate model libIp_floating_point_v7_0_bitacc_cmodel.so), which has no
 * denormal support: subnormal operands are read as signed zero and subnormal
 * results are flushed to signed zero.  With this the oracle reproduces all 42
 * values of the recorded csim log and all 16 of the on-board notebook output
 * bit for bit (tests/test_oracle_golden.py); without it 2 of 58 are 1 ulp off. */
static inline uint16_t f16_ftz(uint16_t h)
{ return ((h & 0x7c00u) == 0) ? (uint16_t)(h & 0x8000u) : h; }
static inline uint16_t f16_mul(uint16_t a, uint16_t b)
{ return f16_ftz(sgo_f32_to_f16(sgo_f16_to_f32(f16_ftz(a)) * sgo_f16_to_f32(f16_ftz(b)))); }
static inline uint16_t f16_add(uint16_t a, uint16_t b)
{ return f16_ftz(sgo_f32_to_f16(sgo_f16_to_f32(f16_ftz(a)) + sgo_f16_to_f32(f16_ftz(b)))); }
static inline int f16_gt0(uint16_t a) { return sgo_f16_to_f32(a) > 0.0f; }

/* EIGHTBIT (H:105-109): ap_fixed<16,2>, Q2.14 two's complement, default modes
 * AP_TRN (drop low bits = round toward -inf) and AP_WRAP.  The 32-bit exact
 * product ap_fixed<32,4> is assigned to ITYPE (K:637-640): >>14 then wrap.   */
static inline int16_t fix_mul(int16_t a, int16_t b)
{ return (int16_t)(uint16_t)(((int32_t)a * (int32_t)b) >> 14); }
static inline int16_t fix_add(int16_t a, int16_t b)
{ return (int16_t)(uint16_t)((int32_t)a + (int32_t)b); }
static inline int fix_gt0(int16_t a) { return a > 0; }

/*
 * One compute stage (FEA: compute1_* K:2605-2929 + dsp_kernel_wrapper_fea
 * K:1960-2152; ADJ: compute2_* K:2262-2603 + dsp_kernel_wrapper_adj_* K:1413-1957).
 *
 *  - rows are split over `threads` hardware threads, thread t owning
 *    [t*floor(N/T), ...) and the last one the remainder (K:3159-3164,
 *    K:3585-3594); every thread restarts its sblock phase at its first row;
 *  - rows are grouped `sb` at a time (readptr_* K:815-895 emit the cumulative
 *    nnz of the group); the non-zeros of a group form one stream;
 *  - FLOAT/HALF: the product at stream position p is added into partial
 *    accumulator lane p % lat of the row it belongs to (K:2009-2042); padding
 *    slots past the end of the stream are never accumulated (K:2040); lanes
 *    are then folded 1..lat-1 into lane 0 in order (K:2050-2061);
 *  - FIX16: a single accumulator per row, products added in stream order
 *    (K:2104-2145) -- the same code with lat = 1;
 *  - dense gemm_mode (FEA only): every row has M stream entries, column
 *    indices are synthesised 0..M-1 (K:847-865, K:985-1013), zeros included.
 * The B operand is read as Bm[ci*brs + j*bcs].
 * Column tiles (B_index loop, K:2995/K:3409) are independent per column, so all
 * P columns are produced in one pass; only j < P_w is ever stored (K:799).
 */
#define DEFINE_STAGE(NAME, T, MUL, ADD, GT0)                                              \
static int NAME(int N, int threads, int sb, int lat, int dense, int Mdense,               \
                const int32_t *rowPtr, const int32_t *colIdx, const T *vals,              \
                const T *Bm, long brs, long bcs, int P, int relu, T *out)                 \
{                                                                                         \
    if (threads < 1 || sb < 1 || lat < 1) return -1;                                      \
    T *part = (T *)malloc(sizeof(T) * (size_t)lat * (size_t)(P > 0 ? P : 1));             \
    if (!part) return -2;                                                                 \
    for (int t = 0; t < threads; t++) {                                                   \
        int blk = N / threads;                                                            \
        int first_row = t * blk;                                                          \
        int row_count = (t == threads - 1) ? blk + N % threads : blk;                     \
        for (int a = 0; a < row_count; a += sb) {                                         \
            long base = dense ? (long)(first_row + a) * Mdense                            \
                              : (long)rowPtr[first_row + a];                              \
            for (int z = 0; z < sb && a + z < row_count; z++) {                           \
                int r = first_row + a + z;                                                \
                long beg = dense ? (long)r * Mdense : (long)rowPtr[r];                    \
                long end = dense ? beg + Mdense : (long)rowPtr[r + 1];                    \
                for (int i = 0; i < lat * P; i++) part[i] = (T)0;                         \
                for (long k = beg; k < end; k++) {                                        \
                    long pos = k - base;            /* position in the sblock stream */   \
                    int lane = (int)(pos % lat);                                          \
                    T v = vals[k];                                                        \
                    long ci = dense ? (k - beg) : (long)colIdx[k];                        \
                    T *acc = part + (size_t)lane * P;                                     \
                    const T *brow = Bm + ci * brs;                                        \
                    for (int j = 0; j < P; j++)                                           \
                        acc[j] = ADD(acc[j], MUL(v, brow[(long)j * bcs]));                \
                }                                                                         \
                for (int j = 0; j < P; j++) {                                             \
                    T acc = part[j];                                                      \
                    for (int l = 1; l < lat; l++) acc = ADD(acc, part[(size_t)l * P + j]);\
                    if (relu && !GT0(acc)) acc = (T)0;   /* K:2586-2590, K:801-804 */     \
                    out[(long)r * P + j] = acc;                                           \
                }                                                                         \
            }                                                                             \
        }                                                                                 \
    }                                                                                     \
    free(part);                                                                           \
    return 0;                                                                             \
}

DEFINE_STAGE(stage_f32, float,    f32_mul, f32_add, f32_gt0)
DEFINE_STAGE(stage_f16, uint16_t, f16_mul, f16_add, f16_gt0)
DEFINE_STAGE(stage_fix, int16_t,  fix_mul, fix_add, fix_gt0)

static int check_csr(const int32_t *rp, const int32_t *ci, int n, int ncols)
{
    if (!rp || !ci) return -1;
    for (int i = 0; i < n; i++) if (rp[i + 1] < rp[i]) return -1;
    for (long k = rp[0]; k < rp[n]; k++) if (ci[k] < 0 || ci[k] >= ncols) return -1;
    return 0;
}

int sgrace_oracle_layer(const sgo_layer_t *d)
{
    if (!d || d->N_adj < 0 || d->M_fea < 0 || d->P_w < 0) return -1;
    const int N = d->N_adj, M = d->M_fea, P = d->P_w;
    const int sb = d->spmm_block > 0 ? d->spmm_block : 1;
    const int ft = d->fea_threads > 0 ? d->fea_threads : 1;
    const int at = d->adj_threads > 0 ? d->adj_threads : 1;
    int lf = d->lat_fea > 0 ? d->lat_fea : 1;
    int la = d->lat_adj > 0 ? d->lat_adj : 1;
    if (d->dtype == SGO_FIX16) lf = la = 1;               /* H:117-118 */
    /* USE_SBLOCKS==1 writes acc2 unmodified and writec's ReLU is commented out (K:748-786) */
    const int relu = d->use_sblocks ? 0 : (d->relu != 0);
    if (N == 0 || P == 0) return 0;
    if (!d->gemm_mode && check_csr(d->rowPtr_fea, d->columnIndex_fea, N, M)) return -3;
    if (check_csr(d->rowPtr_adj, d->columnIndex_adj, N, N)) return -3;

    size_t esz = d->dtype == SGO_F32 ? 4 : 2;
    void *xw = d->XW ? d->XW : malloc(esz * (size_t)N * (size_t)P);
    if (!xw) return -2;
    int rc;
    switch (d->dtype) {
    case SGO_F32:
        rc = stage_f32(N, ft, sb, lf, d->gemm_mode, M, d->rowPtr_fea, d->columnIndex_fea,
                       (const float *)d->values_fea, (const float *)d->B, 1, M, P, 0, (float *)xw);
        if (!rc) rc = stage_f32(N, at, sb, la, 0, 0, d->rowPtr_adj, d->columnIndex_adj,
                       (const float *)d->values_adj, (const float *)xw, P, 1, P, relu, (float *)d->D);
        break;
    case SGO_F16:
        rc = stage_f16(N, ft, sb, lf, d->gemm_mode, M, d->rowPtr_fea, d->columnIndex_fea,
                       (const uint16_t *)d->values_fea, (const uint16_t *)d->B, 1, M, P, 0, (uint16_t *)xw);
        if (!rc) rc = stage_f16(N, at, sb, la, 0, 0, d->rowPtr_adj, d->columnIndex_adj,
                       (const uint16_t *)d->values_adj, (const uint16_t *)xw, P, 1, P, relu, (uint16_t *)d->D);
        break;
    case SGO_FIX16:
        rc = stage_fix(N, ft, sb, lf, d->gemm_mode, M, d->rowPtr_fea, d->columnIndex_fea,
                       (const int16_t *)d->values_fea, (const int16_t *)d->B, 1, M, P, 0, (int16_t *)xw);
        if (!rc) rc = stage_fix(N, at, sb, la, 0, 0, d->rowPtr_adj, d->columnIndex_adj,
                       (const int16_t *)d->values_adj, (const int16_t *)xw, P, 1, P, relu, (int16_t *)d->D);
        break;
    default:
        rc = -1;
    }
    if (!d->XW) free(xw);
    return rc;
}

int sgrace_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int sgrace_oracle_layer_batch(const sgo_layer_t *d, int count, int threads)
{
    int bad = 0;
    if (threads < 1) threads = 1;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) reduction(| : bad)
#endif
    for (int i = 0; i < count; i++) bad |= (sgrace_oracle_layer(&d[i]) != 0);
    return bad ? -1 : 0;
}

/* ------------------------------------------------------------------------- */
/* Full design, emulation semantics (S:563-681)                               */
/* ------------------------------------------------------------------------- */

/* fake_quantization / fake_quantization_b / fake_quantization_b2 (S:177-235):
 * returns the integer code; the value is code / den.                         */
static inline int q_code_unsigned(float x, float inv_s, int z, int qbits)
{   /* quantization_ufbits S:253-265: round-half-even, clip to [0, 2^q - 1] */
    float r = nearbyintf(inv_s * x + (float)z);
    float hi = (float)((1 << qbits) - 1);
    if (r < 0.0f) r = 0.0f;
    if (r > hi) r = hi;
    return (int)r;
}
static inline int q_code_signed(float x, float inv_s, int z, int qbits)
{   /* quantization_fbits S:238-251 */
    if (qbits == 1) {           /* fake_quantization_b: sign -> +-0.5 */
        float t = inv_s * x + (float)z;
        return (t < 0.0f) ? -1 : 1;
    }
    float r = nearbyintf(inv_s * x + (float)z);
    float hi = (float)((1 << (qbits - 1)) - 1);
    if (r < -hi) r = -hi;
    if (r > hi) r = hi;
    return (int)r;
}
static inline float q_den(int qbits)
{   /* x_q / 2^(w_qbits-1) (S:215); the 1-bit variants divide by 2 (S:179-188) */
    return qbits == 1 ? 2.0f : (float)(1 << (qbits - 1));
}

/* torch.round(x, decimals=d) on a float32 CPU tensor: nearbyint(x * T) / T with
 * T = (float)pow(10, d)  (S:616 uses it with d = internal_quantization - 1). */
static inline float round_decimals_f32(float x, float T)
{
    float y = x * T;
    y = nearbyintf(y);
    return y / T;
}

int sgrace_oracle_qlayer(const sgo_qlayer_t *d)
{
    if (!d) return -1;
    const int N = d->N_adj, M = d->M_fea, P = d->P_w, q = d->qbits;
    if (N <= 0 || P <= 0) return 0;
    if (!d->gemm_mode && check_csr(d->rowPtr_fea, d->columnIndex_fea, N, M)) return -3;
    if (check_csr(d->rowPtr_adj, d->columnIndex_adj, N, N)) return -3;
    const long nnz_adj = d->rowPtr_adj[N];

    float *Wh = d->Wh ? d->Wh : (float *)malloc(sizeof(float) * (size_t)N * P);
    if (!Wh) return -2;
    const float inv_fs = d->qscale_fea, inv_ws = d->qscale_w, inv_as = d->qscale_adj;
    int64_t maxabs = 0;

    if (q) {
        /* W codes, P x M transposed like the B buffer */
        int32_t *wq = (int32_t *)malloc(sizeof(int32_t) * (size_t)M * P);
        int64_t *acc = (int64_t *)malloc(sizeof(int64_t) * (size_t)P);
        if (!wq || !acc) { free(wq); free(acc); if (!d->Wh) free(Wh); return -2; }
        for (long i = 0; i < (long)M * P; i++) wq[i] = q_code_signed(d->B[i], inv_ws, d->w_z, q);
        const float den = q_den(q) * q_den(q);
        const float a_hi = (float)(((double)(1u << d->internal_quantization) - 1.0) /
                                   (double)(1u << d->internal_quantization));     /* S:607-608 */
        const float T = (float)pow(10.0, (double)(d->internal_quantization - 1)); /* S:616 */
        const float sc = (float)(1 << d->scale_fea);
        for (int r = 0; r < N; r++) {
            for (int j = 0; j < P; j++) acc[j] = 0;
            if (d->gemm_mode) {
                for (int c = 0; c < M; c++) {
                    int xi = q_code_unsigned(d->values_fea[(long)r * M + c], inv_fs, d->f_z, q);
                    if (!xi) continue;
                    for (int j = 0; j < P; j++) acc[j] += (int64_t)xi * wq[(long)j * M + c];
                }
            } else {
                for (long k = d->rowPtr_fea[r]; k < d->rowPtr_fea[r + 1]; k++) {
                    int xi = q_code_unsigned(d->values_fea[k], inv_fs, d->f_z, q);
                    int c = d->columnIndex_fea[k];
                    if (!xi) continue;
                    for (int j = 0; j < P; j++) acc[j] += (int64_t)xi * wq[(long)j * M + c];
                }
            }
            for (int j = 0; j < P; j++) {
                int64_t a = acc[j] < 0 ? -acc[j] : acc[j];
                if (a > maxabs) maxabs = a;
                /* torch.mm of exact dyadic rationals: exact while |sum| < 2^24 */
                float wh = (float)acc[j] / den;                 /* S:601 */
                wh = wh / sc;                                    /* S:606 */
                if (wh < -a_hi) wh = -a_hi;                      /* S:615 */
                if (wh > a_hi) wh = a_hi;
                Wh[(long)r * P + j] = round_decimals_f32(wh, T); /* S:616 */
            }
        }
        if (d->max_fea) {
            double m = (double)maxabs / (double)den * 65536.0;   /* frac_bits_o = 16, S:1288 */
            *d->max_fea = m > 2147483647.0 ? 2147483647 : (int32_t)m;
        }
        free(wq); free(acc);
    } else {
        /* fake_quantization == 0: plain float32 X W, sequential multiply-add */
        for (int r = 0; r < N; r++) {
            float *o = Wh + (long)r * P;
            for (int j = 0; j < P; j++) o[j] = 0.f;
            if (d->gemm_mode) {
                for (int c = 0; c < M; c++) {
                    float v = d->values_fea[(long)r * M + c];
                    for (int j = 0; j < P; j++) o[j] = o[j] + v * d->B[(long)j * M + c];
                }
            } else {
                for (long k = d->rowPtr_fea[r]; k < d->rowPtr_fea[r + 1]; k++) {
                    float v = d->values_fea[k];
                    int c = d->columnIndex_fea[k];
                    for (int j = 0; j < P; j++) o[j] = o[j] + v * d->B[(long)j * M + c];
                }
            }
        }
        if (d->max_fea) *d->max_fea = 0;
    }

    /* adjacency: quantization_ufbits, zero codes vanish in .to_sparse() (S:626-629) */
    float *aq = (float *)malloc(sizeof(float) * (size_t)(nnz_adj > 0 ? nnz_adj : 1));
    if (!aq) { if (!d->Wh) free(Wh); return -2; }
    for (long k = 0; k < nnz_adj; k++)
        aq[k] = q ? (float)q_code_unsigned(d->values_adj[k], inv_as, d->a_z, q) / q_den(q)
                  : d->values_adj[k];

    float *s1 = NULL, *s2 = NULL, *colmean = NULL, *ebuf = NULL;
    if (d->gat_mode) {
        ebuf = d->E ? d->E : (float *)malloc(sizeof(float) * (size_t)(nnz_adj > 0 ? nnz_adj : 1));
        /* prepare_attentional_mechanism_input S:309-314; attention is quantised
         * with the weight quantiser first (S:624) */
        s1 = (float *)malloc(sizeof(float) * (size_t)N);
        s2 = (float *)malloc(sizeof(float) * (size_t)N);
        float *att = (float *)malloc(sizeof(float) * (size_t)(2 * P));
        for (int j = 0; j < 2 * P; j++)
            att[j] = q ? (float)q_code_signed(d->attention[j], inv_ws, d->w_z, q) / q_den(q)
                       : d->attention[j];
        for (int r = 0; r < N; r++) {
            float a = 0.f, b = 0.f;
            for (int j = 0; j < P; j++) {
                a = a + Wh[(long)r * P + j] * att[j];
                b = b + Wh[(long)r * P + j] * att[P + j];
            }
            s1[r] = a; s2[r] = b;
        }
        free(att);
    }

    const float dq = d->deq_o;
    for (int r = 0; r < N; r++) {
        float *o = d->D + (long)r * P;
        for (int j = 0; j < P; j++) o[j] = 0.f;
        long beg = d->rowPtr_adj[r], end = d->rowPtr_adj[r + 1];
        if (!d->gat_mode) {
            for (long k = beg; k < end; k++) {
                if (aq[k] == 0.f) continue;                      /* pruned */
                const float *w = Wh + (long)d->columnIndex_adj[k] * P;
                for (int j = 0; j < P; j++) o[j] = o[j] + aq[k] * w[j];   /* S:656 */
            }
        } else {
            /* e = LeakyReLU(Wh_i a1 + Wh_j a2) on surviving edges, row softmax,
             * out = att Wh (S:634-650) */
            float mx = -INFINITY;
            int live = 0;
            for (long k = beg; k < end; k++) {
                float e = 0.f;
                if (aq[k] > 0.f) {
                    e = s1[r] + s2[d->columnIndex_adj[k]];
                    e = e > 0.f ? e : d->alpha * e;
                    if (e > mx) mx = e;
                    live++;
                }
                ebuf[k] = e;
            }
            if (live) {
                float sum = 0.f;
                for (long k = beg; k < end; k++)
                    if (aq[k] > 0.f) sum = sum + expf(ebuf[k] - mx);
                for (long k = beg; k < end; k++) {
                    float s = 0.f;
                    if (aq[k] > 0.f) {
                        s = expf(ebuf[k] - mx) / sum;
                        const float *w = Wh + (long)d->columnIndex_adj[k] * P;
                        for (int j = 0; j < P; j++) o[j] = o[j] + s * w[j];
                    }
                    if (d->S) d->S[k] = s;
                }
            } else {
                /* no surviving edge: every masked logit is -9e15, so the dense
                 * softmax of the emulation is uniform 1/N over ALL nodes (S:638-641) */
                if (!colmean) {
                    colmean = (float *)calloc((size_t)P, sizeof(float));
                    for (int j = 0; j < P; j++) {
                        double acc = 0.0;
                        for (int i = 0; i < N; i++) acc += (double)Wh[(long)i * P + j];
                        colmean[j] = (float)(acc / (double)N);
                    }
                }
                for (int j = 0; j < P; j++) o[j] = colmean[j];
                for (long k = beg; k < end; k++) if (d->S) d->S[k] = 0.f;
            }
        }
        for (int j = 0; j < P; j++) {
            float v = o[j];
            if (d->relu && !(v > 0.f)) v = 0.f;                  /* S:660-661 */
            if (q) v = v * dq;                                   /* S:666-667 */
            o[j] = v;
        }
    }
    free(aq); free(s1); free(s2); free(colmean);
    if (ebuf && ebuf != d->E) free(ebuf);
    if (!d->Wh) free(Wh);
    return 0;
}
