/*
 * C entry to the reference's own `mmult_top` (kernelMatrixmult_all.cpp:3762), compiled
 * unmodified from /root/reference against the shim headers in this directory.
 * TEST INFRASTRUCTURE ONLY: validates oracle/sgrace_oracle.c and can serve as the CPU
 * baseline ("kind": "reference").  Build-time configuration is whatever the reference's
 * matrix_mult.h fixes: HALF types, FADD latency 4, SPMM_BLOCK 1, 1 FEA / 1 ADJ thread,
 * B_WIDTH_BLOCK 2, caps MAX_N = MAX_M = 6144.  The fix16 build (ref_kernel_eightbit.cpp) selects the EIGHTBIT types
 * (ap_fixed<16,2> stand-in, latency 1) instead.
 */
#include <stdio.h>
#include <stdlib.h>
#include <hls_stream.h>
#include "ap_int.h"
#include "matrix_mult.h"

void mmult_top(bool gemm_mode, bool relu, ap_int<32> *quantized_multiplier, ap_int<32> *shift,
               ap_int<32> *bias, ap_int<32> bias_count, ap_int<64> *profiling,
               ap_int<8> zero_point_lhs, ap_int<8> zero_point_rhs, ap_int<8> zero_point_dst,
               ap_int<8> clamp_max, ap_int<8> clamp_min, int N_adj, int M_adj, int M_fea, int P_w,
               BTYPE *B, DTYPE *D1, DTYPE *D2, DTYPE *D3, DTYPE *D4, int array_c_adjust,
               int *rowPtr_fea1, int *rowPtr_fea2, int *rowPtr_fea3, int *rowPtr_fea4,
               int *columnIndex_fea1, int *columnIndex_fea2, int *columnIndex_fea3, int *columnIndex_fea4,
               FTYPE *values_fea1, FTYPE *values_fea2, FTYPE *values_fea3, FTYPE *values_fea4,
               int *rowPtr_adj1, int *rowPtr_adj2, int *rowPtr_adj3, int *rowPtr_adj4,
               int *columnIndex_adj1, int *columnIndex_adj2, int *columnIndex_adj3, int *columnIndex_adj4,
               ATYPE *values_adj1, ATYPE *values_adj2, ATYPE *values_adj3, ATYPE *values_adj4);

extern "C" {

int sgrace_ref_elt_bytes(void) { return (int)sizeof(DTYPE); }
int sgrace_ref_max_n(void) { return MAX_N; }
int sgrace_ref_max_m(void) { return MAX_M; }
int sgrace_ref_spmm_block(void) { return SPMM_BLOCK; }
int sgrace_ref_lat(void) { return FTYPE_LATENCY_ADJ; }

/* values/B/D are raw storage: uint16 binary16 patterns (half build) or float (float build) */
int sgrace_ref_mmult_top(int gemm_mode, int relu, int N_adj, int M_adj, int M_fea, int P_w,
                         void *B, void *D,
                         int *rowPtr_fea, int *columnIndex_fea, void *values_fea,
                         int *rowPtr_adj, int *columnIndex_adj, void *values_adj)
{
    if (N_adj > MAX_N || M_fea > MAX_M) return -1;
    static thread_local ap_int<32> qm[1024], sh[1024], bias[1024];
    ap_int<64> prof[16];
    mmult_top(gemm_mode != 0, relu != 0, qm, sh, bias, 0, prof, 0, 0, 0, 0, 0, N_adj, M_adj, M_fea, P_w,
              (BTYPE *)B, (DTYPE *)D, (DTYPE *)D, (DTYPE *)D, (DTYPE *)D, N_adj,
              rowPtr_fea, rowPtr_fea, rowPtr_fea, rowPtr_fea,
              columnIndex_fea, columnIndex_fea, columnIndex_fea, columnIndex_fea,
              (FTYPE *)values_fea, (FTYPE *)values_fea, (FTYPE *)values_fea, (FTYPE *)values_fea,
              rowPtr_adj, rowPtr_adj, rowPtr_adj, rowPtr_adj,
              columnIndex_adj, columnIndex_adj, columnIndex_adj, columnIndex_adj,
              (ATYPE *)values_adj, (ATYPE *)values_adj, (ATYPE *)values_adj, (ATYPE *)values_adj);
    return 0;
}
}
