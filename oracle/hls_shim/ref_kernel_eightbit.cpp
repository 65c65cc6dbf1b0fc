/*
 * The reference's kernel source in its EIGHTBIT configuration (matrix_mult.h:77: ap_fixed<16,2> types, latency 1).
 * The build selector of the reference is an edit of matrix_mult.h ("#define HALF" at line 80); the Makefile writes a copy of
 * that header with the one selector line changed into oracle/_ref/eightbit/ (a build output, git-ignored), which this
 * unit includes first: the include guard then skips the original when the kernel source includes it.
 * TEST INFRASTRUCTURE ONLY.
 */
#include <hls_stream.h>
#include "ap_int.h"
#include "matrix_mult.h"              /* found in oracle/_ref/eightbit (first -I) */
#ifndef EIGHTBIT
#error "the EIGHTBIT copy of matrix_mult.h must come first on the include path"
#endif
#include "kernelMatrixmult_all.cpp"   /* found through -I <reference>/src */
