/*
 * Stand-in for Xilinx Vitis-HLS 2022.1 "ap_int.h" (third-party, NOT vendored by the
 * reference) -- just enough for g++ to compile the reference's own kernel source
 * (gnn-rfsoc-mt-all-2022/src/kernelMatrixmult_all.cpp) unmodified, from where it lies,
 * into oracle/_ref/.  TEST INFRASTRUCTURE ONLY; written for this repo, no Xilinx code.
 *
 *  ap_int<N>/ap_uint<N>: the kernel only uses them as plain scalars (loop bounds, array
 *    indices, register values), so they are thin wrappers over long long.
 *  half: the reference's `half` (via hls_half.h) evaluated by the Xilinx Floating-Point
 *    Operator v7.0 model: round-to-nearest-even binary16, denormals flushed to signed
 *    zero on operands and results (see oracle/sgrace_oracle.c for the evidence).
 *    -DSGRACE_REF_ELT_FLOAT turns `half` into plain float, which yields the reference's
 *    dataflow with float arithmetic and FADD latency 4.
 */
#ifndef SGRACE_SHIM_AP_INT_H
#define SGRACE_SHIM_AP_INT_H
#include <stdint.h>
#include <string.h>

template <int N> struct ap_int {
    long long v;
    ap_int() {}
    ap_int(long long x) : v(x) {}
    operator long long() const { return v; }
    ap_int operator++(int) { ap_int t = *this; ++v; return t; }
    ap_int &operator++() { ++v; return *this; }
};
template <int N> struct ap_uint {
    unsigned long long v;
    ap_uint() {}
    ap_uint(unsigned long long x) : v(x) {}
    operator unsigned long long() const { return v; }
    ap_uint operator++(int) { ap_uint t = *this; ++v; return t; }
    ap_uint &operator++() { ++v; return *this; }
};

/*  ap_fixed<16, 2> (the EIGHTBIT build's ATYPE / BTYPE / DTYPE / FTYPE / ITYPE, matrix_mult.h:105-109) with the
 *  documented Vitis-HLS defaults: quantisation AP_TRN (truncate towards minus infinity) and overflow AP_WRAP.  Raw
 *  two's-complement storage (Q2.14), so a buffer of int16 codes is a buffer of this type.  Only what the kernel source
 *  uses: construction from 0 / numbers, product and sum assigned back to the same type, comparison with 0.
 *  A full-precision product added to an accumulator and then truncated equals the truncated product added
 *  (acc * 2^14 is a multiple of 2^14), so `acc += a * b` and `t = a * b; acc += t` agree, as in the real type. */
template <int W, int I> struct ap_fixed {
    static_assert(W == 16, "shim: 16-bit storage only");
    int16_t v;
    static int16_t wrap(long long r) { return (int16_t)(uint16_t)(unsigned long long)r; }
    static long long floor_to_raw(double x) { double y = x * (double)(1 << (W - I)); long long r = (long long)y; if ((double)r > y) r--; return r; }
    ap_fixed() {}
    ap_fixed(int x) : v(wrap((long long)x * (1 << (W - I)))) {}
    ap_fixed(long long x) : v(wrap(x * (1 << (W - I)))) {}
    ap_fixed(double x) : v(wrap(floor_to_raw(x))) {}
    ap_fixed(float x) : v(wrap(floor_to_raw((double)x))) {}
    static ap_fixed raw(int16_t r) { ap_fixed f; f.v = r; return f; }
    explicit operator double() const { return (double)v / (double)(1 << (W - I)); }
    ap_fixed &operator+=(const ap_fixed &o) { v = wrap((long long)v + (long long)o.v); return *this; }
};
template <int W, int I> inline ap_fixed<W, I> operator*(const ap_fixed<W, I> &a, const ap_fixed<W, I> &b) {
    return ap_fixed<W, I>::raw(ap_fixed<W, I>::wrap(((long long)a.v * (long long)b.v) >> (W - I)));   /* >> on a negative value: floor (AP_TRN) */
}
template <int W, int I> inline ap_fixed<W, I> operator+(const ap_fixed<W, I> &a, const ap_fixed<W, I> &b) { ap_fixed<W, I> r = a; r += b; return r; }
template <int W, int I> inline bool operator>(const ap_fixed<W, I> &a, int z) { return (long long)a.v > (long long)z * (1 << (W - I)); }
template <int W, int I> inline bool operator<(const ap_fixed<W, I> &a, int z) { return (long long)a.v < (long long)z * (1 << (W - I)); }

#if defined(SGRACE_REF_ELT_FLOAT)
typedef float half;
#else
namespace sgshim {
inline float h2f(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (!man) bits = sign;
        else { int e = -1; do { man <<= 1; e++; } while (!(man & 0x400u)); man &= 0x3ffu;
               bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13); }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float f; memcpy(&f, &bits, 4); return f;
}
inline uint16_t f2h(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u, ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (ax > 0x7f800000u ? 0x200u : 0));
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);
    if (ax < 0x38800000u) {
        if (ax < 0x33000000u) return (uint16_t)sign;
        int e = (int)(ax >> 23); uint32_t man = (ax & 0x7fffffu) | 0x800000u; int sh = 126 - e;
        uint32_t q = man >> sh, rem = man & ((1u << sh) - 1u), hf = 1u << (sh - 1);
        if (rem > hf || (rem == hf && (q & 1u))) q++;
        return (uint16_t)(sign | q);
    }
    uint32_t q = ((((ax >> 23) - 112) << 10) | ((ax & 0x7fffffu) >> 13)), rem = ax & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) q++;
    return (uint16_t)(sign | q);
}
inline uint16_t ftz(uint16_t h) { return ((h & 0x7c00u) == 0) ? (uint16_t)(h & 0x8000u) : h; }
}  // namespace sgshim

struct half {
    uint16_t b;
    half() {}
    half(float f) : b(sgshim::f2h(f)) {}
    half(double f) : b(sgshim::f2h((float)f)) {}
    half(int i) : b(sgshim::f2h((float)i)) {}
    operator float() const { return sgshim::h2f(b); }
    static half raw(uint16_t x) { half h; h.b = x; return h; }
    half &operator+=(const half &o) {
        b = sgshim::ftz(sgshim::f2h(sgshim::h2f(sgshim::ftz(b)) + sgshim::h2f(sgshim::ftz(o.b))));
        return *this;
    }
};
inline half operator*(const half &a, const half &c) {
    return half::raw(sgshim::ftz(sgshim::f2h(sgshim::h2f(sgshim::ftz(a.b)) * sgshim::h2f(sgshim::ftz(c.b)))));
}
inline half operator+(const half &a, const half &c) { half r = a; r += c; return r; }
inline bool operator>(const half &a, int z) { return sgshim::h2f(a.b) > (float)z; }
inline bool operator<(const half &a, int z) { return sgshim::h2f(a.b) < (float)z; }
#endif
#endif
