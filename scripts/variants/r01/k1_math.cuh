// Correctly rounded float32 log10 and 10**x for the K1 front end, evaluated in double with small
// tables (k1_tables.cuh) and a rounding test: when the double result is too close to a float32
// rounding boundary to be trusted, the caller falls back to the full-precision libdevice / libm
// routine.  The target value is RN_f32(log10(v)) resp. RN_f32(10**q) -- the platform-independent
// pin of numpy's float32 ufuncs (oracle/oracle_c.c header).
//
// Compiles for host too (tests/host harness): OFP_HD expands to __host__ __device__ under nvcc.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "k1_tables.cuh"

#ifdef __CUDACC__
#define OFP_HD __host__ __device__ __forceinline__
#else
#define OFP_HD static inline
#endif

namespace ofp {

OFP_HD double bits2d(uint64_t b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(static_cast<long long>(b));
#else
    double d; memcpy(&d, &b, 8); return d;
#endif
}
OFP_HD uint64_t d2bits(double d) {
#ifdef __CUDA_ARCH__
    return static_cast<uint64_t>(__double_as_longlong(d));
#else
    uint64_t b; memcpy(&b, &d, 8); return b;
#endif
}
OFP_HD uint32_t f2bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4); return b;
#endif
}
OFP_HD double fma_d(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// Is the double d within `win` units (of 2^-52 relative) of a float32 rounding midpoint?
OFP_HD bool near_f32_midpoint(double d, uint32_t win) {
    const uint32_t low = static_cast<uint32_t>(d2bits(d)) & 0x1fffffffu;  // the 29 bits float32 drops
    return (low - (0x10000000u - win)) < 2u * win;
}

// Polynomial / scaling constants.  The kernel keeps one copy in registers for its whole lifetime
// (a warp-per-CTA kernel has registers to spare) instead of re-materialising 64-bit immediates.
struct MathConst {
    double a1, a2, a3, a4, a5, log10_2;      // log10(1+r)
    double e1, e2, e3, e4, e5, log2_10;      // 2^r
    double shift;                            // 1.5 * 2^52
};
OFP_HD MathConst math_const() {
    MathConst c;
    c.a1 = OFP_LOG_A1; c.a2 = OFP_LOG_A2; c.a3 = OFP_LOG_A3; c.a4 = OFP_LOG_A4; c.a5 = OFP_LOG_A5;
    c.log10_2 = OFP_LOG10_2;
    c.e1 = OFP_EXP_E1; c.e2 = OFP_EXP_E2; c.e3 = OFP_EXP_E3; c.e4 = OFP_EXP_E4; c.e5 = OFP_EXP_E5;
    c.log2_10 = OFP_LOG2_10;
    c.shift = 0x1.8p52;
    return c;
}

// log10 of a positive normal float32 (bit pattern ix), double result with relative error < 2^-41.
// tab: {invc, logc} pairs (OFP_LOGTAB_H) in whatever memory the caller staged them.
OFP_HD double log10_core(uint32_t ix, const double *tab, const MathConst &mc) {
    const uint32_t tmp = ix - OFP_LOG_OFF;
    const int32_t k = static_cast<int32_t>(tmp) >> 23;
    const uint32_t i = (tmp >> (23 - OFP_LOG_N)) & ((1u << OFP_LOG_N) - 1u);
    const uint32_t iz = ix - (tmp & 0xff800000u);  // z in [OFF, 2*OFF)
    // float32 bits -> double bits (z is normal): exponent rebias 127 -> 1023
    const uint64_t zb = (static_cast<uint64_t>((iz >> 3) + 0x38000000u) << 32) | (static_cast<uint64_t>(iz) << 61 >> 32);
    const double z = bits2d(zb);
    const double invc = tab[2 * i], logc = tab[2 * i + 1];
    const double r = fma_d(z, invc, -1.0);
    const double r2 = r * r;
    // r * (A1 + A2 r + A3 r^2 + A4 r^3 + A5 r^4), Estrin
    const double p01 = fma_d(r, mc.a2, mc.a1);
    const double p23 = fma_d(r, mc.a4, mc.a3);
    const double p = fma_d(r2, fma_d(r2, mc.a5, p23), p01);
    const double base = fma_d(static_cast<double>(k), mc.log10_2, logc);
    return fma_d(r, p, base);
}

// 10**q for |q| < 30, double result with relative error < 2^-46.  tab: OFP_EXPTAB_H (2^(j/32)).
OFP_HD double exp10_core(float q, const double *tab, const MathConst &mc) {
    const double t = static_cast<double>(q) * mc.log2_10;
    const double shift = mc.shift;  // round-to-nearest-integer magic number
    const double kd0 = fma_d(t, 32.0, shift);
    const int32_t ki = static_cast<int32_t>(static_cast<uint32_t>(d2bits(kd0)));
    const double kd = kd0 - shift;
    const double r = fma_d(kd, -0.03125, t);  // |r| <= 1/64
    const double r2 = r * r;
    const double p01 = fma_d(r, mc.e2, mc.e1);
    const double p23 = fma_d(r, mc.e4, mc.e3);
    const double p = fma_d(r2, fma_d(r2, mc.e5, p23), p01);
    const double s = tab[ki & 31];
    const double y = fma_d(s * r, p, s);
    // scale by 2^(ki >> 5): add to the exponent field (results here are far from over/underflow)
    return bits2d(d2bits(y) + (static_cast<uint64_t>(static_cast<int64_t>(ki >> 5)) << 52));
}

}  // namespace ofp
