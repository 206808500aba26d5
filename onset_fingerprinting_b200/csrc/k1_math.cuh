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
    double a1, a2, a3, a4, a5;    // log10(1+r) = r (a1 + a2 r + ... + a5 r^4), |r| <= 2^-8
    double log10_2s, kmagic;      // log10(2) * 2^-23 and 2^52 + 2^31 (exponent word -> double without I2F)
    double e1, e2, e3;            // 2^r = 1 + r (e1 + e2 r + e3 r^2), |r| <= 2^-10
    double log2_10;
};
OFP_HD MathConst math_const() {
    MathConst c;
    c.a1 = OFP_LOG_A1; c.a2 = OFP_LOG_A2; c.a3 = OFP_LOG_A3; c.a4 = OFP_LOG_A4; c.a5 = OFP_LOG_A5;
    c.log10_2s = OFP_LOG10_2 * 0x1p-23;
    c.kmagic = 0x1p52 + 0x1p31;
    c.e1 = OFP_EXP_E1; c.e2 = OFP_EXP_E2; c.e3 = OFP_EXP_E3;
    c.log2_10 = OFP_LOG2_10;
    return c;
}

// Rounding windows of the fast paths, in units of 2^-52 relative (double ulps of the result):
// log10: |error| < 2^-41 (2^11 ulps), 10**x: |error| < 2^-46.1 (2^5.9 ulps: truncation 2^-46.7, q log2(10) 2^-48,
// roundings 2^-51.4); the windows leave a factor >= 4.
constexpr float OFP_LOG2_10_F = 3.3219280948873623f;                                   // float32(log2(10))
constexpr float OFP_EXP_MAGIC_F = 1.5f * static_cast<float>(1 << (23 - OFP_EXP_N));   // spacing 2^-OFP_EXP_N
constexpr uint32_t OFP_LOG_WIN = 1u << 13;
constexpr uint32_t OFP_EXP_WIN = 1u << 8;

}  // namespace ofp
