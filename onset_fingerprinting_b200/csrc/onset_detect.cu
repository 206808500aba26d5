// K1: batched multi-channel amplitude onset detector (sm_100a).
//
// Replaces AmplitudeOnsetDetector.__call__/init_minmax_tracker and detect_onsets_amplitude
// (reference detection.py:19-86, 595-840) together with the three ctypes calls per block into
// envelope_follower.so (envelope_follower.c:6-57) and scipy.signal.lfilter (detection.py:499-501).
//
// Design (DESIGN.md "K1"):
//   * time is sequential per channel (non-linear recurrences), so parallelism = channel lanes.
//     One lane per (recording, channel); a warp owns G = floor(32/C) whole recordings because the
//     off-threshold logic couples the channels of a recording -> warp shuffles only.
//   * one warp per CTA: 10k recordings x 3 ch = 1000 warps = 6.8 per SM; single-warp CTAs spread
//     them 6/7 per SM instead of 4/8.
//   * input [R, N, C] is staged by TMA: a 2-D tensor map over (N*C, R) with box (T*C, G) brings the
//     next T samples of all G recordings of the warp with ONE cp.async.bulk.tensor per tile into a
//     warp-private mbarrier ring.  T = 40: the longest tile that keeps 7 warps resident per SM; its row pitch
//     (120 words) leaves the per-sample LDS of the 32 lanes with 3-way bank conflicts -- measured harmless (rows
//     padded to 124 words are conflict-free and equally fast, DESIGN.md "K1"), the LSU is 11 % busy.
//   * the rel envelope of the current block stays in shared memory: the thresholds of a block
//     depend on the END-of-block min/max, so crossings are found by a second pass that
//     only runs when the block maximum exceeded the on-threshold (once per hit).  rel leaves for HBM from
//     that buffer as one bulk async copy per recording and block (cp.async.bulk shared -> global).
//   * the chunk loop is laid out by hand (nested goto loops: the common path is one straight line) and the
//     shared-window base is an opaque register (no S2UR + ULEA rebuild of shared addresses inside loops).
//   * arithmetic follows SURVEY Appendix A op for op: explicit __f*_rn intrinsics (never contracted
//     to FMA), one double add inside the follower, log10/10**x evaluated in double and rounded once.
#include "ofp_common.cuh"

#include <chrono>
#include <mutex>
#include "k1_math.cuh"

// Speed-of-light ladder for profiles/ (scripts/k1_ladder.sh): 0 = TMA in + rel out only, 1 = + high-pass,
// 2 = + dB, 3 = + followers, 4 = + 10**x, 5.. = the full detector (default).  Anything below 5 computes
// something else than the detector and only exists to time the stages.
#ifndef OFP_K1_LADDER
#define OFP_K1_LADDER 9
#endif
#ifndef OFP_K1_KU
#define OFP_K1_KU 8
#endif
#ifndef OFP_K1_SPARSE  // short-cut (v) of chunk_loop, switchable for A/B builds
#define OFP_K1_SPARSE 1
#endif
#ifndef OFP_K1_KQF  // 10**x: k by a float32 magic-constant add (+ one conversion) instead of a double one
#define OFP_K1_KQF 0  // measured: 59.1 ms with the conversion, 57.9 ms without (the XU pipe is the busier resource)
#endif
#ifndef OFP_K1_ICVT_IN  // 10**x: double(q) by integer operations instead of F2F
#define OFP_K1_ICVT_IN 0
#endif
#ifndef OFP_K1_ICVT_OUT  // 10**x: float32 rounding of the result by integer operations instead of F2F
#define OFP_K1_ICVT_OUT 0
#endif
#ifndef OFP_K1_KMAGIC  // log10: exponent -> double by a magic-constant subtraction instead of I2F
#define OFP_K1_KMAGIC 1
#endif
#ifndef OFP_K1_LAUNDER_TAB
#define OFP_K1_LAUNDER_TAB 1  // measured: 55.0 ms with the opaque base, 56.4 ms without
#endif
#ifndef OFP_K1_FOLMAX  // followers: coef * d as max(att * d, rel * d) instead of a select
#define OFP_K1_FOLMAX 1
#endif

#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <cstring>

struct ofp_detector {
    ofp_detector_params p;
    int64_t n_streams;
    int64_t n_lanes;
    float *buf;  // 11 arrays of n_lanes 32-bit words
};

namespace ofp {

// shared memory header: 128 B of mbarriers, then the log10 / 10**x tables
constexpr int K1_SMEM_HEADER = 128 + (2 << OFP_LOG_N) * 8 + (1 << OFP_EXP_N) * 8;

struct DetState {
    float *z0, *z1, *z2, *z3, *yf, *ys, *mn, *mx, *prev;
    int32_t *state, *deb;
};

static DetState state_of(const ofp_detector *d) {
    DetState s;
    float *b = d->buf;
    const int64_t n = d->n_lanes;
    s.z0 = b; s.z1 = b + n; s.z2 = b + 2 * n; s.z3 = b + 3 * n;
    s.yf = b + 4 * n; s.ys = b + 5 * n; s.mn = b + 6 * n; s.mx = b + 7 * n; s.prev = b + 8 * n;
    s.state = reinterpret_cast<int32_t *>(b + 9 * n);
    s.deb = reinterpret_cast<int32_t *>(b + 10 * n);
    return s;
}

struct K1Args {
    ofp_detector_params p;
    float ia_min, ia_max;  // float(1.0 - (double)alpha), envelope_follower.c:31-32
    DetState st;
    const float *x;
    int64_t n_samples, rec_stride;
    int64_t warm_n;  // warm-up region [0, warm_n): HP over all of it, envelopes over full blocks
    int64_t n_main;  // main region [0, n_main), n_main = n_blocks * B
    float *rel;
    int64_t rel_stride;
    int32_t *on_ch, *on_idx, *on_cnt;
    int32_t cap;
    int32_t R, G, T, TC, nst, stage_floats, stride_rel, rel_vec_ok;
    int32_t P;  // row pitch of a stage in floats: TC + padding columns (the TMA box is P wide, tiles advance by TC)
    int32_t floor_skip;  // OFP_K1_FLOOR_SKIP=0 disables the below-floor short-cut (A/B runs)
    int32_t fast_ok;     // follower coefficients in range: the straight-line chunk's short-cuts are proven for those
    int32_t cnt_in;      // continue: the per-recording onset counts in on_cnt are the starting fill levels
    int64_t blk0;        // continue: global index of the first block of this call (x / rel start there)
};

// Per-kernel constants held in registers for the whole launch (a one-warp CTA has registers to
// spare; re-loading them from the constant bank inside the recurrences costs LDC latency).
struct Coef {
    float b0, b1, b2, b3, b4, a1, a2, a3, a4;
    float fa, fr, sa, sr;
    float fA, fR, fS, sA, sR, sS;  // followers of the straight-line chunk: (s att, s rel, s), s = att >= rel ? 1 : -1
    float floor_db, ceil_amp, vfloor, vfloor_h;
    float mxfac;       // chunk maximum < mxfac x max tracker => no sample of the chunk exceeds the tracker
    float overshoot;   // != 0: an attack coefficient above 1 (envelopes are followed through non-skip chunks)
    float sliver_thr;  // -4, or -inf when the straight-line chunk must not be trusted (every chunk re-runs exactly)
    float amin, amax, iamin, iamax, minmin;
};
__device__ __forceinline__ Coef load_coef(const K1Args &a) {
    Coef k;
    k.b0 = a.p.b[0]; k.b1 = a.p.b[1]; k.b2 = a.p.b[2]; k.b3 = a.p.b[3]; k.b4 = a.p.b[4];
    k.a1 = a.p.a[1]; k.a2 = a.p.a[2]; k.a3 = a.p.a[3]; k.a4 = a.p.a[4];
    k.fa = a.p.fast_att; k.fr = a.p.fast_rel; k.sa = a.p.slow_att; k.sr = a.p.slow_rel;
    k.fS = k.fa >= k.fr ? 1.0f : -1.0f; k.fA = k.fS * k.fa; k.fR = k.fS * k.fr;
    k.sS = k.sa >= k.sr ? 1.0f : -1.0f; k.sA = k.sS * k.sa; k.sR = k.sS * k.sr;
    k.floor_db = a.p.floor_db; k.ceil_amp = -a.p.floor_db;
    // |h + 1e-10| below vfloor => 20*log10(.) rounds below the floor => the clipped dB value IS the floor
    // (4e-6 relative margin = 4.5 float32 ulps of the floor in dB, DESIGN.md "K1 arithmetic" (iv))
    k.vfloor = a.floor_skip ? static_cast<float>(exp10(static_cast<double>(a.p.floor_db) / 20.0) * (1.0 - 4e-6)) : 0.0f;
    k.vfloor_h = k.vfloor > 0.0f ? nextafterf(static_cast<float>(static_cast<double>(k.vfloor) * (1.0 - 0x1p-23) - 1e-10), 0.0f) : 0.0f;
    {
        const double ia = static_cast<double>(a.ia_max);
        k.mxfac = (ia > 0.0 && ia < 1.0) ? static_cast<float>(pow(ia, OFP_K1_KU) * (1.0 - 4e-6)) : 0.0f;
    }
    k.overshoot = (k.fa > 1.0f || k.fr > 1.0f || k.sa > 1.0f || k.sr > 1.0f) ? 1.0f : 0.0f;
    k.sliver_thr = a.fast_ok ? -4.0f : -INFINITY;
    k.amin = a.p.alpha_min; k.amax = a.p.alpha_max; k.iamin = a.ia_min; k.iamax = a.ia_max;
    k.minmin = a.p.minmin;
    return k;
}

// Shared memory is addressed with explicit 32-bit shared-space addresses (no generic-pointer
// window arithmetic in the hot loop).  The "memory" clobber keeps these ordered against the
// mbarrier waits / warp syncs that publish the TMA tiles.
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));  // tables are read-only after setup
    return v;
}
__device__ __forceinline__ void lds_2f64(uint32_t addr, double &x, double &y) {
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}

// Rare paths (special inputs, results too close to a float32 rounding boundary): full-precision
// libdevice routines, kept out of line so that the hot loop stays small in the instruction cache.
__device__ __noinline__ double slow_log10(float v) { return log10(static_cast<double>(v)); }
__device__ __noinline__ double slow_exp10(float q) { return exp10(static_cast<double>(q)); }
__device__ __noinline__ float slow_ar_delta(float t) {
    return __double2float_rn(__dadd_rn(static_cast<double>(t), 1e-10));
}

// envelope_follower.c:15-22 (gcc: subss, cvtss2sd, addsd, cvtsd2ss, mulss, addss).
// d = float(double(t) + 1e-10) equals the float32 sum t + 1e-10f whenever t == 0 or |t| >= 2^-26
// (verified over 2.7e8 random t, DESIGN.md "K1 arithmetic"); the double path is kept for the
// remaining sliver (|t| < 2^-22), which needs the signal within 3e-8 dB of the envelope.
__device__ __forceinline__ float ar_step(float y, float x, float att, float rel) {
    const float t = __fsub_rn(x, y);
    float d = __fadd_rn(t, 1e-10f);
    if (__builtin_expect(fabsf(t) < 0x1p-22f && t != 0.0f, 0)) d = slow_ar_delta(t);
    const float coef = d > 0.0f ? att : rel;
    return __fadd_rn(y, __fmul_rn(coef, d));
}

struct Lane {
    float z0, z1, z2, z3, yf, ys, mn, mx, prev, bmax, bmin;
    int32_t state, deb;
};

// Same recurrence when the numerator is symmetric (b0 == b4, b1 == b3 bit for bit, true for every
// Butterworth high-pass): the two repeated products are formed once -- identical values, 2 FMUL less.
__device__ __forceinline__ float hp_step_sym(Lane &L, const Coef &k, float x) {
    const float m0 = __fmul_rn(k.b0, x), m1 = __fmul_rn(x, k.b1);
    const float y = __fadd_rn(L.z0, m0);
    L.z0 = __fsub_rn(__fadd_rn(L.z1, m1), __fmul_rn(y, k.a1));
    L.z1 = __fsub_rn(__fadd_rn(L.z2, __fmul_rn(x, k.b2)), __fmul_rn(y, k.a2));
    L.z2 = __fsub_rn(__fadd_rn(L.z3, m1), __fmul_rn(y, k.a3));
    L.z3 = __fsub_rn(m0, __fmul_rn(y, k.a4));
    return y;
}

// scipy lfilter, DF2T, float32, unfused (detection.py:499-501)
__device__ __forceinline__ float hp_step(Lane &L, const Coef &k, float x) {
    const float y = __fadd_rn(L.z0, __fmul_rn(k.b0, x));
    L.z0 = __fsub_rn(__fadd_rn(L.z1, __fmul_rn(x, k.b1)), __fmul_rn(y, k.a1));
    L.z1 = __fsub_rn(__fadd_rn(L.z2, __fmul_rn(x, k.b2)), __fmul_rn(y, k.a2));
    L.z2 = __fsub_rn(__fadd_rn(L.z3, __fmul_rn(x, k.b3)), __fmul_rn(y, k.a3));
    L.z3 = __fsub_rn(__fmul_rn(x, k.b4), __fmul_rn(y, k.a4));
    return y;
}

// detection.py:747-748: clip(20*log10(|h + 1e-10|), floor) in float32; log10 correctly rounded.
__device__ __forceinline__ float db_of(double ld, float floor_db) {
    return fmaxf(__fmul_rn(20.0f, __double2float_rn(ld)), floor_db);
}
// detection.py:753-754: clip(10**(r/20) - 1e-10, 0, -floor) in float32; 10**x correctly rounded.
__device__ __forceinline__ float amp_of(double ad, float ceil_amp) {
    const float amp = __fsub_rn(__double2float_rn(ad), 1e-10f);
    return fminf(fmaxf(amp, 0.0f), ceil_amp);
}

// Issue costs on B200 (profiles/r02_ubench_op_rates.txt, SMSP cycles per warp instruction): FP32 1, LOP3 / IADD3 /
// FMNMX 1, SHF / IMAD / FSEL / SEL / ISETP / FMNMX3 2, FP64 2, F2F / I2F conversions 8.5 -- and no dual issue: the
// costs add up.  The two pointwise stages below are therefore written to avoid conversions (integer constructions
// of the doubles, integer rounding of the result), per-sample predicate logic (the rare-case tests are ACCUMULATED
// as unsigned min / max words and examined once per chunk) and selects.

// log10 of U samples, step-major (every elementary operation is issued for all U samples before the next one, so
// the instruction stream handed to ptxas is already interleaved).  v = |h + 1e-10|.  Table-driven double
// evaluation, relative error < 2^-41.  Rare cases are accumulated, not branched on:
//   spec = max (bits(v) - 0x00800000): >= 0x7f000000 for zero / denormal / inf / nan inputs;
//   mid  = min distance word of the double result to a float32 rounding boundary (the 29 dropped bits shifted to
//          the top of the word, minus the lower edge of the window): < 2 * OFP_LOG_WIN << 3 when the double cannot
//          be trusted to round correctly (the caller then uses slow_log10).
template <int U>
__device__ __forceinline__ void to_db_vec(const float (&v)[U], float floor_db, uint32_t logtab, const MathConst &mc,
                                          float (&db)[U], uint32_t &spec, uint32_t &mid) {
    uint32_t ix[U], tmp[U], kw[U], iz[U];
    double z[U], kd[U], invc[U], logc[U], r[U], r2[U], p01[U], p23[U], pp[U], base[U], ld[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { ix[u] = __float_as_uint(v[u]); tmp[u] = ix[u] - OFP_LOG_OFF; }
#pragma unroll
    for (int u = 0; u < U; ++u)
        lds_2f64(logtab + ((tmp[u] >> (23 - OFP_LOG_N - 4)) & (((1u << OFP_LOG_N) - 1u) << 4)), invc[u], logc[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) { kw[u] = tmp[u] & 0xff800000u; iz[u] = ix[u] - kw[u]; }
#pragma unroll
    for (int u = 0; u < U; ++u)  // z in [OFF, 2 OFF) as a double: exponent rebias 127 -> 1023
        z[u] = __hiloint2double(static_cast<int>((iz[u] >> 3) + 0x38000000u), static_cast<int>(iz[u] << 29));
#pragma unroll
    for (int u = 0; u < U; ++u)  // exponent * 2^23 as a double: 2^52 + (kw ^ 2^31) minus the magic constant, exact
#if OFP_K1_KMAGIC
        kd[u] = __dsub_rn(__hiloint2double(0x43300000, static_cast<int>(kw[u] ^ 0x80000000u)), mc.kmagic);
#else
        kd[u] = static_cast<double>(static_cast<int32_t>(kw[u]));
#endif
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = __fma_rn(z[u], invc[u], -1.0);
#pragma unroll
    for (int u = 0; u < U; ++u) base[u] = __fma_rn(kd[u], mc.log10_2s, logc[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) r2[u] = __dmul_rn(r[u], r[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) p01[u] = __fma_rn(r[u], mc.a2, mc.a1);
#pragma unroll
    for (int u = 0; u < U; ++u) p23[u] = __fma_rn(r[u], mc.a4, mc.a3);
#pragma unroll
    for (int u = 0; u < U; ++u) pp[u] = __fma_rn(r2[u], mc.a5, p23[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) pp[u] = __fma_rn(r2[u], pp[u], p01[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) ld[u] = __fma_rn(r[u], pp[u], base[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        spec = max(spec, ix[u] - 0x00800000u);
        mid = min(mid, (static_cast<uint32_t>(__double2loint(ld[u])) << 3) - ((0x10000000u - OFP_LOG_WIN) << 3));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) db[u] = db_of(ld[u], floor_db);
}
__device__ __forceinline__ bool db_flagged(uint32_t spec, uint32_t mid) {
    return (spec >= 0x7f000000u) | (mid < ((2u * OFP_LOG_WIN) << 3));
}

// 10**(dr/20) - 1e-10 clipped to the ceiling.  q = dr / 20 correctly rounded without a division (exact: host harness
// over 4e8 values); k = rint(q log2(10) 2^N) by a float32 magic-constant add, reduced argument
// rr = q log2(10) - k / 2^N by one double FMA, 2^(j / 2^N) from the table, degree-3 polynomial, relative error
// < 2^-46.1; the result is rounded to float32 and scaled by 2^(k >> N) in the exponent field.  Rare cases:
//   |q| >= 9.5 (the fast path needs 10**q - 1e-10 > 0 and no overflow of the exponent arithmetic);
//   mid = min distance word of the double result to a float32 rounding boundary: < 2 * OFP_EXP_WIN << 3 => slow_exp10.
// Front of the evaluation for ONE sample (float32 work + two conversions; depends only on dr): a separate piece so
// that the straight-line chunk can issue it between the dependent steps of the follower recurrences.
struct AmpFront {
    float q;        // dr / 20, correctly rounded
    uint32_t ki;    // low mantissa bits: k = rint(q log2(10) 2^N)
    double qd, kq;  // q and k / 2^N as doubles
};
__device__ __forceinline__ AmpFront amp_front(float dr) {
    AmpFront f;
    const float q0 = __fmul_rn(dr, 0.05f);
    f.q = __fmaf_rn(__fmaf_rn(-20.0f, q0, dr), 0.05f, q0);
#if OFP_K1_ICVT_IN
    {   // zero / denormal q become 2^-127-sized doubles: harmless; inf / nan are excluded by the |q| bound
        const uint32_t qb = __float_as_uint(f.q);
        const uint32_t hi = (((qb & 0x7fffffffu) >> 3) + 0x38000000u) | (qb & 0x80000000u);
        f.qd = __hiloint2double(static_cast<int>(hi), static_cast<int>(qb << 29));
    }
#else
    f.qd = static_cast<double>(f.q);
#endif
    // the integer k sits in the low mantissa bits of the sum; a k that is off by one near a tie only makes |rr|
    // 2^-19 larger
#if OFP_K1_KQF
    const float kf0 = __fmaf_rn(f.q, OFP_LOG2_10_F, OFP_EXP_MAGIC_F);
    f.ki = __float_as_uint(kf0);
    f.kq = static_cast<double>(__fsub_rn(kf0, OFP_EXP_MAGIC_F));
#else  // the same in double: one conversion less, one more FP64 issue slot
    const double kd0 = __fma_rn(f.qd, OFP_LOG2_10, 0x1.8p52 / static_cast<double>(1 << OFP_EXP_N));
    f.ki = static_cast<uint32_t>(__double2loint(kd0));
    f.kq = __dsub_rn(kd0, 0x1.8p52 / static_cast<double>(1 << OFP_EXP_N));
#endif
    return f;
}
// Back of the evaluation for U samples, step-major (every elementary operation is issued for all U samples before
// the next one): table, reduced argument, polynomial, rounding, scaling.
template <int U>
__device__ __forceinline__ void amp_back(const AmpFront (&f)[U], float ceil_amp, uint32_t exptab, const MathConst &mc,
                                         float (&amp)[U], uint32_t &mid) {
    uint32_t fb[U];
    double rr[U], sc[U], pp[U], s1[U], y[U];
#pragma unroll
    for (int u = 0; u < U; ++u) sc[u] = lds_f64(exptab + ((f[u].ki & ((1u << OFP_EXP_N) - 1u)) << 3));
#pragma unroll
    for (int u = 0; u < U; ++u) rr[u] = __fma_rn(f[u].qd, mc.log2_10, -f[u].kq);
#pragma unroll
    for (int u = 0; u < U; ++u) pp[u] = __fma_rn(rr[u], mc.e3, mc.e2);
#pragma unroll
    for (int u = 0; u < U; ++u) s1[u] = __dmul_rn(sc[u], rr[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) pp[u] = __fma_rn(rr[u], pp[u], mc.e1);
#pragma unroll
    for (int u = 0; u < U; ++u) y[u] = __fma_rn(s1[u], pp[u], sc[u]);  // in (0.99, 2.01)
#pragma unroll
    for (int u = 0; u < U; ++u) {
#if OFP_K1_ICVT_OUT
        // + 2^28 in the low word (half a float32 ulp), carry into the high word together with the exponent rebias
        // 1023 -> 127 (0x08000000 << 3 == -896 << 23 mod 2^32), funnel shift: round-to-nearest-even except on exact
        // ties, which the window excludes
        uint32_t lo2, hi2;
        asm("add.cc.u32 %0, %2, 0x10000000;\n\taddc.u32 %1, %3, 0x08000000;"
            : "=r"(lo2), "=r"(hi2)
            : "r"(static_cast<uint32_t>(__double2loint(y[u]))), "r"(static_cast<uint32_t>(__double2hiint(y[u]))));
        fb[u] = __funnelshift_l(lo2, hi2, 3) + ((f[u].ki & ~((1u << OFP_EXP_N) - 1u)) << (23 - OFP_EXP_N));
        mid = min(mid, (lo2 << 3) + (OFP_EXP_WIN << 3));
#else
        fb[u] = __float_as_uint(__double2float_rn(y[u])) + ((f[u].ki & ~((1u << OFP_EXP_N) - 1u)) << (23 - OFP_EXP_N));
        mid = min(mid, (static_cast<uint32_t>(__double2loint(y[u])) << 3) + ((0x10000000u + OFP_EXP_WIN) << 3));
#endif
    }
#pragma unroll
    for (int u = 0; u < U; ++u) amp[u] = fminf(__fsub_rn(__uint_as_float(fb[u]), 1e-10f), ceil_amp);
}
__device__ __forceinline__ bool amp_flagged(float qmax, uint32_t mid) {
    return !(qmax < 9.5f) | (mid < ((2u * OFP_EXP_WIN) << 3));
}

// The previous block's bulk copy (block_end) must have read the block buffer before it is overwritten.
// (Unconditional: with nothing in flight the wait is a scoreboard test that falls through; a branch around it was
// one more taken branch in every chunk.)
__device__ __forceinline__ void wait_rel(bool &rel_pending) {
    bulk_wait_read<0>();
    __syncwarp();
    rel_pending = false;
}

// envelope_follower.c:38-52
__device__ __forceinline__ void minmax_step(Lane &L, const Coef &k, float r) {
    const float nm = __fadd_rn(__fmul_rn(L.mn, k.iamin), __fmul_rn(r, k.amin));
    L.mn = r < k.minmin ? k.minmin : (r < L.mn ? r : nm);
    const float nx = __fadd_rn(__fmul_rn(L.mx, k.iamax), __fmul_rn(r, k.amax));
    L.mx = r > L.mx ? r : nx;
}

// ONE sample of one lane, exact in every case (branches, slow paths after a warp vote): the reference semantics
// the straight-line chunk below falls back to, and the path of block tails.
//   xs: shared address of the lane's input sample, rs: of its rel slot.
template <bool USE_HP>
__device__ __forceinline__ void sample_exact(Lane &L, const Coef &k, uint32_t xs, uint32_t rs, bool do_minmax,
                                             bool store, uint32_t logtab, uint32_t exptab, const MathConst &mc,
                                             bool &rel_pending) {
    const float x = lds_f32(xs);
    const float h = USE_HP ? hp_step(L, k, x) : x;
    float v[1] = {fabsf(__fadd_rn(h, 1e-10f))}, db[1], dr[1], amp[1], q[1];
    uint32_t spec = 0, mid = 0xffffffffu;
    to_db_vec<1>(v, k.floor_db, logtab, mc, db, spec, mid);
    const bool redo_db = db_flagged(spec, mid);
    if (__any_sync(0xffffffffu, redo_db)) {
        if (redo_db) db[0] = db_of(slow_log10(v[0]), k.floor_db);
    }
    // detection.py:751 (envelope_follower.c:6-25 twice)
    L.yf = ar_step(L.yf, db[0], k.fa, k.fr);
    L.ys = ar_step(L.ys, db[0], k.sa, k.sr);
    dr[0] = __fsub_rn(L.yf, L.ys);
    mid = 0xffffffffu;
    AmpFront af[1] = {amp_front(dr[0])};
    q[0] = af[0].q;
    amp_back<1>(af, k.ceil_amp, exptab, mc, amp, mid);
    const bool redo_amp = amp_flagged(fabsf(q[0]), mid);
    if (__any_sync(0xffffffffu, redo_amp)) {
        if (redo_amp) {
            const float qq = fabsf(q[0]) < 30.0f ? q[0] : __fdiv_rn(dr[0], 20.0f);
            amp[0] = amp_of(slow_exp10(qq), k.ceil_amp);
        }
    }
    if (do_minmax) minmax_step(L, k, amp[0]);
    L.bmax = fmaxf(L.bmax, amp[0]);
    L.bmin = fminf(L.bmin, amp[0]);
    wait_rel(rel_pending);
    if (store) sts_f32(rs, amp[0]);
}

// Back half of a chunk (see chunk_loop): 10**x, block extrema, stores of rel into the block buffer, and the vote
// on short-cuts (vi) + (vii).  Returns true when the vote passed (the caller then sets L.mn / L.mx from minmin /
// mx_spec), false when the trackers have to be stepped sample by sample over amp[].
template <int U, bool DO_MM>
__device__ __forceinline__ bool chunk_tail(Lane &L, const Coef &k, const AmpFront (&af)[U], const float (&dr)[U],
                                           float (&amp)[U], float &mx_spec, uint32_t rp, uint32_t st, bool store,
                                           uint32_t exptab, const MathConst &mc, bool &rel_pending, bool &bad) {
    if (OFP_K1_LADDER == 3) {  // speed-of-light ladder: + followers
        wait_rel(rel_pending);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (store) sts_f32(rp + u * st, dr[u]);
        return true;
    }
    // |q| < 9.5 needs no test here: both envelopes stay within [floor, -3.99] through the chunk (floor invariant and
    // sliver test in chunk_loop; the host only enables this path for floor >= -180 dB), so |dr| < 190
    uint32_t mid = 0xffffffffu;
    amp_back<U>(af, k.ceil_amp, exptab, mc, amp, mid);
    bad |= amp_flagged(0.0f, mid);
    float cmax = amp[0], cmin = amp[0];
    const float mx0 = L.mx;
    if (OFP_K1_LADDER >= 5) {  // 4: + 10**x only; 5 and above: the full chunk
        // max tracker, block extrema and the stores first: they overlap the drain of the 10**x pipeline that the
        // vote has to wait for
#pragma unroll
        for (int u = 1; u + 1 < U; u += 2) {
            cmax = fmaxf(fmaxf(cmax, amp[u]), amp[u + 1]);
            cmin = fminf(fminf(cmin, amp[u]), amp[u + 1]);
        }
        if (U % 2 == 0) { cmax = fmaxf(cmax, amp[U - 1]); cmin = fminf(cmin, amp[U - 1]); }
        // (vii) The max tracker (envelope_follower.c:48-51) runs on the assumption that no sample exceeds it: every
        // step is then the decay branch, no select in the recurrence.  The assumption holds when the chunk maximum is
        // below mxfac = (1 - alpha_max)^U (1 - 4e-6) x the tracker's start value (the decay steps of a chunk cannot
        // take it lower: the samples are >= 0, roundings cost < 2^-23 per step); it is checked together with
        // short-cut (vi) by the one vote below, after the stores, so that the recurrence overlaps the 10**x pipeline.
        mx_spec = mx0;
        if (DO_MM) {
#pragma unroll
            for (int u = 0; u < U; ++u) mx_spec = __fadd_rn(__fmul_rn(mx_spec, k.iamax), __fmul_rn(amp[u], k.amax));
        }
        L.bmax = fmaxf(L.bmax, cmax);
        L.bmin = fminf(L.bmin, cmin);
    }
    wait_rel(rel_pending);
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (store) sts_f32(rp + u * st, amp[u]);
    if (!DO_MM || OFP_K1_LADDER < 5) return true;
    return __all_sync(0xffffffffu, (amp[U - 1] < k.minmin) & (cmax < __fmul_rn(mx0, k.mxfac)));
}

// The chunk loop: U samples per iteration in straight-line form, one large basic block per path that ptxas can
// software-pipeline (needs __launch_bounds__(32, 1): with a higher occupancy target ptxas keeps the dependency
// chains back to back to save registers).  Every rare case is only FLAGGED (see to_db_vec / amp_back, plus follower
// steps that could fall into the sliver 0 < |x - y| < 2^-22 where the float32 short-cut of ar_step is not proven
// exact).  Returns true when this lane hit a flag; the caller then restores the lane state and re-runs the samples
// through sample_exact.  CT: compile-time channel count (0 = use `step`), so that the shared-memory accesses of a
// chunk are one base register plus immediates.
//
// Data-dependent exact short-cuts, all decided by warp votes (DESIGN.md "K1"):
//  (iv)  no sample of the warp above the floor -> the dB values are the floor, no logarithm, and the followers
//        only release;
//  (v)   at most ONE sample per lane above the floor (noise peaks: one value in a few hundred) -> one logarithm
//        per lane (of the lane's maximum) instead of U;
//  (vi)  the last rel value of every lane below `minmin` -> the min tracker ends at `minmin` whatever came
//        before (envelope_follower.c:42-43), its recurrence is skipped;
//  (vii) no sample above the decayed max tracker -> its recurrence has no select.
//
// Layout.  A taken branch costs this kernel about 2.5 % of its time (one or two warps per scheduler: nothing
// hides the refetch).  ptxas orders basic blocks topologically (every forward predecessor of a block before it), so
// an if/else or an if around a slow path always costs the common path one taken branch, whatever __builtin_expect
// says.  The loop is therefore written as two nested loops with gotos: the inner one is the common path -- (iv),
// (vi) and (vii) all hold -- in one straight line from `top` to its back edge; every other case LEAVES the inner
// loop, finishes its chunk behind it (with its own copy of chunk_tail) and re-enters through `outer`.  Every local
// is declared before the first label (a goto must not cross an initialisation).
template <bool USE_HP, bool HP_SYM, int U, bool DO_MM, int CT>
__device__ __forceinline__ bool chunk_loop(Lane &L, const Coef &k, float (&xin)[U], int &i_io, const int nfast,
                                           uint32_t &xp_io, uint32_t &rp_io, uint32_t step, bool store,
                                           uint32_t logtab, uint32_t exptab, const MathConst &mc, bool &rel_pending) {
    const uint32_t st = CT ? 4u * CT : step;
    constexpr bool TRACK = DO_MM && OFP_K1_LADDER >= 5;
    float h[U], v[U], db[U], dr[U], amp[U];
    AmpFront af[U];  // fronts of the 10**x evaluations, issued inside the follower loops
    uint32_t spec, mid, xp, rp;
    float hmax, dbmax, m1, m2, mx_spec, ytop;
    bool bad, skip;
    int i;
    i = i_io; xp = xp_io; rp = rp_io;
    bad = false;
outer:
    if (i >= nfast) goto done;
top:
#pragma unroll
    for (int u = 0; u < U; ++u)
        h[u] = (USE_HP && OFP_K1_LADDER >= 1) ? (HP_SYM ? hp_step_sym(L, k, xin[u]) : hp_step(L, k, xin[u])) : xin[u];
    // the NEXT chunk's input is fetched now (its shared-memory latency hides behind this chunk); past the end of a
    // tile this reads the neighbouring stage / the block buffer -- defined addresses, values never used
#pragma unroll
    for (int u = 0; u < U; ++u) xin[u] = lds_f32(xp + (U + u) * st);
    if (OFP_K1_LADDER <= 1) {  // speed-of-light ladder (profiles/): memory path only / + high-pass
        wait_rel(rel_pending);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (store) sts_f32(rp + u * st, h[u]);
        goto next;
    }
    // |h| < vfloor_h implies |h + 1e-10| < vfloor (vfloor_h = vfloor (1 - 2^-23) - 1e-10, rounded down)
    hmax = 0.0f;
#pragma unroll
    for (int u = 0; u < U; ++u) hmax = fmaxf(hmax, fabsf(h[u]));
    skip = __all_sync(0xffffffffu, hmax < k.vfloor_h);
    // Followers (envelope_follower.c:15-22).  0 < |x - y| < 2^-22 (the sliver in which float(double(t) + 1e-10)
    // differs from t + 1e-10f) needs min(|x|, |y|) < 2.  With every dB value of the chunk and both envelopes at
    // its start <= -4, the envelopes stay <= -3.99 throughout (a step moves y towards x by a factor <= 1 up to
    // rounding), so one test per chunk is enough; anything else re-runs exactly.
    bad |= !((k.floor_db <= k.sliver_thr) & (L.yf <= k.sliver_thr) & (L.ys <= k.sliver_thr));
    // floor invariant of the envelopes (every dB value is >= floor, a step moves y towards x): holds from the reset
    // on, but a state injected through ofp_detector_set_state may violate it
    bad |= !((L.yf >= k.floor_db) & (L.ys >= k.floor_db));
    if (!skip) goto general;
    // (iv): x is the floor and y >= floor (tested above, and kept through the chunk: release coefficients <= 1/2, a
    // step covers at most half the distance, rounding is monotone and the floor is a float), so
    // d = (floor - y) + 1e-10 <= 1e-10 and the release coefficient applies; d > 0 only for y == floor, where either
    // coefficient (<= 16) leaves y unchanged (|coef d| <= 1.6e-9, far below half an ulp of y <= -3.99), so
    // y + rel * d is the reference's value.
    if (OFP_K1_LADDER == 2) {
        wait_rel(rel_pending);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (store) sts_f32(rp + u * st, k.floor_db);
        goto next;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float d1 = __fadd_rn(__fsub_rn(k.floor_db, L.yf), 1e-10f), d2 = __fadd_rn(__fsub_rn(k.floor_db, L.ys), 1e-10f);
        L.yf = __fadd_rn(L.yf, __fmul_rn(k.fr, d1));
        L.ys = __fadd_rn(L.ys, __fmul_rn(k.sr, d2));
        dr[u] = __fsub_rn(L.yf, L.ys);
        af[u] = amp_front(dr[u]);
    }
    if (!chunk_tail<U, DO_MM>(L, k, af, dr, amp, mx_spec, rp, st, store, exptab, mc, rel_pending, bad)) goto trackers;
    if (TRACK) { L.mn = k.minmin; L.mx = mx_spec; }  // (vi), (vii)
next:
    i += U; xp += U * st; rp += U * st;
    if (i < nfast) goto top;
done:
    i_io = i; xp_io = xp; rp_io = rp;
    return bad;

    // ---- behind the inner loop: a chunk with samples above the floor ----
general:
    spec = 0; mid = 0xffffffffu;
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = fabsf(__fadd_rn(h[u], 1e-10f));
    m1 = v[0]; m2 = 0.0f;  // largest and second largest of the lane
#pragma unroll
    for (int u = 1; u < U; ++u) { m2 = fmaxf(m2, fminf(m1, v[u])); m1 = fmaxf(m1, v[u]); }
    if (OFP_K1_SPARSE && __all_sync(0xffffffffu, m2 < k.vfloor)) {  // (v)
        // the lane's other samples are below the floor; if its maximum is too, db1 comes out as the floor
        float v1[1], db1[1];
        v1[0] = m1;
        to_db_vec<1>(v1, k.floor_db, logtab, mc, db1, spec, mid);
        bad |= db_flagged(spec, mid) & (m1 >= k.vfloor);
        dbmax = m1 >= k.vfloor ? db1[0] : k.floor_db;  // (a special value below the floor must not leak through)
#pragma unroll
        for (int u = 0; u < U; ++u) db[u] = v[u] == m1 ? dbmax : k.floor_db;
    } else {
        to_db_vec<U>(v, k.floor_db, logtab, mc, db, spec, mid);
        bad |= db_flagged(spec, mid);
        dbmax = k.floor_db;
#pragma unroll
        for (int u = 0; u < U; ++u) dbmax = fmaxf(dbmax, db[u]);
    }
    if (OFP_K1_LADDER == 2) {  // + dB
        wait_rel(rel_pending);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (store) sts_f32(rp + u * st, db[u]);
        i += U; xp += U * st; rp += U * st;
        goto outer;
    }
    bad |= !(dbmax <= k.sliver_thr);  // largest dB value of the chunk: the sliver test above
    // coef * d with coef = d > 0 ? att : rel is max(att * d, rel * d) for att >= rel >= 0 (rounding is monotone)
    // and -max(-att * d, -rel * d) for rel > att >= 0: the host passes (A, R, s) = (s att, s rel, s = +-1).
    // An attack coefficient above 1 (the reference's realtime settings use 1 / 0.3) overshoots its input: the
    // envelopes are then followed through the chunk and must stay <= -4 for the sliver argument above (attack
    // coefficients <= 1 cannot overshoot: the chunk-start test covers them, no per-step tracking).
    {
        ytop = -INFINITY;
        auto steps = [&](auto track) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float d1 = __fadd_rn(__fsub_rn(db[u], L.yf), 1e-10f), d2 = __fadd_rn(__fsub_rn(db[u], L.ys), 1e-10f);
#if OFP_K1_FOLMAX
                L.yf = __fmaf_rn(k.fS, fmaxf(__fmul_rn(k.fA, d1), __fmul_rn(k.fR, d1)), L.yf);
                L.ys = __fmaf_rn(k.sS, fmaxf(__fmul_rn(k.sA, d2), __fmul_rn(k.sR, d2)), L.ys);
#else
                L.yf = __fadd_rn(L.yf, __fmul_rn(d1 > 0.0f ? k.fa : k.fr, d1));
                L.ys = __fadd_rn(L.ys, __fmul_rn(d2 > 0.0f ? k.sa : k.sr, d2));
#endif
                if (decltype(track)::value) ytop = fmaxf(fmaxf(ytop, L.yf), L.ys);
                dr[u] = __fsub_rn(L.yf, L.ys);
                af[u] = amp_front(dr[u]);
            }
        };
        if (k.overshoot != 0.0f) { steps(std::true_type{}); bad |= !(ytop <= k.sliver_thr); }
        else steps(std::false_type{});
    }
    if (chunk_tail<U, DO_MM>(L, k, af, dr, amp, mx_spec, rp, st, store, exptab, mc, rel_pending, bad)) {
        if (TRACK) { L.mn = k.minmin; L.mx = mx_spec; }
        i += U; xp += U * st; rp += U * st;
        goto outer;
    }
trackers:  // (vi) or (vii) failed: the reference's recurrences, sample by sample
#pragma unroll
    for (int u = 0; u < U; ++u) minmax_step(L, k, amp[u]);
    i += U; xp += U * st; rp += U * st;
    goto outer;
}

__device__ __align__(16) double g_logtab[2 << OFP_LOG_N];
__device__ __align__(16) double g_exptab[1 << OFP_EXP_N];

// Round-trip the launch constants through shared memory with volatile loads.  To nvcc/ptxas the
// reloaded values are opaque, so they stay in registers; otherwise they are re-materialised inside
// the recurrences as constant-bank loads (LDC, ~30 cycles in a dependent chain) and 64-bit
// immediates (2 UMOV each).  scratch: >= 256 bytes of shared memory not yet in use.
#ifndef OFP_LAUNDER_COEF
#define OFP_LAUNDER_COEF 0
#endif
__device__ __forceinline__ void launder(Coef &k, MathConst &mc, uint32_t scratch) {
    float *f = reinterpret_cast<float *>(&k);
    double *m = reinterpret_cast<double *>(&mc);
    constexpr int NF = OFP_LAUNDER_COEF ? sizeof(Coef) / 4 : 0, ND = sizeof(MathConst) / 8;
#pragma unroll
    for (int i = 0; i < NF; ++i) sts_f32(scratch + 4 * i, f[i]);
#pragma unroll
    for (int i = 0; i < ND; ++i)
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(scratch + 128 + 8 * i), "d"(m[i]) : "memory");
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NF; ++i) asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(f[i]) : "r"(scratch + 4 * i) : "memory");
#pragma unroll
    for (int i = 0; i < ND; ++i)
        asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(m[i]) : "r"(scratch + 128 + 8 * i) : "memory");
    __syncwarp();
}

// End of a block (main phase): the reference's threshold FSM (detection.py:759-792) on the block held in shared
// memory, onset compaction in the reference's order, and the copy of the block's rel envelope to HBM
// (rcol_s = shared address of this lane's column of row 0).  With 16-byte aligned rows the copy is one bulk async copy per recording
// (cp.async.bulk shared -> global, issued by lane 0): the warp does not touch the data again, and
// `rel_pending` tells the next writer of the block buffer to wait until the copy engine has read it.
__device__ __forceinline__ void block_end(Lane &L, const K1Args &a, uint32_t rcol_s, const float *relbuf,
                                          uint32_t relbuf_s, int lane,
                                          int g, int c, int rec, int rec0, bool active, unsigned rec_mask,
                                          unsigned lower_mask, int32_t &cnt, int64_t blk, bool &rel_pending) {
    const int C = a.p.n_channels, B = a.p.block_size, G = a.G;
    // ---- block FSM, detection.py:759-792 ----
    const float last = lds_f32(rcol_s + 4u * ((B - 1) * C));
    const float thr_on = a.p.manual ? a.p.on_thr
                                    : __fadd_rn(__fmul_rn(L.mx, a.p.on_thr), L.mn);
    const float thr_off = a.p.manual ? a.p.off_thr
                                     : __fadd_rn(__fmul_rn(L.mx, a.p.off_thr), L.mn);
    int oi = 0;
    bool hit = false;
    if (!L.state && L.deb < 1 && L.bmax > thr_on) {
        float before = L.prev;
        for (int k = 0; k < B; ++k) {
            const float r = lds_f32(rcol_s + 4u * (k * C));
            if (r > thr_on && before < thr_on) { oi = k; hit = true; break; }
            before = r;
        }
    }
    if (hit) { L.state = 1; L.deb = a.p.cooldown; }
    if (L.deb > 0) L.deb -= B;
    const unsigned hits = __ballot_sync(0xffffffffu, hit && active);
    int M = 0;  // max first-crossing index over the recording's channels (Q3)
    if (hits) {
        for (int jj = 0; jj < C; ++jj) M = max(M, __shfl_sync(0xffffffffu, oi, g * C + jj));
    }
    bool off = false;
    if (M == 0) off = L.bmin < thr_off;
    else {
        for (int k = M; k < B; ++k)
            if (lds_f32(rcol_s + 4u * (k * C)) < thr_off) { off = true; break; }
    }
    if (off) L.state = 0;
    L.prev = last;
    if (hits) {
        const int pos = cnt + __popc(hits & lower_mask);
        if (hit && active && pos < a.cap) {
            a.on_ch[static_cast<int64_t>(rec) * a.cap + pos] = c;
            a.on_idx[static_cast<int64_t>(rec) * a.cap + pos] =
                static_cast<int32_t>(blk * B + oi);
        }
        cnt += __popc(hits & rec_mask);
    }
    if (a.rel != nullptr) {
        const int nBC = B * C;
        if (a.rel_vec_ok) {
            fence_proxy_async();  // this lane's st.shared of the block -> visible to the async proxy
            __syncwarp();
            // lane 0 issues the G copies one after the other: the copy instruction takes uniform operands, so G
            // lanes issuing one copy each become a divergence loop of G iterations anyway, with more bookkeeping
            if (lane == 0) {
                float *dst = a.rel + static_cast<int64_t>(rec0) * a.rel_stride + (blk - a.blk0) * nBC;
                uint32_t src = relbuf_s;
                const int ng = min(G, a.R - rec0);
                for (int gi = 0; gi < ng; ++gi, dst += a.rel_stride, src += 4u * a.stride_rel)
                    bulk_store(dst, src, static_cast<uint32_t>(nBC) * 4u);
            }
            bulk_commit();
            rel_pending = true;
        } else {
            __syncwarp();
            for (int gi = 0; gi < G; ++gi) {
                if (rec0 + gi >= a.R) break;
                float *dst = a.rel + (rec0 + gi) * a.rel_stride + (blk - a.blk0) * nBC;
                const float *src = relbuf + gi * a.stride_rel;
                for (int i = lane; i < nBC; i += 32) __stcs(dst + i, src[i]);
            }
            __syncwarp();
        }
    }
}

constexpr int KU = OFP_K1_KU;  // samples per straight-line chunk of the single-warp kernel

template <bool USE_HP, bool USE_TMA, bool HP_SYM, int CT>
__global__ void __launch_bounds__(32, 1) k1_detect(const __grid_constant__ CUtensorMap tmap, const K1Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    double *logtab = reinterpret_cast<double *>(smem + 128);
    double *exptab = logtab + (2 << OFP_LOG_N);
    float *stages = reinterpret_cast<float *>(smem + K1_SMEM_HEADER);
    float *relbuf = stages + static_cast<size_t>(a.nst) * a.stage_floats;

    const int lane = threadIdx.x;
    const int C = CT ? CT : a.p.n_channels, B = a.p.block_size, G = a.G, T = a.T, TC = a.TC;
    const int g_raw = lane / C;
    const bool in_group = g_raw < G;
    const int g = in_group ? g_raw : 0;
    const int c = in_group ? lane - g_raw * C : 0;
    const int rec0 = blockIdx.x * G;
    const int rec = rec0 + g;
    const bool active = in_group && rec < a.R;
    const int64_t lid = static_cast<int64_t>(rec) * C + c;
    const unsigned rec_mask = (C == 32 ? 0xffffffffu : ((1u << C) - 1u)) << (g * C);
    const unsigned lower_mask = rec_mask & ((1u << lane) - 1u);

    Coef kf = load_coef(a);
    const uint32_t step = 4u * C;
    MathConst mc = math_const();
    launder(kf, mc, smem_u32(relbuf));
#if OFP_K1_LAUNDER_TAB
    // The base of the shared-memory window as an opaque register: ptxas otherwise rebuilds every shared address in
    // the uniform datapath at its use (S2UR of the CTA's rank in the cluster + ULEA, a 20-cycle chain in front of
    // the table loads of every chunk and of the barrier / TMA operations of every tile).
    uint32_t smem_s = smem_u32(smem);
    {
        const uint32_t scratch = smem_u32(relbuf) + 512;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(scratch), "r"(smem_s) : "memory");
        __syncwarp();
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(smem_s) : "r"(scratch) : "memory");
        __syncwarp();
    }
#else
    const uint32_t smem_s = smem_u32(smem);
#endif
    const uint32_t bars_s = smem_s, logtab_s = smem_s + 128, exptab_s = logtab_s + (2 << OFP_LOG_N) * 8;
    const uint32_t stages_s = smem_s + K1_SMEM_HEADER;
    Lane L;
    if (active) {
        L.z0 = a.st.z0[lid]; L.z1 = a.st.z1[lid]; L.z2 = a.st.z2[lid]; L.z3 = a.st.z3[lid];
        L.yf = a.st.yf[lid]; L.ys = a.st.ys[lid]; L.mn = a.st.mn[lid]; L.mx = a.st.mx[lid];
        L.prev = a.st.prev[lid]; L.state = a.st.state[lid]; L.deb = a.st.deb[lid];
    } else {
        L.z0 = L.z1 = L.z2 = L.z3 = 0.f; L.yf = L.ys = a.p.floor_db; L.mn = 0.f; L.mx = 10.f;
        L.prev = 0.f; L.state = 0; L.deb = 0;
    }
    L.bmax = -INFINITY; L.bmin = INFINITY;

    {   // tables -> shared memory: all loads of a lane in flight together (a dependent load / store loop costs one L2
        // round trip per iteration -- 10 us of a 25 us single-block launch of the realtime path)
        constexpr int NL = (2 << OFP_LOG_N) / 64, NE = (1 << OFP_EXP_N) / 64;  // double2 per lane
        double2 tl[NL], te[NE];
#pragma unroll
        for (int i = 0; i < NL; ++i) tl[i] = reinterpret_cast<const double2 *>(g_logtab)[lane + 32 * i];
#pragma unroll
        for (int i = 0; i < NE; ++i) te[i] = reinterpret_cast<const double2 *>(g_exptab)[lane + 32 * i];
#pragma unroll
        for (int i = 0; i < NL; ++i) reinterpret_cast<double2 *>(logtab)[lane + 32 * i] = tl[i];
#pragma unroll
        for (int i = 0; i < NE; ++i) reinterpret_cast<double2 *>(exptab)[lane + 32 * i] = te[i];
    }
    __syncwarp();
    if (USE_TMA) {
        if (lane == 0) {
            for (int s = 0; s < a.nst; ++s) mbar_init(&bars[s], 1);
            fence_mbar_init();
            tma_prefetch_desc(&tmap);
        }
        __syncwarp();
    }
    const int P = a.P;
    const uint32_t box_bytes = static_cast<uint32_t>(G) * P * 4u;
    int32_t cnt = (a.cnt_in && active) ? a.on_cnt[rec] : 0;  // onsets emitted for this lane's recording
    int64_t blk = a.blk0;  // main-phase block index (global; a.blk0 != 0 when a recording is fed in segments)

    const uint32_t relbuf_s = stages_s + static_cast<uint32_t>(a.nst) * 4u * static_cast<uint32_t>(a.stage_floats);
    const uint32_t rcol_s = relbuf_s + 4u * static_cast<uint32_t>(g * a.stride_rel + c);
    const uint32_t stage0_s = stages_s + 4u * (g * P + c);

    bool rel_pending = false;  // a bulk copy of the block buffer to HBM is in flight (block_end)
    int s_cur = 0;             // ring position of the tile being consumed
    uint32_t par = 0;          // its mbarrier phase parity
    const int nst = a.nst;
    const uint32_t stage_bytes = 4u * static_cast<uint32_t>(a.stage_floats);
    for (int phase = 0; phase < 2; ++phase) {
        // both lengths fit 32 bits (int32 sample indices, ofp_detect_offline)
        const int len = static_cast<int>(phase == 0 ? a.warm_n : a.n_main);
        if (len <= 0) continue;
        const int env_len = phase == 0 ? (len / B) * B : len;
        const bool do_minmax = phase == 0 || !a.p.manual;
        const int ntiles = (len + T - 1) / T;
        int kpos = 0;
        if (USE_TMA && lane == 0) {
            int sp = s_cur;
            for (int p = 0; p < nst - 1 && p < ntiles; ++p) {
                mbar_expect_tx(bars_s + 8u * sp, box_bytes);
                tma_load_2d(stages_s + static_cast<uint32_t>(sp) * stage_bytes, &tmap, bars_s + 8u * sp, p * TC, rec0);
                sp = sp + 1 == nst ? 0 : sp + 1;
            }
        }
        int t0 = 0;
        for (int ti = 0; ti < ntiles; ++ti, t0 += T) {
            int s = s_cur;
            if (USE_TMA) {
                const int nx = ti + nst - 1;
                if (lane == 0 && nx < ntiles) {
                    const int sn = s_cur == 0 ? nst - 1 : s_cur - 1;  // the stage consumed last
                    mbar_expect_tx(bars_s + 8u * sn, box_bytes);
                    tma_load_2d(stages_s + static_cast<uint32_t>(sn) * stage_bytes, &tmap, bars_s + 8u * sn, nx * TC, rec0);
                }
                mbar_wait(bars_s + 8u * s, par);
                if (++s_cur == nst) { s_cur = 0; par ^= 1u; }
            } else {
                // generic path (unaligned input): cooperative copy of the tile into stage 0
                s = 0;
                const int64_t row_elems = a.n_samples * C;
                for (int idx = lane; idx < G * P; idx += 32) {
                    const int gi = idx / P, e = idx - gi * P;
                    const int64_t col = static_cast<int64_t>(t0) * C + e;
                    float v = 0.f;
                    if (rec0 + gi < a.R && col < row_elems) v = a.x[(rec0 + gi) * a.rec_stride + col];
                    stages[idx] = v;
                }
                __syncwarp();
            }
            const uint32_t sp = stage0_s + static_cast<uint32_t>(s) * stage_bytes;
            const int tl = min(T, len - t0);
            int j = 0;
            while (j < tl) {
                if (t0 + j < env_len) {
                    const int seg = min(tl - j, B - kpos);
                    uint32_t xp = sp + j * step;
                    uint32_t rp = rcol_s + kpos * step;
                    int i = 0;
                    // Straight-line chunks.  Their rare-case flags are collected over the whole segment (<= one tile,
                    // inside one block) and voted on once: a flagged segment is re-run sample by sample on the exact
                    // path from the state saved at its start (0.2 % of the segments of the benchmark signal).  The
                    // min/max trackers only rest in the main phase of manual-threshold detectors: that choice is made
                    // outside the chunk loop (inside it costs a constant-bank load and a dependent branch per chunk).
                    const int nfast = seg - seg % KU;
                    if (nfast > 0) {
                        const Lane saved = L;
                        const uint32_t xp0 = xp, rp0 = rp;
                        bool bad = false;
                        float xin[KU];  // the chunk's input samples, fetched one chunk ahead
#pragma unroll
                        for (int u = 0; u < KU; ++u) xin[u] = lds_f32(xp + u * step);
                        if (do_minmax)
                            bad = chunk_loop<USE_HP, HP_SYM, KU, true, CT>(L, kf, xin, i, nfast, xp, rp, step, in_group,
                                                                           logtab_s, exptab_s, mc, rel_pending);
                        else
                            bad = chunk_loop<USE_HP, HP_SYM, KU, false, CT>(L, kf, xin, i, nfast, xp, rp, step, in_group,
                                                                            logtab_s, exptab_s, mc, rel_pending);
                        if (__builtin_expect(__any_sync(0xffffffffu, bad), 0)) {
                            L = saved;
                            for (int e = 0; e < nfast; ++e)
                                sample_exact<USE_HP>(L, kf, xp0 + e * step, rp0 + e * step, do_minmax, in_group, logtab_s,
                                                     exptab_s, mc, rel_pending);
                        }
                    }
                    for (; i < seg; ++i, xp += step, rp += step)
                        sample_exact<USE_HP>(L, kf, xp, rp, do_minmax, in_group, logtab_s, exptab_s, mc, rel_pending);
                    j += seg;
                    kpos += seg;
                    if (kpos == B) {
                        kpos = 0;
                        if (phase == 1) {
                            block_end(L, a, rcol_s, relbuf, relbuf_s, lane, g, c, rec, rec0, active, rec_mask, lower_mask,
                                      cnt, blk, rel_pending);
                            ++blk;
                        }
                        L.bmax = -INFINITY; L.bmin = INFINITY;
                    }
                } else {
                    // warm-up tail beyond the last full block: only the high-pass advances
                    // (detection.py:828-829 filters the whole half second in one call)
                    const int seg = tl - j;
                    if (USE_HP) {
                        for (int i = 0; i < seg; ++i) hp_step(L, kf, lds_f32(sp + (j + i) * step));
                    }
                    j += seg;
                }
            }
            __syncwarp();  // every lane is done with stage s before the producer refills it
        }
    }
    if (rel_pending) bulk_wait_read<0>();  // the block buffer must outlive the copy engine's reads

    if (active) {
        a.st.z0[lid] = L.z0; a.st.z1[lid] = L.z1; a.st.z2[lid] = L.z2; a.st.z3[lid] = L.z3;
        a.st.yf[lid] = L.yf; a.st.ys[lid] = L.ys; a.st.mn[lid] = L.mn; a.st.mx[lid] = L.mx;
        a.st.prev[lid] = L.prev; a.st.state[lid] = L.state; a.st.deb[lid] = L.deb;
        if (c == 0 && a.on_cnt != nullptr) a.on_cnt[rec] = cnt;
    }
}


__global__ void k1_reset(DetState st, int64_t n, float floor_db) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    st.z0[i] = st.z1[i] = st.z2[i] = st.z3[i] = 0.f;
    st.yf[i] = st.ys[i] = floor_db;  // detection.py:697-702
    st.mn[i] = 0.f; st.mx[i] = 10.f; // detection.py:703-708
    st.prev[i] = 0.f; st.state[i] = 0; st.deb[i] = 0;  // detection.py:710-712
}

// Twins of the DLL entry points, one thread per channel (envelope_follower.c:6-57).
__global__ void k_ar_envelope(const float *x, float *y, float att, float rel, int size, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    float yi = y[static_cast<int64_t>(n - 1) * size + i];
    for (int j = 0; j < n; ++j) {
        yi = ar_step(yi, x[static_cast<int64_t>(j) * size + i], att, rel);
        y[static_cast<int64_t>(j) * size + i] = yi;
    }
}

__global__ void k_minmax_envelope(const float *x, float *mnp, float *mxp, float a_min, float a_max,
                                  float ia_min, float ia_max, float minmin, int n, int nch) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nch) return;
    float mn = mnp[c], mx = mxp[c];
    for (int i = 0; i < n; ++i) {
        const float r = x[static_cast<int64_t>(i) * nch + c];
        const float nm = __fadd_rn(__fmul_rn(mn, ia_min), __fmul_rn(r, a_min));
        mn = r < minmin ? minmin : (r < mn ? r : nm);
        const float nx = __fadd_rn(__fmul_rn(mx, ia_max), __fmul_rn(r, a_max));
        mx = r > mx ? r : nx;
    }
    mnp[c] = mn; mxp[c] = mx;
}

// backtrack_onsets (detection.py:800-825 == envelope_follower.c:59-85), one thread per onset.
// hist: rel rows in time order, [R, n_rows, C].  The ring buffer of the reference holds the last N
// rows ending with the last row of the onset's block; rows before the start of `hist` read as 0
// (the reference's buffer starts uninitialised, np.empty).
__global__ void k_backtrack(const float *hist, int64_t rec_stride, int64_t n_rows, int C, int B, int N, float alpha,
                            float omba, float tol, int streaming, const int32_t *on_ch, int32_t *on_idx,
                            const int32_t *on_cnt, int R, int cap) {
    const int64_t gid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (gid >= static_cast<int64_t>(R) * cap) return;
    const int r = static_cast<int>(gid / cap), k = static_cast<int>(gid % cap);
    if (k >= min(on_cnt[r], cap)) return;
    const int c = on_ch[gid];
    const int64_t s = on_idx[gid];
    // last row of the block the onset was detected in
    const int64_t end_row = streaming ? n_rows - 1 : (s / B + 1) * B - 1;
    int64_t delta = streaming ? s : s - (s / B) * B;
    const float *col = hist + r * rec_stride + c;
    auto at = [&](int64_t i) -> float {  // buffer[-i]
        const int64_t row = end_row - i + 1;
        return row >= 0 ? col[row * C] : 0.0f;
    };
    int64_t i = B - delta;
    float cur = at(i);
    i += 1;
    float prev = at(i);
    float ps = __fadd_rn(__fmul_rn(alpha, prev), __fmul_rn(omba, cur));
    while (cur > ps && fabsf(__fsub_rn(ps, prev)) > tol && i + 1 < N) {
        delta -= 1;
        i += 1;
        cur = ps;
        prev = at(i);
        ps = __fadd_rn(__fmul_rn(alpha, prev), __fmul_rn(omba, cur));
    }
    on_idx[gid] = static_cast<int32_t>(streaming ? delta : (s / B) * B + delta);
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// Tile length: T*C floats per recording row must be a multiple of 4 (16-byte TMA rows), at most 256
// (TMA box limit) and, if possible, == roundup4(C) (mod 32) so that the G rows of a stage start in
// distinct shared-memory banks.
// Tile length T (samples per TMA box row).  Constraints: T*C*4 bytes a multiple of 16 (TMA), T*C <= 256
// (box limit), T a multiple of `multiple` (so that chunks never straddle tiles).  Among those, prefer
// the T whose row pitch T*C spreads the G rows of a stage over the most shared-memory banks (the
// per-sample LDS of a warp touches one word per lane: rows starting in the same bank conflict), then
// the longest.
static int pick_tile(int C, int tcap = 64, int multiple = 1) {
    const int forced = env_int("OFP_K1_TILE", 0);
    if (forced > 0 && (forced * C) % 4 == 0 && forced * C <= 256 && forced % multiple == 0) return forced;
    const int G = 32 / C;
    const int tmax = std::min(tcap, 256 / C);
    int best = 0, best_deg = 1 << 30;
    for (int t = multiple; t <= tmax; t += multiple) {
        if ((t * C) % 4 != 0) continue;
        int count[32] = {0}, deg = 0;
        for (int g = 0; g < G; ++g)
            for (int c = 0; c < C; ++c) deg = std::max(deg, ++count[(g * t * C + c) % 32]);
        if (deg < best_deg || (deg == best_deg && t > best)) { best = t; best_deg = deg; }
    }
    return best;
}

constexpr int MAX_DEVICES = 64;
static int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < MAX_DEVICES ? dev : 0;
}

// __device__ tables exist once per device: upload on first use of EACH device (a process may drive several)
static int upload_tables() {
    static std::mutex mu;
    static bool done[MAX_DEVICES] = {false};
    std::lock_guard<std::mutex> lock(mu);
    const int dev = current_device();
    if (done[dev]) return OFP_OK;
    OFP_CUDA_CHECK(cudaMemcpyToSymbol(g_logtab, OFP_LOGTAB_H, sizeof(OFP_LOGTAB_H)));
    OFP_CUDA_CHECK(cudaMemcpyToSymbol(g_exptab, OFP_EXPTAB_H, sizeof(OFP_EXPTAB_H)));
    done[dev] = true;
    return OFP_OK;
}

static int launch_k1(ofp_detector *det, const float *x, int64_t n_samples, int64_t rec_stride, int64_t warm_n,
                     int64_t n_main, float *rel, int64_t rel_stride, int32_t *on_ch, int32_t *on_idx,
                     int32_t *on_cnt, int32_t cap, cudaStream_t stream, int64_t blk0 = 0, bool cnt_in = false) {
    const ofp_detector_params &p = det->p;
    const int C = p.n_channels, B = p.block_size;
    K1Args a;
    memset(&a, 0, sizeof a);
    a.p = p;
    a.ia_min = static_cast<float>(1.0 - static_cast<double>(p.alpha_min));
    a.ia_max = static_cast<float>(1.0 - static_cast<double>(p.alpha_max));
    a.floor_skip = env_int("OFP_K1_FLOOR_SKIP", 1);
    {
        // release coefficients in (0, 1/2] (a release step covers at most half the distance to its input: the floor
        // invariant), attack coefficients in (0, 16] (att * 1e-10 stays far below half an ulp of the floor)
        a.fast_ok = (p.fast_rel > 0.0f && p.fast_rel <= 0.5f && p.slow_rel > 0.0f && p.slow_rel <= 0.5f &&
                     p.fast_att > 0.0f && p.fast_att <= 16.0f && p.slow_att > 0.0f && p.slow_att <= 16.0f) ? 1 : 0;
        a.fast_ok &= (p.floor_db >= -180.0f && p.floor_db <= -4.0f) ? 1 : 0;
    }
    a.blk0 = blk0; a.cnt_in = cnt_in ? 1 : 0;
    a.st = state_of(det);
    a.x = x; a.n_samples = n_samples; a.rec_stride = rec_stride;
    a.warm_n = std::min(warm_n, n_samples);
    a.n_main = n_main;
    a.rel = rel; a.rel_stride = rel_stride;
    a.on_ch = on_ch; a.on_idx = on_idx; a.on_cnt = on_cnt; a.cap = cap;
    a.R = static_cast<int32_t>(det->n_streams);
    a.G = 32 / C;
    const int tile_mult = (B % KU == 0) ? KU : 1;
    a.T = pick_tile(C, env_int("OFP_K1_TILECAP", 48), tile_mult);
    OFP_REQUIRE(a.T > 0, "no valid tile length for %d channels", C);
    if (env_int("OFP_K1_TILE", 0) <= 0) {
        // The tile length also decides how many one-warp CTAs an SM holds (stages + block buffer + tables).  A grid
        // that needs k CTAs per SM to be resident at once should get them: with one fewer the launch runs in two
        // waves (2 and 4 channels: 72-74 ms instead of ~52 for the 10 000 x 5 s batch).  Among the valid tiles take
        // the one with the highest useful residency, longer tiles first.
        const int grid_ctas = (a.R + a.G - 1) / a.G, need = (grid_ctas + sm_count() - 1) / sm_count();
        const int want_rel = ((C + 3) / 4 * 4) % 32, bc4r = (B * C + 3) / 4 * 4;
        const size_t fixed = K1_SMEM_HEADER + static_cast<size_t>(a.G) * (bc4r + ((want_rel - bc4r % 32) + 32) % 32) * 4;
        const int nst_env = std::max(2, std::min(8, env_int("OFP_K1_STAGES", 2)));
        int best_t = a.T, best_r = -1;
        for (int cap_t = env_int("OFP_K1_TILECAP", 48); cap_t >= tile_mult; cap_t -= tile_mult) {
            const int t = pick_tile(C, cap_t, tile_mult);
            if (t <= 0) continue;
            const size_t stage = (static_cast<size_t>(a.G) * t * C * 4 + 127) / 128 * 128;
            const int resident = static_cast<int>((227 * 1024) / (fixed + nst_env * stage + 1024));
            const int useful = std::min(resident, need);
            if (useful > best_r) { best_r = useful; best_t = t; }
        }
        a.T = best_t;
    }
    a.TC = a.T * C;
    {
        // pad the rows of a stage so that the G rows start 4 banks apart (the per-sample LDS of a warp touches one
        // word per lane: rows starting in the same bank conflict); the TMA box simply reads the extra columns
        const int want = ((C + 3) / 4 * 4) % 32;
        // rows `want` banks apart in either direction are equally good: take the smaller pad
        int pad = std::min(((want - a.TC % 32) + 32) % 32, ((32 - want - a.TC % 32) + 64) % 32);
        // off by default: at T = 40 the padded stages cost the seventh resident CTA per SM (84 vs 59 ms) and the
        // 3-way conflicts of the unpadded rows are not visible in the kernel time; T = 32 padded is equally fast
        if (!env_int("OFP_K1_PAD", 0) || a.TC + pad > 256 || (a.TC + pad) % 4 != 0) pad = 0;
        a.P = a.TC + pad;
    }
    a.nst = std::max(2, std::min(8, env_int("OFP_K1_STAGES", 2)));
    const int stage_bytes = (a.G * a.P * 4 + 127) / 128 * 128;
    a.stage_floats = stage_bytes / 4;
    const bool tma_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (a.R == 1 || rec_stride % 4 == 0) &&
                        (n_samples * C < (1ll << 31)) && !env_int("OFP_K1_NO_TMA", 0);
    const int NR = B;
    const int want = ((C + 3) / 4 * 4) % 32;
    const int bc4 = (NR * C + 3) / 4 * 4;
    a.stride_rel = bc4 + ((want - bc4 % 32) + 32) % 32;
    a.rel_vec_ok = rel != nullptr && (reinterpret_cast<uintptr_t>(rel) % 16 == 0) && (rel_stride % 4 == 0) &&
                   ((B * C) % 4 == 0);
    const size_t smem = K1_SMEM_HEADER + static_cast<size_t>(a.nst) * stage_bytes + static_cast<size_t>(a.G) * a.stride_rel * 4;
    OFP_REQUIRE(smem <= 227 * 1024, "block_size %d x %d channels needs %zu bytes of shared memory per warp (max 232448)",
                B, C, smem);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    if (tma_ok) {
        const uint64_t stride1 = a.R == 1 ? static_cast<uint64_t>((n_samples * C * 4 + 15) / 16 * 16)
                                          : static_cast<uint64_t>(rec_stride) * 4;
        int rc = encode_tmap_2d_f32(&tmap, x, static_cast<uint64_t>(n_samples) * C, static_cast<uint64_t>(a.R),
                                    stride1, static_cast<uint32_t>(a.P), static_cast<uint32_t>(a.G));
        if (rc != OFP_OK) return rc;
    }
    { int rc = upload_tables(); if (rc != OFP_OK) return rc; }
    const int grid = (a.R + a.G - 1) / a.G;
    const bool sym = p.use_hp && memcmp(&p.b[0], &p.b[4], 4) == 0 && memcmp(&p.b[1], &p.b[3], 4) == 0;
    // channel count as a compile-time constant for the common 3-microphone layout (immediate shared-memory offsets)
    auto kern = C == 3 ? (p.use_hp ? (tma_ok ? (sym ? k1_detect<true, true, true, 3> : k1_detect<true, true, false, 3>)
                                             : (sym ? k1_detect<true, false, true, 3> : k1_detect<true, false, false, 3>))
                                   : (tma_ok ? k1_detect<false, true, false, 3> : k1_detect<false, false, false, 3>))
                       : (p.use_hp ? (tma_ok ? (sym ? k1_detect<true, true, true, 0> : k1_detect<true, true, false, 0>)
                                             : (sym ? k1_detect<true, false, true, 0> : k1_detect<true, false, false, 0>))
                                   : (tma_ok ? k1_detect<false, true, false, 0> : k1_detect<false, false, false, 0>));
    OFP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, 32, smem, stream>>>(tmap, a);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

struct HostCache {
    float *x[2] = {nullptr, nullptr}, *rel[2] = {nullptr, nullptr};
    size_t x_cap[2] = {0, 0}, rel_cap[2] = {0, 0};
    int32_t *och = nullptr, *oix = nullptr, *ocn = nullptr;
    size_t och_cap = 0, oix_cap = 0, ocn_cap = 0;
    cudaStream_t copy = nullptr, comp = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
};
static HostCache g_host_cache[MAX_DEVICES];  // staging buffers, streams and events live on ONE device each

}  // namespace ofp

using namespace ofp;

extern "C" {

int ofp_host_release(void) {
    HostCache &c = g_host_cache[current_device()];
    for (int i = 0; i < 2; ++i) {
        cudaFree(c.x[i]); cudaFree(c.rel[i]);
        if (c.copied[i]) cudaEventDestroy(c.copied[i]);
        if (c.done[i]) cudaEventDestroy(c.done[i]);
    }
    cudaFree(c.och); cudaFree(c.oix); cudaFree(c.ocn);
    if (c.copy) cudaStreamDestroy(c.copy);
    if (c.comp) cudaStreamDestroy(c.comp);
    c = HostCache();
    return OFP_OK;
}

int ofp_detector_create(ofp_detector **out, int64_t n_streams, const ofp_detector_params *p) {
    OFP_REQUIRE(out && p, "null argument");
    OFP_REQUIRE(n_streams > 0 && n_streams < (1ll << 31), "n_streams out of range");
    OFP_REQUIRE(p->n_channels >= 1 && p->n_channels <= 32, "n_channels must be in 1..32 (got %d)", p->n_channels);
    OFP_REQUIRE(p->block_size >= 1, "block_size must be positive");
    ofp_detector *d = new ofp_detector;
    d->p = *p;
    d->n_streams = n_streams;
    d->n_lanes = n_streams * p->n_channels;
    d->buf = nullptr;
    cudaError_t e = cudaMalloc(&d->buf, sizeof(float) * 11 * d->n_lanes);
    if (e != cudaSuccess) {
        set_error("cudaMalloc detector state: %s", cudaGetErrorString(e));
        delete d;
        return e == cudaErrorMemoryAllocation ? OFP_ENOMEM : OFP_ECUDA;
    }
    *out = d;
    return ofp_detector_reset(d, nullptr);
}

int ofp_detector_destroy(ofp_detector *det) {
    if (!det) return OFP_OK;
    cudaFree(det->buf);
    delete det;
    return OFP_OK;
}

int ofp_detector_reset(ofp_detector *det, void *stream) {
    OFP_REQUIRE(det, "null detector");
    const int64_t n = det->n_lanes;  // all allocated lanes, also when fewer streams are active
    k1_reset<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        state_of(det), n, det->p.floor_db);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

int ofp_detector_get_state(ofp_detector *det, int field, void *buf_dev, void *stream) {
    OFP_REQUIRE(det && buf_dev && field >= 0 && field <= 10, "bad argument");
    OFP_CUDA_CHECK(cudaMemcpyAsync(buf_dev, det->buf + static_cast<int64_t>(field) * det->n_lanes,
                                   sizeof(float) * det->n_lanes, cudaMemcpyDeviceToDevice,
                                   static_cast<cudaStream_t>(stream)));
    return OFP_OK;
}

int ofp_detector_set_state(ofp_detector *det, int field, const void *buf_dev, void *stream) {
    OFP_REQUIRE(det && buf_dev && field >= 0 && field <= 10, "bad argument");
    OFP_CUDA_CHECK(cudaMemcpyAsync(det->buf + static_cast<int64_t>(field) * det->n_lanes, buf_dev,
                                   sizeof(float) * det->n_lanes, cudaMemcpyDeviceToDevice,
                                   static_cast<cudaStream_t>(stream)));
    return OFP_OK;
}

int ofp_detect_offline(ofp_detector *det, const float *x_dev, int64_t n_samples, int64_t rec_stride,
                       int64_t warm_n, float *rel_dev, int64_t rel_stride, int32_t *on_channel_dev,
                       int32_t *on_sample_dev, int32_t *on_count_dev, int32_t cap, void *stream) {
    OFP_REQUIRE(det && x_dev, "null argument");
    OFP_REQUIRE(n_samples >= 0 && warm_n >= 0 && cap >= 0, "negative size");
    OFP_REQUIRE(cap == 0 || (on_channel_dev && on_sample_dev), "onset buffers missing");
    const int64_t nb = n_samples / det->p.block_size;
    OFP_REQUIRE(nb * det->p.block_size < (1ll << 31), "recording too long for int32 sample indices");
    return launch_k1(det, x_dev, n_samples, rec_stride, warm_n, nb * det->p.block_size, rel_dev, rel_stride,
                     on_channel_dev, on_sample_dev, on_count_dev, cap, static_cast<cudaStream_t>(stream));
}

int ofp_detect_continue(ofp_detector *det, const float *x_dev, int64_t n_samples, int64_t rec_stride,
                        int64_t first_block, float *rel_dev, int64_t rel_stride, int32_t *on_channel_dev,
                        int32_t *on_sample_dev, int32_t *on_count_dev, int32_t cap, void *stream) {
    OFP_REQUIRE(det && x_dev && on_channel_dev && on_sample_dev && on_count_dev, "null argument");
    OFP_REQUIRE(n_samples >= 0 && first_block >= 0 && cap >= 0, "negative size");
    OFP_REQUIRE(n_samples % det->p.block_size == 0, "a continuation must be whole blocks");
    OFP_REQUIRE((first_block + n_samples / det->p.block_size) * det->p.block_size < (1ll << 31),
                "recording too long for int32 sample indices");
    if (n_samples == 0) return OFP_OK;
    return launch_k1(det, x_dev, n_samples, rec_stride, 0, n_samples, rel_dev, rel_stride, on_channel_dev,
                     on_sample_dev, on_count_dev, cap, static_cast<cudaStream_t>(stream), first_block, true);
}

int ofp_detect_block(ofp_detector *det, const float *x_dev, int64_t stream_stride, float *rel_dev, int32_t *ch_dev,
                     int32_t *delta_dev, int32_t *count_dev, void *stream) {
    OFP_REQUIRE(det && x_dev && ch_dev && delta_dev && count_dev, "null argument");
    const int64_t bc = static_cast<int64_t>(det->p.block_size) * det->p.n_channels;
    if (det->n_streams == 1) stream_stride = bc;  // a size-1 leading dimension carries no meaningful stride
    OFP_REQUIRE(stream_stride >= bc, "stream_stride smaller than one block");
    return launch_k1(det, x_dev, det->p.block_size, stream_stride, 0, det->p.block_size, rel_dev, bc, ch_dev, delta_dev,
                     count_dev, det->p.n_channels, static_cast<cudaStream_t>(stream));
}

int ofp_detect_warmup(ofp_detector *det, const float *x_dev, int64_t n_samples, int64_t rec_stride, void *stream) {
    OFP_REQUIRE(det && x_dev, "null argument");
    return launch_k1(det, x_dev, n_samples, rec_stride, n_samples, 0, nullptr, 0, nullptr, nullptr, nullptr, 0,
                     static_cast<cudaStream_t>(stream));
}

int ofp_detect_offline_host(const ofp_detector_params *p, const float *x_host, int64_t n_rec, int64_t n_samples,
                            int64_t warm_n, float *rel_host, int32_t *on_channel_host, int32_t *on_sample_host,
                            int32_t *on_count_host, int32_t cap) {
    OFP_REQUIRE(p && x_host && on_channel_host && on_sample_host && on_count_host, "null argument");
    OFP_REQUIRE(n_rec > 0 && n_samples >= 0 && warm_n >= 0, "bad size");
    // The batch is fed in TIME segments of all recordings at once (2-D copies out of the [R, N, C] host
    // array), double buffered: the copy of segment s+1 overlaps the kernel of segment s, which continues
    // from the detector state segment s-1 left behind (ofp_detect_continue).  Every launch therefore keeps
    // all recordings' lanes busy and the wall time is the host->device transfer plus one segment's kernel.
    // Batches too large for one detector go through the same pipeline in chunks of recordings.
    const int64_t C = p->n_channels, B = p->block_size;
    const int64_t n_main = (n_samples / B) * B;
    const int64_t rel_row = n_main * C;
    int64_t seg = std::max<int64_t>(env_int("OFP_HOST_SEGMENT", 49152), std::min(warm_n, n_samples));
    seg = (seg + B - 1) / B * B;
    if (seg % 4) seg *= 4;  // 16-byte rows for the TMA path
    const int64_t chunk = std::min<int64_t>(n_rec, env_int("OFP_HOST_CHUNK", 8192));
    const int64_t seg_alloc = std::min(seg + B, std::max<int64_t>(n_samples, 1));  // + an absorbed partial block
    // Staging buffers, streams and events are kept between calls (cudaMalloc / cudaFree of multi-GB buffers
    // cost more than the whole pipeline); ofp_host_release() frees them.  One call at a time.
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    HostCache &cx = g_host_cache[current_device()];
    ofp_detector *det = nullptr;
    auto cleanup = [&]() { ofp_detector_destroy(det); };
#define HOST_CHECK(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            cleanup();                                                                    \
            return OFP_ECUDA;                                                             \
        }                                                                                 \
    } while (0)
    const bool trace = env_int("OFP_HOST_TRACE", 0) != 0;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    int rc = ofp_detector_create(&det, chunk, p);
    if (rc != OFP_OK) { cleanup(); return rc; }
    if (!cx.copy) HOST_CHECK(cudaStreamCreateWithFlags(&cx.copy, cudaStreamNonBlocking));
    if (!cx.comp) HOST_CHECK(cudaStreamCreateWithFlags(&cx.comp, cudaStreamNonBlocking));
    auto ensure = [&](void **ptr, size_t &cap, size_t need) -> cudaError_t {
        if (need <= cap) return cudaSuccess;
        cudaFree(*ptr); *ptr = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(ptr, need);
        if (e == cudaSuccess) cap = need;
        return e;
    };
    const size_t seg_bytes = sizeof(float) * static_cast<size_t>(std::max<int64_t>(chunk * ((seg_alloc + B) * C + 4), 4));
    for (int i = 0; i < 2; ++i) {
        HOST_CHECK(ensure(reinterpret_cast<void **>(&cx.x[i]), cx.x_cap[i], seg_bytes));
        if (rel_host) HOST_CHECK(ensure(reinterpret_cast<void **>(&cx.rel[i]), cx.rel_cap[i], seg_bytes));
        if (!cx.copied[i]) HOST_CHECK(cudaEventCreateWithFlags(&cx.copied[i], cudaEventDisableTiming));
        if (!cx.done[i]) HOST_CHECK(cudaEventCreateWithFlags(&cx.done[i], cudaEventDisableTiming));
    }
    const size_t on_bytes = sizeof(int32_t) * static_cast<size_t>(std::max<int64_t>(chunk * cap, 1));
    HOST_CHECK(ensure(reinterpret_cast<void **>(&cx.och), cx.och_cap, on_bytes));
    HOST_CHECK(ensure(reinterpret_cast<void **>(&cx.oix), cx.oix_cap, on_bytes));
    HOST_CHECK(ensure(reinterpret_cast<void **>(&cx.ocn), cx.ocn_cap, sizeof(int32_t) * static_cast<size_t>(chunk)));
    HOST_CHECK(cudaDeviceSynchronize());  // the detector's first reset ran on the default stream
    const size_t host_pitch = sizeof(float) * n_samples * C;
    const double t_setup = now();
    int64_t use = 0;  // buffer uses so far (parity = slot)
    for (int64_t r0 = 0; r0 < n_rec; r0 += chunk) {
        const int64_t n = std::min(chunk, n_rec - r0);
        det->n_streams = n;  // the state arrays are sized for `chunk` recordings
        rc = ofp_detector_reset(det, cx.comp);
        if (rc != OFP_OK) { cleanup(); return rc; }
        if (n_samples == 0) HOST_CHECK(cudaMemsetAsync(cx.ocn, 0, sizeof(int32_t) * n, cx.comp));
        for (int64_t t0 = 0; t0 < n_samples; ++use) {
            const int slot = static_cast<int>(use & 1);
            // the last segment carries the trailing partial block (only the warm-up's high-pass reads it)
            int64_t len = std::min(seg, n_samples - t0);
            if (t0 + len < n_samples && n_samples - (t0 + len) < B) len = n_samples - t0;
            if (use >= 2) HOST_CHECK(cudaStreamWaitEvent(cx.copy, cx.done[slot], 0));
            const int64_t pitch = (len * C + 3) / 4 * 4;  // device row stride in floats (16-byte rows: TMA path)
            HOST_CHECK(cudaMemcpy2DAsync(cx.x[slot], sizeof(float) * pitch, x_host + r0 * n_samples * C + t0 * C,
                                         host_pitch, sizeof(float) * len * C, n, cudaMemcpyHostToDevice, cx.copy));
            HOST_CHECK(cudaEventRecord(cx.copied[slot], cx.copy));
            HOST_CHECK(cudaStreamWaitEvent(cx.comp, cx.copied[slot], 0));
            const int64_t blocks = len / B;
            if (t0 == 0)
                rc = ofp_detect_offline(det, cx.x[slot], len, pitch, warm_n, cx.rel[slot], blocks * B * C, cx.och,
                                        cx.oix, cx.ocn, cap, cx.comp);
            else
                rc = ofp_detect_continue(det, cx.x[slot], blocks * B, pitch, t0 / B, cx.rel[slot], blocks * B * C,
                                         cx.och, cx.oix, cx.ocn, cap, cx.comp);
            if (rc != OFP_OK) { cleanup(); return rc; }
            if (rel_host && blocks > 0)
                HOST_CHECK(cudaMemcpy2DAsync(rel_host + r0 * rel_row + t0 * C, sizeof(float) * rel_row, cx.rel[slot],
                                             sizeof(float) * blocks * B * C, sizeof(float) * blocks * B * C, n,
                                             cudaMemcpyDeviceToHost, cx.comp));
            HOST_CHECK(cudaEventRecord(cx.done[slot], cx.comp));
            t0 += len;
        }
        HOST_CHECK(cudaMemcpyAsync(on_channel_host + r0 * cap, cx.och, sizeof(int32_t) * n * cap, cudaMemcpyDeviceToHost, cx.comp));
        HOST_CHECK(cudaMemcpyAsync(on_sample_host + r0 * cap, cx.oix, sizeof(int32_t) * n * cap, cudaMemcpyDeviceToHost, cx.comp));
        HOST_CHECK(cudaMemcpyAsync(on_count_host + r0, cx.ocn, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, cx.comp));
        HOST_CHECK(cudaStreamSynchronize(cx.comp));  // och / oix / ocn are reused by the next chunk
    }
    HOST_CHECK(cudaStreamSynchronize(cx.copy));
#undef HOST_CHECK
    const double t_done = now();
    cleanup();
    if (trace)
        fprintf(stderr, "[ofp_detect_offline_host] setup %.1f ms, pipeline %.1f ms, cleanup %.1f ms (segment %lld samples)\n",
                t_setup - t_start, t_done - t_setup, now() - t_done, static_cast<long long>(seg));
    return OFP_OK;
}

int ofp_backtrack_onsets(const float *rel_dev, int64_t rec_stride, int64_t n_rows, int32_t n_channels,
                         int32_t block_size, int32_t buffer_size, float alpha, float tol, int32_t streaming,
                         const int32_t *on_channel_dev, int32_t *on_sample_dev, const int32_t *on_count_dev,
                         int32_t n_rec, int32_t cap, void *stream) {
    OFP_REQUIRE(rel_dev && on_channel_dev && on_sample_dev && on_count_dev, "null argument");
    OFP_REQUIRE(buffer_size >= block_size, "backtrack_buffer_size should be at least block_size");
    if (n_rec == 0 || cap == 0) return OFP_OK;
    const float omba = static_cast<float>(1.0 - static_cast<double>(alpha));
    const int64_t total = static_cast<int64_t>(n_rec) * cap;
    k_backtrack<<<static_cast<unsigned>((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        rel_dev, rec_stride, n_rows, n_channels, block_size, buffer_size, alpha, omba, tol, streaming, on_channel_dev,
        on_sample_dev, on_count_dev, n_rec, cap);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

int ofp_ar_envelope(const float *x_dev, float *y_dev, float attack, float release, int size, int num_samples,
                    void *stream) {
    OFP_REQUIRE(x_dev && y_dev && size > 0 && num_samples > 0, "bad argument");
    k_ar_envelope<<<(size + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(x_dev, y_dev, attack, release,
                                                                                      size, num_samples);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

int ofp_minmax_envelope(const float *x_dev, float *min_dev, float *max_dev, float alpha_min, float alpha_max,
                        float minmin, int n_samples, int n_channels, void *stream) {
    OFP_REQUIRE(x_dev && min_dev && max_dev && n_samples >= 0 && n_channels > 0, "bad argument");
    const float ia_min = static_cast<float>(1.0 - static_cast<double>(alpha_min));
    const float ia_max = static_cast<float>(1.0 - static_cast<double>(alpha_max));
    k_minmax_envelope<<<(n_channels + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        x_dev, min_dev, max_dev, alpha_min, alpha_max, ia_min, ia_max, minmin, n_samples, n_channels);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

}  // extern "C"
