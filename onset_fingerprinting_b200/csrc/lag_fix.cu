// K3 (group_onsets) and K4 (lag_fix): onset groups -> bounded-lag cross-correlation -> adjusted onsets.
//
// Replaces find_onset_groups (reference detection.py:131-189), fix_onsets (373-451),
// cross_correlation_lag (195-268) and adjust_onset (299-352).
//
// K4 design (DESIGN.md "K4"): one CTA per hit.  The audio section [a-look, b+look) x C is read from
// HBM once (coalesced, straight from the [R, N, C] recording -- no materialised sections), median
// filtered / differenced / rectified in shared memory, then the later channels are aligned one after
// the other (the reference moves the reference onset between pairs, SURVEY Q6, so pairs are
// sequential).  Per pair only the 2*tol lags of the legal window are evaluated (the reference
// computes all 2n-1 and slices): one thread per 4 consecutive lags, double accumulation in index
// order (bit-identical to the oracle's restatement), sliding register window so that each step is
// 2 LDS.64 + 4 DFMA.  argmax / max / weighted sums are block reductions.
#include "ofp_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace ofp {

constexpr int LAG_NONE = INT32_MIN;
constexpr int FIX_OK = 0, FIX_DEGENERATE = 1, FIX_REF_CRASH = 2, FIX_TOO_LONG = 3, FIX_INCOMPLETE = 4;
#ifndef OFP_K4_THREADS
#define OFP_K4_THREADS 192
#endif
#ifndef OFP_K4_MINCTA
#define OFP_K4_MINCTA 3
#endif
// CTA size: the kernels are compiled for three sizes (lag_fix_body.cuh is included once per size, K4_THREADS
// a compile-time constant in each) and the launch picks one: 192 threads for many-channel sections, fewer
// for 3-channel ones whose 60-lag windows over ~200 samples cannot feed more.
constexpr int K4_MAX_THREADS = OFP_K4_THREADS;
// Lags per thread.  ODD on purpose: neighbouring lanes then read x at a stride of LPT doubles, and an
// odd stride spreads the 32 lanes of an LDS.64 over all bank pairs (LPT = 4 gave 8-way conflicts).
constexpr int LPT = 3;
constexpr int CC_UN = 4;  // time steps per unrolled iteration

struct FixParams {
    int32_t filter_size, d, direction, take_abs, zero_left, cutoff, tol, shift;
    int32_t to_end;  // section runs to the end of the recording (streaming ring, multilateration.py:457-466)
};

struct K4Args {
    const float *audio;
    int64_t rec_stride, n_samples;
    int32_t C, H, Lmax;
    const int32_t *hit_rec;  // [H] recording of each hit (may be null: rec = hit)
    const int32_t *onsets;   // [H, C]
    FixParams fp;
    int32_t *out_onsets;     // [H, C]
    int32_t *out_lags;       // [H, C] (may be null)
    int32_t *out_status;     // [H]
    int32_t screen;          // float32 screening of the lag window (OFP_K4_SCREEN, default on)
    int32_t columns;         // column mode: shared memory holds two channel columns instead of the [L, C] section
};
// ---- float32 screening of the lag window ---------------------------------------------------------
// The result of cross_correlation_lag is an argmax, so the exact (double, index-order) sums are only
// needed for lags that can still win.  Pass 1 evaluates every lag of the window with float32 FMAs
// (LPF lags per thread, the x window sliding through registers; threads split (lag group, time
// segment)), pass 2 bounds each value rigorously:
//     |v32[w] - V[w]| <= delta[w] = (n + 8) * 2^-24 * ||x|| * ||y|| / cnt[w] + 2^-21 * |v32[w]|
// (V = the oracle's float32(double sum) / cnt; float32 recursive summation error <= n u sum|x_i y_i|,
// Cauchy-Schwarz; the last term covers the roundings of both quotients) and keeps the lags with
// v32 + delta >= max_w (v32 - delta).  The true first maximum is always among them.  One survivor
// (the usual case) IS the answer; a handful are recomputed exactly, one thread each, and compared
// with np.argmax's tie rule; more than CAND_CAP, or non-finite data, falls back to cc_argmax.
#ifndef OFP_K4_LPF
#define OFP_K4_LPF 9
#endif
#ifndef OFP_K4_CCF_UN
#define OFP_K4_CCF_UN 8
#endif
constexpr int LPF = OFP_K4_LPF;        // lags per thread (odd: conflict-free x reads across lanes)
constexpr int CCF_UN = OFP_K4_CCF_UN;  // time steps per unrolled iteration
constexpr int XPAD = 16;     // zeros on both sides of the float copy of x
constexpr int CAND_CAP = 32;

// screening statistics since the last reset: pairs screened, decided by one survivor, decided by an exact
// recomputation of a few survivors, handed to the exact path
__device__ unsigned long long g_cc_stats[4];

struct CcScratch {
    float *xf;     // [L + 2 * XPAD], xf[XPAD + i] = x[i]
    float *yf;     // [L]
    float *part;   // [K4_THREADS * LPF] partial sums, then v32[w]
    int *cand;     // [CAND_CAP + 1]: count, then window indices
    double *red_d; // block reduction scratch
    float *red_f;
};

// Stand-alone twins of cross_correlation_lag (detection.py:195-268) and adjust_onset (299-352) for
// P independent signal pairs of equal length n: x, y [P, n] float32.
struct PairArgs {
    const float *x, *y;
    int32_t P, n, d, take_abs, use_legal, cutoff, tol;
    const int32_t *onsets;  // [P, 2] (onset of x, onset of y) or legal lags (l0, l1) when use_legal
    const int32_t *new_lag; // [P] (adjust only)
    int32_t *out;           // cc: [P] lag; adjust: [P, 2]
    int32_t screen;
};

#ifndef OFP_K4_SMALL_LPF
#define OFP_K4_SMALL_LPF 9
#endif
// lags per thread in the float32 screening (tunable per CTA size; measured 1 / 3 / 5 / 9 on 3-channel hits:
// 8.13-8.18 ms for 380 k hits, no difference -- those hits are barrier / latency bound, scripts/ncu_lines.py)
namespace t32 { constexpr int K4_THREADS = 32, K4_MINCTA = 32, K4_LPF = OFP_K4_SMALL_LPF;
#include "lag_fix_body.cuh"
}
#ifndef OFP_K4_T64_MINCTA
#define OFP_K4_T64_MINCTA 16
#endif
namespace t64 { constexpr int K4_THREADS = 64, K4_MINCTA = OFP_K4_T64_MINCTA, K4_LPF = OFP_K4_SMALL_LPF;
#include "lag_fix_body.cuh"
}
#ifndef OFP_K4_T128_MINCTA
#define OFP_K4_T128_MINCTA 8
#endif
namespace t128 { constexpr int K4_THREADS = 128, K4_MINCTA = OFP_K4_T128_MINCTA, K4_LPF = OFP_K4_SMALL_LPF;
#include "lag_fix_body.cuh"
}
namespace t192 { constexpr int K4_THREADS = K4_MAX_THREADS, K4_MINCTA = OFP_K4_MINCTA, K4_LPF = LPF;
#include "lag_fix_body.cuh"
}

// ---------------------------------------------------------------------------------------------
// K3: find_onset_groups (detection.py:131-189), one thread per recording (sequential scan).
// ---------------------------------------------------------------------------------------------
struct K3Args {
    const int32_t *on_ch, *on_idx, *on_cnt;  // [R, cap], [R, cap], [R]
    int32_t R, cap, C, max_distance, min_channels, close_channel, max_groups;
    int32_t *groups;   // [R, max_groups, C], -1 = missing
    int32_t *n_groups; // [R]
};

__global__ void k3_group(const K3Args a) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const int n = min(a.on_cnt[r], a.cap);
    const int32_t *ch = a.on_ch + static_cast<int64_t>(r) * a.cap;
    const int32_t *ix = a.on_idx + static_cast<int64_t>(r) * a.cap;
    int32_t *out = a.groups + static_cast<int64_t>(r) * a.max_groups * a.C;
    int32_t cur[32];
    unsigned mask = 0;
    int first = 0, ng = 0, members = 0;
    auto flush = [&]() {
        if (members > 0 && __popc(mask) >= a.min_channels) {
            bool keep = true;
            if (a.close_channel >= 0) {  // detection.py:184-185: all(x[close] <= x), -1 entries included
                const int32_t cv = (mask >> a.close_channel) & 1u ? cur[a.close_channel] : -1;
                for (int c = 0; c < a.C; ++c) {
                    const int32_t v = (mask >> c) & 1u ? cur[c] : -1;
                    if (!(cv <= v)) keep = false;
                }
            }
            if (keep) {
                if (ng < a.max_groups)
                    for (int c = 0; c < a.C; ++c) out[ng * a.C + c] = (mask >> c) & 1u ? cur[c] : -1;
                ++ng;
            }
        }
    };
    for (int i = 0; i < n; ++i) {
        const int s = ix[i], c = ch[i];
        if (members > 0 && abs(s - first) > a.max_distance) { flush(); mask = 0; members = 0; }
        if (members == 0) first = s;
        cur[c] = s;  // a later onset of the same channel overwrites (detection.py:171-172)
        mask |= 1u << c;
        ++members;
    }
    flush();
    a.n_groups[r] = ng;
}

// Hit list from the per-recording groups: offsets = exclusive prefix sum of min(n_groups, max_groups).
__global__ void k3_compact(const int32_t *groups, const int32_t *n_groups, const int64_t *offsets, int R,
                           int max_groups, int C, int32_t *hit_rec, int32_t *hit_onsets) {
    const int r = blockIdx.x;
    if (r >= R) return;
    const int ng = min(n_groups[r], max_groups);
    const int64_t off = offsets[r];
    for (int e = threadIdx.x; e < ng * C; e += blockDim.x)
        hit_onsets[off * C + e] = groups[static_cast<int64_t>(r) * max_groups * C + e];
    for (int g = threadIdx.x; g < ng; g += blockDim.x) hit_rec[off + g] = r;
}

}  // namespace ofp

using namespace ofp;

extern "C" {

int ofp_cc_screen_stats(uint64_t *stats4_host, int32_t reset) {
    OFP_REQUIRE(stats4_host, "null argument");
    OFP_CUDA_CHECK(cudaMemcpyFromSymbol(stats4_host, g_cc_stats, 4 * sizeof(uint64_t)));
    if (reset) {
        const uint64_t z[4] = {0, 0, 0, 0};
        OFP_CUDA_CHECK(cudaMemcpyToSymbol(g_cc_stats, z, sizeof z));
    }
    return OFP_OK;
}

static int fix_smem_bytes(int32_t n_channels, int32_t max_section, int threads, bool columns = false) {
    const size_t cc = (static_cast<size_t>(max_section) + 2 * XPAD + max_section + 16 + threads * LPF) * sizeof(float) +
                      (CAND_CAP + 1) * sizeof(int);
    return static_cast<int>(2 * (max_section + 16) * sizeof(double) +
                            static_cast<size_t>(max_section | 1) * (columns ? 2 : n_channels) * sizeof(float) + cc);
}
// upper bound over the CTA sizes a launch may pick
int ofp_fix_onsets_smem_bytes(int32_t n_channels, int32_t max_section) {
    return fix_smem_bytes(n_channels, max_section, K4_MAX_THREADS);
}

int ofp_fix_onsets(const float *audio_dev, int64_t n_samples, int64_t rec_stride, int32_t n_channels,
                   const int32_t *hit_rec_dev, const int32_t *onsets_dev, int32_t n_hits, int32_t filter_size,
                   int32_t d, int32_t direction, int32_t take_abs, int32_t zero_left, int32_t cutoff, int32_t tol,
                   int32_t shift, int32_t max_section, int32_t *out_onsets_dev, int32_t *out_lags_dev,
                   int32_t *out_status_dev, void *stream) {
    return ofp_fix_onsets_ex(audio_dev, n_samples, rec_stride, n_channels, hit_rec_dev, onsets_dev, n_hits,
                             filter_size, d, direction, take_abs, zero_left, cutoff, tol, shift, max_section, 0,
                             out_onsets_dev, out_lags_dev, out_status_dev, stream);
}

int ofp_fix_onsets_ex(const float *audio_dev, int64_t n_samples, int64_t rec_stride, int32_t n_channels,
                      const int32_t *hit_rec_dev, const int32_t *onsets_dev, int32_t n_hits, int32_t filter_size,
                      int32_t d, int32_t direction, int32_t take_abs, int32_t zero_left, int32_t cutoff,
                      int32_t tol, int32_t shift, int32_t max_section, int32_t flags, int32_t *out_onsets_dev,
                      int32_t *out_lags_dev, int32_t *out_status_dev, void *stream) {
    if (n_hits == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(audio_dev && onsets_dev && out_onsets_dev && out_status_dev, "null argument");
    OFP_REQUIRE(n_channels >= 2 && n_channels <= 32, "n_channels must be in 2..32");
    OFP_REQUIRE(filter_size >= 1 && filter_size <= 15, "filter_size must be in 1..15");
    OFP_REQUIRE(d >= 0 && d <= 3 && cutoff >= 0 && tol >= 0 && direction >= 0 && direction <= 2,
                "bad option (difference order d must be 0..3)");
    K4Args a;
    a.audio = audio_dev; a.rec_stride = rec_stride; a.n_samples = n_samples; a.C = n_channels; a.H = n_hits;
    a.Lmax = max_section; a.hit_rec = hit_rec_dev; a.onsets = onsets_dev;
    a.fp = FixParams{filter_size, d, direction, take_abs, zero_left, cutoff, tol, shift, flags & 1};
    a.out_onsets = out_onsets_dev; a.out_lags = out_lags_dev; a.out_status = out_status_dev;
    { const char *e = getenv("OFP_K4_SCREEN"); a.screen = e ? atoi(e) : 1; }
    // CTA size by the work a hit carries: 16-channel sections feed 192 threads; a 3-channel hit (60-lag windows
    // over ~300 samples) is barrier / latency bound and runs best as ONE WARP per hit at 32 CTAs per SM
    // (380 k hits: 8.0 ms with 128 threads and the worst-case section size, 5.7 / 4.2 / 3.1 ms with 128 / 64 / 32
    // threads once the section is sized by the data).  The float32 screening needs 2 tol lags <= 9 per thread.
    // OFP_K4_SMALL forces the size used below the 192-thread threshold (A/B runs).
    static const int small = getenv("OFP_K4_SMALL") ? atoi(getenv("OFP_K4_SMALL")) : 0;
    const int64_t work = static_cast<int64_t>(n_channels) * max_section;
    // Column mode: shared memory holds two channel columns instead of the [L, C] section; the later channel of
    // each pair is median filtered from the L1-resident section when its pair comes up.  It lifts the 3-hits-per-SM
    // cap of 16-channel sections (49 KB each) but measured no faster (2.03e6 vs 2.09e6 hits/s: the per-pair column
    // reads are uncoalesced and 16-channel hits are not occupancy bound), so it is only taken when the full
    // section does not fit shared memory at all (e.g. 32 channels x 1500 samples).  OFP_K4_COLUMNS=0/1 and
    // OFP_K4_COL_THREADS override (A/B).
    const char *ec = getenv("OFP_K4_COLUMNS");
    a.columns = ec ? atoi(ec) : (fix_smem_bytes(n_channels, max_section, K4_MAX_THREADS) > 200 * 1024 ? 1 : 0);
    int threads = work >= 4096 ? K4_MAX_THREADS : (work > 1536 ? 64 : 32);
    if (a.columns && work >= 4096) threads = getenv("OFP_K4_COL_THREADS") ? atoi(getenv("OFP_K4_COL_THREADS")) : 128;
    if (small > 0 && threads != K4_MAX_THREADS) threads = small <= 32 ? 32 : (small <= 64 ? 64 : 128);
    while (threads < 128 && 2 * tol > threads * LPF) threads *= 2;
    if (threads > 128) threads = K4_MAX_THREADS;
    auto kern = threads == K4_MAX_THREADS ? t192::k4_fix
                                          : (threads == 32 ? t32::k4_fix : (threads == 64 ? t64::k4_fix : t128::k4_fix));
    const int smem = fix_smem_bytes(n_channels, max_section, threads, a.columns != 0);  // partial sums follow the CTA size
    OFP_REQUIRE(smem <= 220 * 1024, "max_section %d x %d channels needs %d bytes of shared memory", max_section,
                n_channels, smem);
    OFP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<n_hits, threads, smem, static_cast<cudaStream_t>(stream)>>>(a);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

static int launch_pairs(bool adjust, const float *x, const float *y, int32_t P, int32_t n, int32_t d, int32_t take_abs,
                        int32_t use_legal, int32_t cutoff, int32_t tol, const int32_t *onsets, const int32_t *new_lag,
                        int32_t *out, void *stream) {
    OFP_REQUIRE(x && y && onsets && out, "null argument");
    OFP_REQUIRE(n >= 1 && n <= 8 * K4_MAX_THREADS, "signal length must be in 1..%d", 8 * K4_MAX_THREADS);
    OFP_REQUIRE(d >= 0 && d < n, "bad difference order");
    if (P == 0) return OFP_OK;
    const char *e = getenv("OFP_K4_SCREEN");
    PairArgs a{x, y, P, n, d, take_abs, use_legal, cutoff, tol, onsets, new_lag, out, e ? atoi(e) : 1};
    const int smem = 2 * (n + 16) * sizeof(double) + 2 * n * sizeof(float) +
                     (static_cast<size_t>(n) + 2 * XPAD + n + 16 + K4_MAX_THREADS * LPF) * sizeof(float) +
                     (CAND_CAP + 1) * sizeof(int);
    // n above ~1300 needs more than the default 48 KB of dynamic shared memory: opt in (as k4_fix does)
    OFP_REQUIRE(smem <= 220 * 1024, "signal length %d needs %d bytes of shared memory", n, smem);
    if (adjust) {
        OFP_CUDA_CHECK(cudaFuncSetAttribute(t192::k4_adjust_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        t192::k4_adjust_pairs<<<P, K4_MAX_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(a);
    } else {
        OFP_CUDA_CHECK(cudaFuncSetAttribute(t192::k4_cc_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        t192::k4_cc_pairs<<<P, K4_MAX_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(a);
    }
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

int ofp_cross_correlation_lag(const float *x_dev, const float *y_dev, int32_t n_pairs, int32_t n, int32_t d,
                              int32_t take_abs, int32_t use_legal_lags, int32_t cutoff, int32_t tol,
                              const int32_t *onsets_or_legal_dev, int32_t *lag_dev, void *stream) {
    return launch_pairs(false, x_dev, y_dev, n_pairs, n, d, take_abs, use_legal_lags, cutoff, tol,
                        onsets_or_legal_dev, nullptr, lag_dev, stream);
}

int ofp_adjust_onset(const float *x_dev, const float *y_dev, int32_t n_pairs, int32_t n, const int32_t *onsets_dev,
                     const int32_t *new_lag_dev, int32_t *out_dev, void *stream) {
    OFP_REQUIRE(new_lag_dev, "null argument");
    return launch_pairs(true, x_dev, y_dev, n_pairs, n, 0, 0, 0, 0, 0, onsets_dev, new_lag_dev, out_dev, stream);
}

int ofp_group_onsets(const int32_t *on_channel_dev, const int32_t *on_sample_dev, const int32_t *on_count_dev,
                     int32_t n_rec, int32_t cap, int32_t n_channels, int32_t max_distance, int32_t min_channels,
                     int32_t close_channel, int32_t max_groups, int32_t *groups_dev, int32_t *n_groups_dev,
                     void *stream) {
    OFP_REQUIRE(on_channel_dev && on_sample_dev && on_count_dev && groups_dev && n_groups_dev, "null argument");
    OFP_REQUIRE(n_channels >= 1 && n_channels <= 32, "n_channels must be in 1..32");
    K3Args a{on_channel_dev, on_sample_dev, on_count_dev, n_rec, cap, n_channels, max_distance, min_channels,
             close_channel, max_groups, groups_dev, n_groups_dev};
    k3_group<<<(n_rec + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

int ofp_compact_groups(const int32_t *groups_dev, const int32_t *n_groups_dev, const int64_t *offsets_dev,
                       int32_t n_rec, int32_t max_groups, int32_t n_channels, int32_t *hit_rec_dev,
                       int32_t *hit_onsets_dev, void *stream) {
    OFP_REQUIRE(groups_dev && n_groups_dev && offsets_dev && hit_rec_dev && hit_onsets_dev, "null argument");
    k3_compact<<<n_rec, 64, 0, static_cast<cudaStream_t>(stream)>>>(groups_dev, n_groups_dev, offsets_dev, n_rec,
                                                                    max_groups, n_channels, hit_rec_dev,
                                                                    hit_onsets_dev);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

}  // extern "C"
