// K2: spectral-flux onset features (sm_100a).
//
// Replaces the per-hop numpy FFT of the reference's realtime analysis -- RecAnalysis.fft /
// onset_strength (realtime/recording.py:273-311: hann(2048) * mean_c(audio[-2048:]) -> rfft ->
// |.|^2 -> 10 log10(max(1e-10, .)) -> clamp at max-80 -> mean(max(0, s - s_prev))) -- and the STFT
// front half of detect_onsets_spectral (detection.py:96-110: |stft(n_fft=256, hop=32)| x per-bin
// weight -> mean(max(0, dD))).  librosa and loopmate are absent from the reference tree, so this
// row is a restatement of their documented behaviour ("parity unpinned" for K2).
//
// One CTA walks F consecutive frames of one recording: the frame is gathered (channel mean on the
// fly from the interleaved [R, N, C] audio; the 16x overlap between frames is served by L1/L2),
// windowed, transformed by an in-shared-memory radix-2 Stockham FFT of n_fft/2 complex points plus
// the real-input split, and reduced to one flux value against the previous frame's spectrum, which
// never leaves shared memory.  FP32 throughout (the reference uses numpy's float32 FFT).
#include "ofp_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace ofp {

constexpr int K2_THREADS = 256;

struct K2Args {
    const float *x;
    int64_t rec_stride, n_samples;
    int32_t C, n_fft, hop, n_frames, frames_per_cta;
    int32_t center;      // 0: frame j = x[(j+1)*hop - n_fft : (j+1)*hop] (realtime); 1: centred on j*hop (librosa)
    int32_t reflect;     // padding of a centred transform: 0 zeros, 1 reflect
    int32_t mode;        // 0: log-power flux, 1: weighted magnitude flux
    float top_db;        // mode 0: clamp at (frame max - top_db); <= 0 disables
    const float *window; // [n_fft]
    const float *weight; // [n_fft/2 + 1] or null (mode 1)
    float *flux;         // [R, n_frames]
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ float k2_block_sum(float v, float *scratch) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < K2_THREADS / 32; ++i) r += scratch[i];
    return r;
}
__device__ __forceinline__ float k2_block_max(float v, float *scratch) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = scratch[0];
    for (int i = 1; i < K2_THREADS / 32; ++i) r = fmaxf(r, scratch[i]);
    return r;
}

// Twiddle exp(-2 pi i m / N) for m < N from the half-circle table tw[k], k < N/2.
__device__ __forceinline__ float2 tw_full(const float2 *tw, int m, int H) {
    const float2 w = tw[m & (H - 1)];
    return m >= H ? make_float2(-w.x, -w.y) : w;
}

__global__ void __launch_bounds__(K2_THREADS) k2_flux(const K2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.n_fft, H = N / 2, tid = threadIdx.x;
    float *win = reinterpret_cast<float *>(smem_raw);            // [N]
    float2 *tw = reinterpret_cast<float2 *>(win + N);            // [H]: exp(-2 pi i k / N)
    float2 *bufA = tw + H;                                        // [H]
    float2 *bufB = bufA + H;                                      // [H]
    float *prevS = reinterpret_cast<float *>(bufB + H);           // [H + 1]
    float *curS = prevS + H + 1;                                  // [H + 1]
    float *mono = curS + H + 1;                                   // [frames_per_cta * hop + N]: channel mean
    __shared__ float red[K2_THREADS / 32];

    const int r = blockIdx.y;
    const float *xr = a.x + static_cast<int64_t>(r) * a.rec_stride;
    for (int i = tid; i < N; i += K2_THREADS) win[i] = a.window[i];
    for (int k = tid; k < H; k += K2_THREADS) {
        float s, c;
        sincospif(-2.0f * static_cast<float>(k) / static_cast<float>(N), &s, &c);
        tw[k] = make_float2(c, s);
    }
    const float invC = 1.0f / static_cast<float>(a.C);
    const int j0 = blockIdx.x * a.frames_per_cta;
    const int j1 = min(j0 + a.frames_per_cta, a.n_frames);
    // ---- channel mean of every sample this CTA's frames touch, once (the frames overlap N/hop-fold) ----
    // frame j starts at fs(j); the buffer covers frames j0-1 (seed of the previous spectrum) .. j1-1
    const int64_t fs0 = a.center ? static_cast<int64_t>(j0 - 1) * a.hop - H : static_cast<int64_t>(j0) * a.hop - N;
    const int span = (j1 - j0) * a.hop + N;
    for (int i = tid; i < span; i += K2_THREADS) {
        int64_t t = fs0 + i;
        if (a.center && a.reflect) {  // numpy pad mode 'reflect' (no edge repeat)
            if (t < 0) t = -t;
            if (t >= a.n_samples) t = 2 * (a.n_samples - 1) - t;
        }
        float sv = 0.f;
        if (t >= 0 && t < a.n_samples) {
            const float *p = xr + t * a.C;
            for (int c = 0; c < a.C; ++c) sv += p[c];
            sv = a.C > 1 ? sv * invC : sv;
        }
        mono[i] = sv;
    }
    __syncthreads();

    for (int j = j0 - 1; j < j1; ++j) {  // frame j0-1 only seeds prevS
        if (j < 0) {  // before the recording: the spectrum of silence
            const float s0 = a.mode == 0 ? 10.0f * log10f(1e-10f) : 0.0f;
            for (int k = tid; k <= H; k += K2_THREADS) prevS[k] = s0;
            __syncthreads();
            continue;
        }
        // ---- window: z[n] = x[2n] + i x[2n+1] ----
        const float *fr = mono + static_cast<int64_t>(j - (j0 - 1)) * a.hop;
        for (int n = tid; n < H; n += K2_THREADS)
            bufA[n] = make_float2(fr[2 * n] * win[2 * n], fr[2 * n + 1] * win[2 * n + 1]);
        __syncthreads();
        // ---- Stockham autosort FFT of H complex points: radix-4 passes, one radix-2 pass if log2 H is odd ----
        float2 *src = bufA, *dst = bufB;
        int L = 1;
        for (; L * 4 <= H; L *= 4) {
            const int q = H / 4, tstep = N / (4 * L);  // W_{4L}^{k} = exp(-2 pi i k tstep / N)
            for (int i = tid; i < q; i += K2_THREADS) {
                const int k = i & (L - 1);
                const float2 v0 = src[i];
                float2 v1 = src[i + q], v2 = src[i + 2 * q], v3 = src[i + 3 * q];
                if (L > 1) {
                    v1 = cmul(v1, tw_full(tw, k * tstep, H));
                    v2 = cmul(v2, tw_full(tw, 2 * k * tstep, H));
                    v3 = cmul(v3, tw_full(tw, 3 * k * tstep, H));
                }
                const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
                const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
                const int o = ((i - k) << 2) + k;
                dst[o] = make_float2(s02.x + s13.x, s02.y + s13.y);
                dst[o + L] = make_float2(d02.x + d13.y, d02.y - d13.x);      // d02 - i d13
                dst[o + 2 * L] = make_float2(s02.x - s13.x, s02.y - s13.y);
                dst[o + 3 * L] = make_float2(d02.x - d13.y, d02.y + d13.x);  // d02 + i d13
            }
            __syncthreads();
            float2 *tmp = src; src = dst; dst = tmp;
        }
        if (L < H) {  // remaining radix-2 pass (L == H / 2)
            for (int i = tid; i < H / 2; i += K2_THREADS) {
                const float2 u = src[i], t = cmul(src[i + H / 2], tw_full(tw, i * (N / (2 * L)), H));
                dst[i] = make_float2(u.x + t.x, u.y + t.y);
                dst[i + L] = make_float2(u.x - t.x, u.y - t.y);
            }
            __syncthreads();
            float2 *tmp = src; src = dst; dst = tmp;
        }
        // src now holds Z[k], k < H.  Real-input split: X[k], k = 0..H
        float pmax = -INFINITY;
        for (int k = tid; k <= H; k += K2_THREADS) {
            const float2 zk = src[k & (H - 1)], zc = src[(H - k) & (H - 1)];
            const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));   // (Z[k] + conj Z[H-k]) / 2
            const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));  // (Z[k] - conj Z[H-k]) / 2i
            const float2 w = k < H ? tw[k] : make_float2(-1.f, 0.f);
            const float2 ow = cmul(o, w);
            const float re = e.x + ow.x, im = e.y + ow.y;
            const float p = re * re + im * im;
            float sv;
            if (a.mode == 0) sv = 10.0f * log10f(fmaxf(1e-10f, p));
            else sv = sqrtf(p) * (a.weight ? a.weight[k] : 1.0f);
            curS[k] = sv;
            pmax = fmaxf(pmax, sv);
        }
        float lo = -INFINITY;
        if (a.mode == 0 && a.top_db > 0.f) lo = k2_block_max(pmax, red) - a.top_db;
        __syncthreads();
        float acc = 0.f;
        for (int k = tid; k <= H; k += K2_THREADS) {
            const float s = fmaxf(curS[k], lo), sp = fmaxf(prevS[k], lo);
            acc += fmaxf(0.f, s - sp);
        }
        const float total = k2_block_sum(acc, red);
        if (tid == 0 && j >= j0) a.flux[static_cast<int64_t>(r) * a.n_frames + j] = total / static_cast<float>(H + 1);
        __syncthreads();
        float *t2 = prevS; prevS = curS; curS = t2;
    }
}

// Complex short-time spectra of already cut frames -- data.stft / stft_frame (data.py:581-654): frame f of
// `frame_length` samples is centred in n_fft points (librosa.util.pad_center), multiplied by the window and
// transformed; numpy computes `np.fft.rfft(window * x)` in double (the window is float64) and the caller stores
// complex64, so the transform runs in double here too (radix-2 Stockham over n_fft complex points in shared memory,
// one CTA per frame -- a handful of 512-point frames per onset, nowhere near a hot loop) and rounds once at the end.
__global__ void __launch_bounds__(128) k2_stft_frames(const float *frames, int64_t n_frames, int frame_length, int n_fft,
                                                      const double *window, float2 *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *bufA = reinterpret_cast<double2 *>(smem_raw), *bufB = bufA + n_fft;
    const int tid = threadIdx.x, N = n_fft, lpad = (n_fft - frame_length) / 2;
    const int64_t f = blockIdx.x;
    const float *x = frames + f * frame_length;
    for (int n = tid; n < N; n += 128) {
        const int i = n - lpad;
        const double v = (i >= 0 && i < frame_length) ? static_cast<double>(x[i]) * window[n] : 0.0;
        bufA[n] = make_double2(v, 0.0);
    }
    __syncthreads();
    double2 *src = bufA, *dst = bufB;
    for (int L = 1; L < N; L *= 2) {  // Stockham autosort, radix 2: after the pass sub-transforms have length 2 L
        for (int i = tid; i < N / 2; i += 128) {
            const int k = i & (L - 1);
            double sn, cs;
            sincospi(-static_cast<double>(k) / static_cast<double>(L), &sn, &cs);
            const double2 u = src[i], v = src[i + N / 2];
            const double2 t = make_double2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
            const int o = ((i - k) << 1) + k;
            dst[o] = make_double2(u.x + t.x, u.y + t.y);
            dst[o + L] = make_double2(u.x - t.x, u.y - t.y);
        }
        __syncthreads();
        double2 *tmp = src; src = dst; dst = tmp;
    }
    float2 *o = out + f * (N / 2 + 1);
    for (int k = tid; k <= N / 2; k += 128)
        o[k] = make_float2(static_cast<float>(src[k].x), static_cast<float>(src[k].y));
}

}  // namespace ofp
#include "spectral_flux_warp.cuh"
namespace ofp {

// librosa.util.peak_pick restated (greedy, sequential; one thread per recording): x[n] is a peak iff
//   x[n] == max(x[n-pre_max : n+post_max])  and  x[n] >= mean(x[n-pre_avg : n+post_avg]) + delta  and  n - last > wait
// with HALF-OPEN windows clipped at the ends (librosa's documented conditions; its 0.9 implementation builds them
// from maximum_filter1d / uniform_filter1d of length pre+post with a shifted origin -- oracle/librosa_standin.py).
// A zero-valued sample is never a peak (librosa multiplies the detections into x).  The mean is accumulated in double.
__global__ void k2_peak_pick(const float *oe, int n_frames, int R, int pre_max, int post_max, int pre_avg,
                             int post_avg, float delta, int wait, int32_t *peaks, int32_t *n_peaks, int cap) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float *x = oe + static_cast<int64_t>(r) * n_frames;
    int cnt = 0;
    int last = -(1 << 30);
    for (int n = 0; n < n_frames; ++n) {
        if (n - last <= wait) continue;
        const int a0 = max(0, n - pre_max), a1 = min(n_frames, max(n + post_max, n + 1));
        float mx = -INFINITY;
        for (int i = a0; i < a1; ++i) mx = fmaxf(mx, x[i]);
        if (x[n] != mx || x[n] == 0.0f) continue;
        const int b0 = max(0, n - pre_avg), b1 = min(n_frames, max(n + post_avg, n + 1));
        double s = 0.0;
        for (int i = b0; i < b1; ++i) s += static_cast<double>(x[i]);
        const float avg = static_cast<float>(s / static_cast<double>(b1 - b0));
        if (x[n] < __fadd_rn(avg, delta)) continue;
        if (cnt < cap) peaks[static_cast<int64_t>(r) * cap + cnt] = n;
        ++cnt;
        last = n;
    }
    n_peaks[r] = cnt;
}

}  // namespace ofp

using namespace ofp;

extern "C" {

int ofp_spectral_flux(const float *x_dev, int64_t n_rec, int64_t n_samples, int64_t rec_stride, int32_t n_channels,
                      int32_t n_fft, int32_t hop, int32_t center, int32_t reflect, int32_t mode, float top_db,
                      const float *window_dev, const float *weight_dev, int32_t n_frames, float *flux_dev,
                      void *stream) {
    if (n_frames <= 0 || n_rec == 0) return OFP_OK;  // nothing to compute: the buffers may be null
    OFP_REQUIRE(x_dev && window_dev && flux_dev, "null argument");
    OFP_REQUIRE(n_fft >= 64 && n_fft <= 8192 && (n_fft & (n_fft - 1)) == 0, "n_fft must be a power of two in 64..8192");
    OFP_REQUIRE(hop >= 1 && n_channels >= 1 && n_rec <= 65535, "bad argument");
    K2Args a;
    a.x = x_dev; a.rec_stride = rec_stride; a.n_samples = n_samples; a.C = n_channels; a.n_fft = n_fft; a.hop = hop;
    a.n_frames = n_frames; a.center = center; a.reflect = reflect; a.mode = mode;
    a.top_db = top_db; a.window = window_dev; a.weight = weight_dev; a.flux = flux_dev;
    const int H = n_fft / 2;
    // 2048-point fast path (the realtime / config-5 shape): one warp per frame, transform in registers
    const size_t smem_w = sizeof(K2WSmem) + sizeof(float) * (static_cast<size_t>(K2W_WARPS * K2W_FRAMES) * hop + n_fft);
    const bool force_generic = getenv("OFP_K2_GENERIC") != nullptr;  // read per call: tests compare both kernels
    if (n_fft == 2 * K2W_H && hop % 2 == 0 && smem_w <= (227 * 1024) / OFP_K2W_MINCTA && !force_generic) {
        auto kern = mode == 0 ? (top_db > 0.f ? k2_flux_warp<0, true> : k2_flux_warp<0, false>) : k2_flux_warp<1, false>;
        OFP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_w)));
        const int fpc = K2W_WARPS * K2W_FRAMES;
        dim3 grid((n_frames + fpc - 1) / fpc, static_cast<unsigned>(n_rec));
        kern<<<grid, K2W_WARPS * 32, smem_w, static_cast<cudaStream_t>(stream)>>>(a);
        OFP_CUDA_CHECK(cudaGetLastError());
        return OFP_OK;
    }
    // frames per CTA: enough that the shared channel-mean buffer is read ~1.5x instead of n_fft/hop-fold,
    // few enough that 3 CTAs fit an SM and the grid keeps every SM busy
    int fpc = std::max(8, std::min(64, (3 * n_fft / 2 + hop - 1) / hop));
    const size_t fixed = sizeof(float) * n_fft + sizeof(float2) * 3 * H + sizeof(float) * 2 * (H + 1) + 16;
    while (fpc > 1 && fixed + sizeof(float) * (static_cast<size_t>(fpc) * hop + n_fft) > 220 * 1024) fpc /= 2;
    a.frames_per_cta = fpc;
    const size_t smem = fixed + sizeof(float) * (static_cast<size_t>(fpc) * hop + n_fft);
    OFP_REQUIRE(smem <= 227 * 1024, "n_fft %d with hop %d needs %zu bytes of shared memory", n_fft, hop, smem);
    OFP_CUDA_CHECK(cudaFuncSetAttribute(k2_flux, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    dim3 grid((n_frames + a.frames_per_cta - 1) / a.frames_per_cta, static_cast<unsigned>(n_rec));
    k2_flux<<<grid, K2_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(a);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

int ofp_stft_frames(const float *frames_dev, int64_t n_frames, int32_t frame_length, int32_t n_fft,
                    const double *window_dev, float *out_dev, void *stream) {
    if (n_frames == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(frames_dev && window_dev && out_dev, "null argument");
    OFP_REQUIRE(n_fft >= 2 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0, "n_fft must be a power of two in 2..4096");
    OFP_REQUIRE(frame_length >= 1 && frame_length <= n_fft, "frame_length must be in 1..n_fft");
    OFP_REQUIRE(n_frames < (1ll << 31), "too many frames for one launch");
    const size_t smem = 2 * static_cast<size_t>(n_fft) * sizeof(double2);
    OFP_CUDA_CHECK(cudaFuncSetAttribute(k2_stft_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    k2_stft_frames<<<static_cast<unsigned>(n_frames), 128, smem, static_cast<cudaStream_t>(stream)>>>(
        frames_dev, n_frames, frame_length, n_fft, window_dev, reinterpret_cast<float2 *>(out_dev));
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

int ofp_peak_pick(const float *oe_dev, int32_t n_rec, int32_t n_frames, int32_t pre_max, int32_t post_max,
                  int32_t pre_avg, int32_t post_avg, float delta, int32_t wait, int32_t *peaks_dev,
                  int32_t *n_peaks_dev, int32_t cap, void *stream) {
    if (n_rec == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(oe_dev && peaks_dev && n_peaks_dev, "null argument");
    k2_peak_pick<<<(n_rec + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(
        oe_dev, n_frames, n_rec, pre_max, post_max, pre_avg, post_avg, delta, wait, peaks_dev, n_peaks_dev, cap);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

}  // extern "C"
