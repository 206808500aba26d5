// Small device twins of the helper functions around the detector / lag path (sm_100a):
//   ofp_lfilter              ButterworthFilter.__call__ = scipy.signal.lfilter(b, a, x, axis=0, zi)  (detection.py:487-501)
//   ofp_filter_data          filter_data                                                        (detection.py:355-370)
//   ofp_detect_onset_region  detect_onset_region                                                (detection.py:454-484)
//   ofp_correlate_full       np.correlate(a, b, "full") as used by find_lag / find_lag_multi     (multilateration.py:878-899)
// None of them is on the throughput-critical path (they are per-call helpers of a few hundred
// samples); they exist so that the whole detection.py / multilateration.py surface runs on the device.
#include "ofp_common.cuh"

namespace ofp {

constexpr int LF_MAX_ORDER = 8;

struct LfArgs {
    float b[LF_MAX_ORDER + 1], a[LF_MAX_ORDER + 1];
    int32_t order, n, C;
    const float *x;
    float *y, *zi;
};

// scipy's float_filt (direct form II transposed), float32, evaluated left to right without
// contraction: y = z0 + b0*x; z[i] = (z[i+1] + x*b[i+1]) - y*a[i+1]; z[last] = x*b[last] - y*a[last].
__global__ void k_lfilter(const LfArgs q) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= q.C) return;
    float z[LF_MAX_ORDER];
#pragma unroll
    for (int i = 0; i < LF_MAX_ORDER; ++i) z[i] = i < q.order ? q.zi[i * q.C + c] : 0.f;
    for (int t = 0; t < q.n; ++t) {
        const float x = q.x[static_cast<int64_t>(t) * q.C + c];
        const float y = q.order > 0 ? __fadd_rn(z[0], __fmul_rn(q.b[0], x)) : __fmul_rn(q.b[0], x);
#pragma unroll
        for (int i = 0; i < LF_MAX_ORDER; ++i) {
            if (i < q.order - 1)
                z[i] = __fsub_rn(__fadd_rn(z[i + 1], __fmul_rn(x, q.b[i + 1])), __fmul_rn(y, q.a[i + 1]));
            else if (i == q.order - 1)
                z[i] = __fsub_rn(__fmul_rn(x, q.b[i + 1]), __fmul_rn(y, q.a[i + 1]));
        }
        q.y[static_cast<int64_t>(t) * q.C + c] = y;
    }
#pragma unroll
    for (int i = 0; i < LF_MAX_ORDER; ++i)
        if (i < q.order) q.zi[i * q.C + c] = z[i];
}

// filter_data: diff = x[t] - x[t-1] (0 for the first row); "up" zeroes samples with diff < 0,
// "down" those with diff > 0.  Out of place (the reference masks with the diff of the ORIGINAL x).
__global__ void k_filter_data(const float *x, float *out, int64_t n, int C, int direction) {
    const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (e >= n * C) return;
    const float v = x[e];
    const float diff = e >= C ? __fsub_rn(v, x[e - C]) : 0.f;
    const bool kill = direction == 1 ? diff < 0.f : diff > 0.f;
    out[e] = kill ? 0.f : v;
}

// detect_onset_region, one warp-sized block per signal.  audio [P, len] float32.
//   region = audio[start:end], start = max(onset - n/2, 0), end = min(onset + n/2, len)
//   filtered = scipy.signal.medfilt(|region|, ks) (zero padded)
//   binary = filtered > factor * max(filtered)        (float32, numpy >= 2 scalar rules)
//   binary_opening(binary, ones(5)) = erosion then dilation with border value 0
//   -> start + index of the first True (0 if none)
__global__ void k_onset_region(const float *audio, int64_t len, const int32_t *onsets, int32_t n, int32_t ks,
                               float factor, int32_t *out) {
    extern __shared__ float sm[];
    const int p = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int64_t on = onsets[p];
    int64_t s = on - n / 2, e = on + n / 2;
    // python slicing of audio[start:end] with start clamped at 0 and end at len
    if (s < 0) s = 0;
    if (e > len) e = len;
    if (e < 0) { e += len; if (e < 0) e = 0; }
    if (s > len) s = len;
    const int m = e > s ? static_cast<int>(e - s) : 0;
    float *a = sm, *f = sm + m;
    unsigned char *b0 = reinterpret_cast<unsigned char *>(f + m), *b1 = b0 + m;
    const float *src = audio + static_cast<int64_t>(p) * len + s;
    for (int i = tid; i < m; i += nt) a[i] = fabsf(src[i]);
    __syncthreads();
    const int half = ks / 2;
    float mx = -INFINITY;
    for (int i = tid; i < m; i += nt) {
        // median of ks values (zeros outside) by rank counting
        float med = 0.f;
        for (int u = 0; u < ks; ++u) {
            const int iu = i - half + u;
            const float vu = (iu >= 0 && iu < m) ? a[iu] : 0.f;
            int rank = 0;
            for (int w = 0; w < ks; ++w) {
                const int iw = i - half + w;
                const float vw = (iw >= 0 && iw < m) ? a[iw] : 0.f;
                rank += (vw < vu) || (vw == vu && w < u);
            }
            if (rank == half) med = vu;
        }
        f[i] = med;
        mx = fmaxf(mx, med);
    }
    __shared__ float red[32];
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    float gm = -INFINITY;
    for (int w = 0; w < (nt + 31) / 32; ++w) gm = fmaxf(gm, red[w]);
    const float thr = __fmul_rn(factor, gm);
    for (int i = tid; i < m; i += nt) b0[i] = f[i] > thr;
    __syncthreads();
    for (int i = tid; i < m; i += nt) {  // erosion, structure ones(5), outside = 0
        bool all = true;
        for (int k = -2; k <= 2; ++k) all = all && (i + k >= 0 && i + k < m && b0[i + k]);
        b1[i] = all;
    }
    __syncthreads();
    int first = INT32_MAX;
    for (int i = tid; i < m; i += nt) {  // dilation
        bool any = false;
        for (int k = -2; k <= 2; ++k) any = any || (i + k >= 0 && i + k < m && b1[i + k]);
        if (any) first = min(first, i);
    }
    __shared__ int redi[32];
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_down_sync(0xffffffffu, first, o));
    if ((tid & 31) == 0) redi[tid >> 5] = first;
    __syncthreads();
    if (tid == 0) {
        int g = INT32_MAX;
        for (int w = 0; w < (nt + 31) / 32; ++w) g = min(g, redi[w]);
        out[p] = static_cast<int32_t>(s + (g == INT32_MAX ? 0 : g));
    }
}

// np.correlate(x, y, "full")[k] = sum_i x[i + k - (n-1)] * y[i], double accumulation in index order,
// rounded once to float32.  x, y [P, n]; out [P, 2n-1].
__global__ void k_correlate_full(const float *x, const float *y, int n, float *out) {
    extern __shared__ float sm[];
    float *xs = sm, *ys = sm + n;
    const int p = blockIdx.x;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        xs[i] = x[static_cast<int64_t>(p) * n + i];
        ys[i] = y[static_cast<int64_t>(p) * n + i];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * n - 1; k += blockDim.x) {
        const int m = k - (n - 1);
        const int i0 = m < 0 ? -m : 0, i1 = m > 0 ? n - m : n;
        double acc = 0.0;
        for (int i = i0; i < i1; ++i) acc = __fma_rn(static_cast<double>(xs[i + m]), static_cast<double>(ys[i]), acc);
        out[static_cast<int64_t>(p) * (2 * n - 1) + k] = __double2float_rn(acc);
    }
}

}  // namespace ofp

using namespace ofp;

extern "C" {

int ofp_lfilter(const float *b_host, const float *a_host, int32_t order, const float *x_dev, float *y_dev,
                float *zi_dev, int32_t n_samples, int32_t n_channels, void *stream) {
    OFP_REQUIRE(b_host && a_host && x_dev && y_dev && zi_dev, "null argument");
    OFP_REQUIRE(order >= 1 && order <= LF_MAX_ORDER, "filter order must be in 1..%d", LF_MAX_ORDER);
    OFP_REQUIRE(n_samples >= 0 && n_channels >= 1, "bad size");
    OFP_REQUIRE(a_host[0] == 1.0f, "a[0] must be 1 (scipy.signal.butter output)");
    if (n_samples == 0) return OFP_OK;
    LfArgs q;
    for (int i = 0; i <= LF_MAX_ORDER; ++i) {
        q.b[i] = i <= order ? b_host[i] : 0.f;
        q.a[i] = i <= order ? a_host[i] : 0.f;
    }
    q.order = order; q.n = n_samples; q.C = n_channels; q.x = x_dev; q.y = y_dev; q.zi = zi_dev;
    k_lfilter<<<(n_channels + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(q);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

int ofp_filter_data(const float *x_dev, float *out_dev, int64_t n_samples, int32_t n_channels, int32_t direction,
                    void *stream) {
    OFP_REQUIRE(x_dev && out_dev && x_dev != out_dev, "null or aliased argument");
    OFP_REQUIRE(direction == 1 || direction == 2, "direction must be 1 (\"up\") or 2 (\"down\")");
    const int64_t total = n_samples * n_channels;
    if (total <= 0) return OFP_OK;
    k_filter_data<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x_dev, out_dev, n_samples, n_channels, direction);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

int ofp_detect_onset_region(const float *audio_dev, int32_t n_signals, int64_t len, const int32_t *onsets_dev,
                            int32_t n, int32_t median_filter_size, float threshold_factor, int32_t *out_dev,
                            void *stream) {
    if (n_signals == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(audio_dev && onsets_dev && out_dev, "null argument");
    OFP_REQUIRE(n >= 0 && n <= 8192, "n must be in 0..8192");
    OFP_REQUIRE(median_filter_size >= 1 && median_filter_size % 2 == 1 && median_filter_size <= 31,
                "kernel_size must be odd and <= 31");
    const int m = n + 2;
    const size_t smem = static_cast<size_t>(m) * (2 * sizeof(float) + 2);
    k_onset_region<<<n_signals, 128, smem, static_cast<cudaStream_t>(stream)>>>(
        audio_dev, len, onsets_dev, n, median_filter_size, threshold_factor, out_dev);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

int ofp_correlate_full(const float *x_dev, const float *y_dev, int32_t n_pairs, int32_t n, float *out_dev,
                       void *stream) {
    if (n_pairs == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(x_dev && y_dev && out_dev, "null argument");
    OFP_REQUIRE(n >= 1 && n <= 24 * 1024, "signal length must be in 1..24576");
    const size_t smem = 2 * static_cast<size_t>(n) * sizeof(float);
    OFP_CUDA_CHECK(cudaFuncSetAttribute(k_correlate_full, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    k_correlate_full<<<n_pairs, 256, smem, static_cast<cudaStream_t>(stream)>>>(x_dev, y_dev, n, out_dev);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Onset-window extraction (data.py:55-192 FrameExtractor / FastFrameExtractor): the gather
// sliding_window_view(audio, F, axis=0)[start] -> [C, F] per hit, fused after K4 so that the
// [hits, C, F] training windows are cut on the device from the resident recordings.
// One CTA per hit: the F x C block is contiguous in the time-major recording (coalesced read),
// transposed through shared memory, written as C contiguous rows of F samples.
// ---------------------------------------------------------------------------------------------
namespace ofp {

struct FrameArgs {
    const float *audio;
    int64_t n_samples, rec_stride;
    int32_t C, H, F, pre, use_min;
    const int32_t *hit_rec, *onsets, *shifts;  // [H] or null, [H, C], [H] or null
    float *out;                                // [H, C, F]
    int32_t *status;                           // [H]: 0 ok, 1 index out of range (numpy raises IndexError)
};

__global__ void k_extract_frames(const FrameArgs a) {
    extern __shared__ float sm[];
    const int h = blockIdx.x, C = a.C, F = a.F;
    const int64_t rec = a.hit_rec ? a.hit_rec[h] : 0;
    const float *src = a.audio + rec * a.rec_stride;
    const int64_t nwin = a.n_samples - F + 1;  // rows of the sliding-window view
    const int32_t off = a.pre - (a.shifts ? a.shifts[h] : 0);
    __shared__ int64_t start[32];
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    if (threadIdx.x < C) {
        int64_t o;
        if (a.use_min) {
            o = a.onsets[static_cast<int64_t>(h) * C];
            for (int c = 1; c < C; ++c) o = min(o, static_cast<int64_t>(a.onsets[static_cast<int64_t>(h) * C + c]));
        } else {
            o = a.onsets[static_cast<int64_t>(h) * C + threadIdx.x];
        }
        int64_t s = o - off;
        if (s < 0) s += nwin;  // numpy's negative indexing into the view
        if (s < 0 || s >= nwin) { s = 0; bad = 1; }
        start[threadIdx.x] = s;
    }
    __syncthreads();
    float *dst = a.out + static_cast<int64_t>(h) * C * F;
    if (bad) {
        for (int e = threadIdx.x; e < C * F; e += blockDim.x) dst[e] = 0.f;
        if (threadIdx.x == 0) a.status[h] = 1;
        return;
    }
    if (a.use_min) {
        const float *p = src + start[0] * C;
        for (int e = threadIdx.x; e < C * F; e += blockDim.x) {  // e = t * C + c
            const int t = e / C, c = e - t * C;
            sm[c * (F + 1) + t] = p[e];
        }
    } else {
        for (int e = threadIdx.x; e < C * F; e += blockDim.x) {
            const int t = e / C, c = e - t * C;
            sm[c * (F + 1) + t] = src[(start[c] + t) * C + c];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < C * F; e += blockDim.x) {  // e = c * F + t
        const int c = e / F, t = e - c * F;
        dst[e] = sm[c * (F + 1) + t];
    }
    if (threadIdx.x == 0) a.status[h] = 0;
}

}  // namespace ofp

extern "C" int ofp_extract_frames(const float *audio_dev, int64_t n_samples, int64_t rec_stride, int32_t n_channels,
                                  const int32_t *hit_rec_dev, const int32_t *onsets_dev, const int32_t *shifts_dev,
                                  int32_t n_hits, int32_t frame_length, int32_t pre_samples, int32_t use_min_onset,
                                  float *frames_dev, int32_t *status_dev, void *stream) {
    if (n_hits == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(audio_dev && onsets_dev && frames_dev && status_dev, "null argument");
    OFP_REQUIRE(n_channels >= 1 && n_channels <= 32, "n_channels must be in 1..32");
    OFP_REQUIRE(frame_length >= 1 && frame_length <= n_samples, "frame_length must be in 1..n_samples");
    const size_t smem = static_cast<size_t>(n_channels) * (frame_length + 1) * sizeof(float);
    OFP_REQUIRE(smem <= 200 * 1024, "frame of %d x %d samples does not fit shared memory", frame_length, n_channels);
    ofp::FrameArgs a{audio_dev, n_samples, rec_stride, n_channels, n_hits, frame_length, pre_samples, use_min_onset,
                     hit_rec_dev, onsets_dev, shifts_dev, frames_dev, status_dev};
    OFP_CUDA_CHECK(cudaFuncSetAttribute(ofp::k_extract_frames, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    ofp::k_extract_frames<<<n_hits, 256, smem, static_cast<cudaStream_t>(stream)>>>(a);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

// ---------------------------------------------------------------------------------------------
// Peak refinement of the notebooks' dataset builder (notebooks/refresh.org:262-279):
//     onsets[h, c] + np.argmax(audio[onsets[h, c] : onsets[h, c] + tolerance, c])
// One warp per (hit, channel): strided window read, first maximum wins (np.argmax).
// ---------------------------------------------------------------------------------------------
namespace ofp {
__global__ void k_window_argmax(const float *audio, int64_t n_samples, int64_t rec_stride, int C, const int32_t *hit_rec,
                                const int32_t *onsets, int64_t n_pairs, int tol, int32_t *out) {
    const int64_t w = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_pairs) return;
    const int64_t h = w / C;
    const int c = static_cast<int>(w - h * C);
    const int64_t on = onsets[w];
    if (on < 0) { if (lane == 0) out[w] = static_cast<int32_t>(on); return; }  // missing channel stays missing
    const float *col = audio + (hit_rec ? hit_rec[h] : h) * rec_stride + c;
    const int64_t end = (on + tol < n_samples) ? on + tol : n_samples;  // python slice clips at the end of the recording
    float best = -INFINITY;
    int64_t best_i = INT64_MAX;
    for (int64_t i = on + lane; i < end; i += 32) {
        const float v = col[i * C];
        if (v > best || (best_i == INT64_MAX && !(v < best))) { best = v; best_i = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int64_t oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) out[w] = static_cast<int32_t>(best_i == INT64_MAX ? on : best_i);  // empty window: argmax raises; keep
}

// ---------------------------------------------------------------------------------------------
// RecAnalysis.tempogram (realtime/recording.py:313-327): for frame j the autocorrelation of the Hann-windowed
// last W onset-envelope values, irfft(|rfft(w * oe[j-W+1 .. j], n = 2W-1)|^2)[:W] (no circular wrap at that
// padding => the plain linear autocorrelation), normalised by (max + 1e-10).  Values before the first frame
// read 0 (the reference's ring is zero-initialised).  One CTA per (recording, selected frame).
// ---------------------------------------------------------------------------------------------
__global__ void k_tempogram(const float *oe, int64_t n_frames, const float *window, int W, int64_t first, int64_t every,
                            int64_t n_sel, float *tg) {
    extern __shared__ float tgs[];  // [W] windowed envelope, then [W] raw autocorrelation
    float *x = tgs, *ac = tgs + W;
    const int64_t r = blockIdx.y, j = first + static_cast<int64_t>(blockIdx.x) * every;
    const float *row = oe + r * n_frames;
    for (int t = threadIdx.x; t < W; t += blockDim.x) {
        const int64_t f = j - W + 1 + t;
        x[t] = __fmul_rn(window[t], f >= 0 && f < n_frames ? row[f] : 0.0f);
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int l = threadIdx.x; l < W; l += blockDim.x) {
        double acc = 0.0;
        for (int t = 0; t + l < W; ++t) acc = fma(static_cast<double>(x[t]), static_cast<double>(x[t + l]), acc);
        const float v = static_cast<float>(acc);
        ac[l] = v;
        mx = fmaxf(mx, v);
    }
    __shared__ float red[32];
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int k = 1; k < (blockDim.x + 31) / 32; ++k) mx = fmaxf(mx, red[k]);
    const float inv = __fadd_rn(mx, 1e-10f);
    float *dst = tg + (r * n_sel + blockIdx.x) * W;
    for (int l = threadIdx.x; l < W; l += blockDim.x) dst[l] = __fdiv_rn(ac[l], inv);
}
}  // namespace ofp

extern "C" int ofp_window_argmax(const float *audio_dev, int64_t n_samples, int64_t rec_stride, int32_t n_channels,
                                 const int32_t *hit_rec_dev, const int32_t *onsets_dev, int32_t n_hits,
                                 int32_t tolerance, int32_t *out_dev, void *stream) {
    if (n_hits == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(audio_dev && onsets_dev && out_dev, "null argument");
    OFP_REQUIRE(n_channels >= 1 && tolerance >= 1, "bad size");
    const int64_t pairs = static_cast<int64_t>(n_hits) * n_channels;
    ofp::k_window_argmax<<<static_cast<unsigned>((pairs * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        audio_dev, n_samples, rec_stride, n_channels, hit_rec_dev, onsets_dev, pairs, tolerance, out_dev);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

extern "C" int ofp_tempogram(const float *oe_dev, int32_t n_rec, int64_t n_frames, const float *window_dev,
                             int32_t win_length, int64_t first_frame, int64_t every, int64_t n_selected, float *tg_dev,
                             void *stream) {
    if (n_rec == 0 || n_selected == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(oe_dev && window_dev && tg_dev, "null argument");
    OFP_REQUIRE(win_length >= 1 && win_length <= 8192 && every >= 1 && first_frame >= 0, "bad tempogram shape");
    OFP_REQUIRE(n_selected >= 0 && (n_selected == 0 || first_frame + (n_selected - 1) * every < n_frames),
                "selected frames exceed the envelope");
    if (n_rec == 0 || n_selected == 0) return OFP_OK;
    OFP_REQUIRE(n_selected < (1ll << 31) && n_rec <= 65535, "grid too large: select fewer frames or recordings per call");
    const size_t smem = 2 * static_cast<size_t>(win_length) * sizeof(float);
    OFP_CUDA_CHECK(cudaFuncSetAttribute(ofp::k_tempogram, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    dim3 grid(static_cast<unsigned>(n_selected), static_cast<unsigned>(n_rec));
    ofp::k_tempogram<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(oe_dev, n_frames, window_dev, win_length,
                                                                           first_frame, every, n_selected, tg_dev);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}
