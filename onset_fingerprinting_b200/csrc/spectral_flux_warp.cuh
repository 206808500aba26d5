// K2, 2048-point fast path: one WARP per frame, the transform lives in registers.
//
// The 2048 real samples of a frame are packed into 1024 complex points z[n] = x[2n] + i x[2n+1] and
// transformed by the four-step factorisation 1024 = 32 x 32 with n = n1 + 32 n2, k = 32 k1 + k2:
//   step 1  lane n1 holds z[n1 + 32 n2], n2 = 0..31 (conflict-free 8-byte shared loads of the staged
//           channel mean times the window) and runs a 32-point DFT over n2 in registers;
//   step 2  Y[n1][k2] *= W_1024^(n1 k2)   (table T[k2][n1], conflict-free);
//   step 3  transpose through a padded warp-private shared tile (the only exchange, __syncwarp only),
//           lane k2 runs a 32-point DFT over n1 -> Z[32 k1 + k2] in register k1.
// The real-input split needs Z[k] and Z[1024 - k]: bin 32 k1 + k2 pairs with register 31 - k1 of lane
// 32 - k2, i.e. one warp shuffle per word (lane 0 pairs with its own registers).  Log-power / magnitude,
// the frame maximum, the rectified difference against the previous frame and the mean over the bins
// are register + shuffle work; the previous spectrum stays in registers because a warp walks
// K2W_FRAMES consecutive frames.  No block-wide barrier inside the frame loop.
#pragma once

namespace ofp {

#ifndef OFP_K2W_WARPS
#define OFP_K2W_WARPS 4
#endif
#ifndef OFP_K2W_FRAMES
#define OFP_K2W_FRAMES 16
#endif
#ifndef OFP_K2W_MINCTA
#define OFP_K2W_MINCTA 2
#endif
constexpr int K2W_WARPS = OFP_K2W_WARPS;     // warps per CTA
constexpr int K2W_FRAMES = OFP_K2W_FRAMES;   // consecutive frames per warp (+1 seed frame)
constexpr int K2W_H = 1024;      // complex points
constexpr int K2W_PAD = 33;      // row stride (float2) of the transpose tile

__host__ __device__ constexpr int brev5(int v) {
    return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

// d * W_32^m, m = 0..15 (compile-time m after unrolling: the branches fold away)
__device__ __forceinline__ float2 mul_w32(float2 d, int m) {
    constexpr float R = 0.70710678118654752f;
    // cos(2 pi m / 32), sin(2 pi m / 32)
    constexpr float CS[16] = {1.0f, 0.98078528040323044f, 0.92387953251128676f, 0.83146961230254524f,
                              0.70710678118654752f, 0.55557023301960222f, 0.38268343236508977f,
                              0.19509032201612827f, 0.0f, -0.19509032201612827f, -0.38268343236508977f,
                              -0.55557023301960222f, -0.70710678118654752f, -0.83146961230254524f,
                              -0.92387953251128676f, -0.98078528040323044f};
    constexpr float SN[16] = {0.0f, 0.19509032201612827f, 0.38268343236508977f, 0.55557023301960222f,
                              0.70710678118654752f, 0.83146961230254524f, 0.92387953251128676f,
                              0.98078528040323044f, 1.0f, 0.98078528040323044f, 0.92387953251128676f,
                              0.83146961230254524f, 0.70710678118654752f, 0.55557023301960222f,
                              0.38268343236508977f, 0.19509032201612827f};
    if (m == 0) return d;
    if (m == 8) return make_float2(d.y, -d.x);                       // -i
    if (m == 4) return make_float2((d.x + d.y) * R, (d.y - d.x) * R);   // (1 - i) / sqrt 2
    if (m == 12) return make_float2((d.y - d.x) * R, -(d.x + d.y) * R); // (-1 - i) / sqrt 2
    const float c = CS[m], s = SN[m];                                // W = c - i s
    return make_float2(fmaf(d.y, s, d.x * c), fmaf(-d.x, s, d.y * c));
}

// In-place radix-2 decimation-in-frequency DFT of 32 points held in registers.
// Result X[k] sits in v[brev5(k)].
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
#pragma unroll
    for (int len = 32; len >= 2; len >>= 1) {
        const int half = len >> 1, tstep = 32 / len;
#pragma unroll
        for (int b = 0; b < 32; b += len) {
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const float2 p = v[b + i], q = v[b + i + half];
                v[b + i] = make_float2(p.x + q.x, p.y + q.y);
                v[b + i + half] = mul_w32(make_float2(p.x - q.x, p.y - q.y), i * tstep);
            }
        }
    }
}

struct K2WSmem {
    float2 win2[K2W_H];                   // window pairs (w[2n], w[2n+1])
    float2 T[32 * 32];                    // T[k2][n1] = exp(-2 pi i n1 k2 / 1024)
    float2 T2[16 * 32];                   // T2[k1][l] = exp(-2 pi i (32 k1 + l) / 2048), k1 < 16 (lower half of the bins)
    float2 tile[K2W_WARPS][32 * K2W_PAD]; // warp-private transpose tiles
    // followed by the channel-mean buffer: float mono[K2W_WARPS * K2W_FRAMES * hop + 2048]
};

// Per-bin value from FOUR TIMES the power 4|X|^2 (the split below never forms X/2): mode 0 works on
// s' = 10 log10(max(4e-10, 4 p)) = s + 6.02 dB -- the constant cancels in the frame difference and in the top_db
// clamp; mode 1 on sqrt(4 p) * (weight / 2).
template <int MODE>
__device__ __forceinline__ float k2w_val(float p4, float half_wt) {
    if (MODE == 0) return 3.0102999566398120f * __log2f(fmaxf(4e-10f, p4));
    return sqrtf(p4) * half_wt;
}

// Channel mean of samples [fs0, fs0 + span) of one recording into shared memory (zeros outside the
// recording, numpy 'reflect' padding for a centred transform).  3-channel audio, the configs[1..4]
// shape, is read as three 16-byte loads per four samples with a batch of loads in flight per thread.
template <int NT>
__device__ __forceinline__ void k2_fill_mono(const K2Args &a, const float *__restrict__ xr, int64_t fs0, int span,
                                             float *mono, int tid) {
    const float invC = 1.0f / static_cast<float>(a.C);
    auto scalar = [&](int i) {
        int64_t t = fs0 + i;
        if (a.center && a.reflect) {
            if (t < 0) t = -t;
            if (t >= a.n_samples) t = 2 * (a.n_samples - 1) - t;
        }
        float sv = 0.f;
        if (t >= 0 && t < a.n_samples) {
            const float *p = xr + t * a.C;
            for (int c = 0; c < a.C; ++c) sv += __ldg(p + c);
            sv = a.C > 1 ? sv * invC : sv;
        }
        mono[i] = sv;
    };
    const bool vec3 = a.C == 3 && (reinterpret_cast<uintptr_t>(xr + fs0 * 3) & 15) == 0;
    if (!vec3) {
        for (int i = tid; i < span; i += NT) scalar(i);
        return;
    }
    constexpr int U = 4;
    const int groups = span >> 2;  // four samples = 12 floats = 3 float4
    for (int g0 = 0; g0 < groups; g0 += NT * U) {
        float4 q[U][3];
        bool in[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int g = g0 + u * NT + tid;
            const int64_t t = fs0 + 4 * static_cast<int64_t>(g);
            in[u] = g < groups && t >= 0 && t + 4 <= a.n_samples;
            if (in[u]) {
                const float4 *p = reinterpret_cast<const float4 *>(xr + t * 3);
                q[u][0] = __ldg(p); q[u][1] = __ldg(p + 1); q[u][2] = __ldg(p + 2);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int g = g0 + u * NT + tid;
            if (in[u]) {
                float4 m;  // same summation order as the scalar path: ((c0 + c1) + c2) * (1/3)
                m.x = ((q[u][0].x + q[u][0].y) + q[u][0].z) * invC;
                m.y = ((q[u][0].w + q[u][1].x) + q[u][1].y) * invC;
                m.z = ((q[u][1].z + q[u][1].w) + q[u][2].x) * invC;
                m.w = ((q[u][2].y + q[u][2].z) + q[u][2].w) * invC;
                *reinterpret_cast<float4 *>(mono + 4 * g) = m;
            } else if (g < groups) {
                for (int e = 0; e < 4; ++e) scalar(4 * g + e);
            }
        }
    }
    for (int i = (groups << 2) + tid; i < span; i += NT) scalar(i);
}

template <int MODE, bool TOPDB>
__global__ void __launch_bounds__(K2W_WARPS * 32, OFP_K2W_MINCTA) k2_flux_warp(const K2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K2WSmem &sm = *reinterpret_cast<K2WSmem *>(smem_raw);
    float *mono = reinterpret_cast<float *>(smem_raw + sizeof(K2WSmem));
    constexpr int N = 2 * K2W_H, H = K2W_H, NT = K2W_WARPS * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int r = blockIdx.y;
    const float *xr = a.x + static_cast<int64_t>(r) * a.rec_stride;
    for (int i = tid; i < H; i += NT) {
        sm.win2[i] = make_float2(a.window[2 * i], a.window[2 * i + 1]);
        const int k2 = i >> 5, n1 = i & 31;
        float s, c;
        sincospif(-2.0f * static_cast<float>((n1 * k2) & (H - 1)) / static_cast<float>(H), &s, &c);
        sm.T[i] = make_float2(c, s);
        if (i < 16 * 32) {
            sincospif(-2.0f * static_cast<float>(i) / static_cast<float>(N), &s, &c);
            sm.T2[i] = make_float2(c, s);
        }
    }
    const int j0 = blockIdx.x * (K2W_WARPS * K2W_FRAMES);
    const int j1 = min(j0 + K2W_WARPS * K2W_FRAMES, a.n_frames);
    // channel mean of every sample the CTA's frames touch (frames j0-1 .. j1-1), once
    const int64_t fs0 = a.center ? static_cast<int64_t>(j0 - 1) * a.hop - H : static_cast<int64_t>(j0) * a.hop - N;
    const int span = (j1 - j0) * a.hop + N;
    k2_fill_mono<NT>(a, xr, fs0, span, mono, tid);
    __syncthreads();

    const int jw0 = j0 + warp * K2W_FRAMES;
    const int jw1 = min(jw0 + K2W_FRAMES, j1);
    if (jw0 >= jw1) return;
    float2 *tile = sm.tile[warp];
    const int partner = (32 - lane) & 31;
    const bool l0 = lane == 0;

    float prevS[33], curS[33];
    for (int j = jw0 - 1; j < jw1; ++j) {
        if (j < 0) {
            const float s0 = MODE == 0 ? 3.0102999566398120f * __log2f(4e-10f) : 0.0f;
#pragma unroll
            for (int k = 0; k < 33; ++k) prevS[k] = s0;
            continue;
        }
        float2 v[32];
        {   // ---- step 1: windowed load, DFT over n2 ----
            const float2 *fr2 = reinterpret_cast<const float2 *>(mono + static_cast<int64_t>(j - (j0 - 1)) * a.hop);
#pragma unroll
            for (int n2 = 0; n2 < 32; ++n2) {
                const float2 x2 = fr2[lane + 32 * n2], w2 = sm.win2[lane + 32 * n2];
                v[n2] = make_float2(x2.x * w2.x, x2.y * w2.y);
            }
            fft32(v);
            // ---- step 2 + transpose: tile[k2][n1] = Y[n1][k2] W^(n1 k2) ----
#pragma unroll
            for (int k2 = 0; k2 < 32; ++k2) {
                float2 y = v[brev5(k2)];
                if (k2 > 0) y = cmul(y, sm.T[k2 * 32 + lane]);
                tile[k2 * K2W_PAD + lane] = y;
            }
            __syncwarp();
            // ---- step 3: lane = k2, DFT over n1 ----
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) v[n1] = tile[lane * K2W_PAD + n1];
            __syncwarp();
            fft32(v);
        }
        // ---- real-input split, a PAIR of bins per step; Z[32 k1 + lane] = v[brev5(k1)] ----
        // With E2 = Z[k] + conj Z[H-k], O2 = (Z[k] - conj Z[H-k]) / i and T = W^k O2:
        //   2 X[k] = E2 + T,   2 X[H-k] = conj(E2 - T)
        // so one evaluation of (E2, T) gives bins k and H - k.  A lane owns k = 32 k1 + lane for k1 < 16 (its
        // lower 16 registers) together with their mirrors H - k, whose Z values are the upper registers of lane
        // 32 - lane: one shuffle per word of the upper half only.  Lane 0 mirrors onto its own registers
        // (k1 = 0 pairs bin 0 with the Nyquist bin) and also owns the self-paired bin H/2.
        float pmax = -INFINITY;
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            const float2 za = v[brev5(k1)], up = v[brev5(31 - k1)];
            float2 pa;
            pa.x = __shfl_sync(0xffffffffu, up.x, partner);
            pa.y = __shfl_sync(0xffffffffu, up.y, partner);
            if (l0) pa = v[brev5((32 - k1) & 31)];
            const int ka = 32 * k1 + lane;
            const float2 w = sm.T2[ka];
            const float ex = za.x + pa.x, ey = za.y - pa.y, ox = za.y + pa.y, oy = pa.x - za.x;
            const float tx = ox * w.x - oy * w.y, ty = ox * w.y + oy * w.x;
            const float rx = ex + tx, ry = ey + ty, qx = ex - tx, qy = ey - ty;
            const float ha = (MODE == 1) ? 0.5f * (a.weight ? a.weight[ka] : 1.0f) : 0.f;
            const float hb = (MODE == 1) ? 0.5f * (a.weight ? a.weight[H - ka] : 1.0f) : 0.f;
            curS[k1] = k2w_val<MODE>(rx * rx + ry * ry, ha);        // bin k
            curS[16 + k1] = k2w_val<MODE>(qx * qx + qy * qy, hb);   // bin H - k
            if (TOPDB) pmax = fmaxf(pmax, fmaxf(curS[k1], curS[16 + k1]));
        }
        {   // bin H/2 = conj Z[H/2] (lane 0, register 16)
            const float2 zh = v[brev5(16)];
            curS[32] = k2w_val<MODE>(4.0f * (zh.x * zh.x + zh.y * zh.y), (MODE == 1) ? 0.5f * (a.weight ? a.weight[H / 2] : 1.0f) : 0.f);
            if (TOPDB && l0) pmax = fmaxf(pmax, curS[32]);
        }
        float acc = 0.f;
        if (TOPDB) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
            const float lo = pmax - a.top_db;
#pragma unroll
            for (int k = 0; k < 32; ++k) acc += fmaxf(0.f, fmaxf(curS[k], lo) - fmaxf(prevS[k], lo));
            if (l0) acc += fmaxf(0.f, fmaxf(curS[32], lo) - fmaxf(prevS[32], lo));
        } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) acc += fmaxf(0.f, curS[k] - prevS[k]);
            if (l0) acc += fmaxf(0.f, curS[32] - prevS[32]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (l0 && j >= jw0) a.flux[static_cast<int64_t>(r) * a.n_frames + j] = acc / static_cast<float>(H + 1);
#pragma unroll
        for (int k = 0; k < 33; ++k) prevS[k] = curS[k];
    }
}

}  // namespace ofp
