// Device code of K4 (lag_fix.cu), compiled once per CTA size: included inside a namespace that defines
// `constexpr int K4_THREADS` and `K4_MINCTA`.  No include guard on purpose.


__device__ __forceinline__ void py_slice(int64_t &s, int64_t &e, int64_t len) {
    if (s < 0) { s += len; if (s < 0) s = 0; } else if (s > len) s = len;
    if (e < 0) { e += len; if (e < 0) e = 0; } else if (e > len) e = len;
}

// median of `size` values by rank counting (stable: equal values ranked by position)
__device__ __forceinline__ float median_rank(const float *w, int size) {
    const int want = size / 2;
    float med = w[0];
    for (int i = 0; i < size; ++i) {
        int rank = 0;
        for (int j = 0; j < size; ++j) rank += (w[j] < w[i]) || (w[j] == w[i] && j < i);
        if (rank == want) med = w[i];
    }
    return med;
}

// Median of SIZE values through an optimal sorting network (3 / 9 / 16 / 25 compare-exchanges for
// 3 / 5 / 7 / 9 inputs, each validated with the 0-1 principle): 2 FMNMX per exchange instead of SIZE^2 rank
// comparisons.  Any correct selection returns the same VALUE as scipy's median_filter; NaNs (for which
// fminf / fmaxf are not an ordering) are detected by the caller and sent through rank counting.
__device__ __forceinline__ void cswap(float &a, float &b) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo; b = hi;
}
template <int SIZE>
__device__ __forceinline__ float median_network(float (&w)[SIZE > 0 ? SIZE : 16]) {
#define CS(i, j) cswap(w[i], w[j])
    if (SIZE == 3) { CS(0, 2); CS(0, 1); CS(1, 2); return w[1]; }
    if (SIZE == 5) { CS(0, 3); CS(1, 4); CS(0, 2); CS(1, 3); CS(0, 1); CS(2, 4); CS(1, 2); CS(3, 4); CS(2, 3); return w[2]; }
    if (SIZE == 7) {
        CS(0, 6); CS(2, 3); CS(4, 5); CS(0, 2); CS(1, 4); CS(3, 6); CS(0, 1); CS(2, 5); CS(3, 4); CS(1, 2); CS(4, 6);
        CS(2, 3); CS(4, 5); CS(1, 2); CS(3, 4); CS(5, 6);
        return w[3];
    }
    if (SIZE == 9) {
        CS(0, 3); CS(1, 7); CS(2, 5); CS(4, 8); CS(0, 7); CS(2, 4); CS(3, 8); CS(5, 6); CS(0, 2); CS(1, 3); CS(4, 5);
        CS(7, 8); CS(1, 4); CS(3, 6); CS(5, 7); CS(0, 1); CS(2, 4); CS(3, 5); CS(6, 8); CS(2, 3); CS(4, 5); CS(6, 7);
        CS(1, 2); CS(3, 4); CS(5, 6);
        return w[4];
    }
#undef CS
    return w[0];
}

template <int SIZE>
__device__ __forceinline__ float median_window(const float *src, int64_t L, int C, int64_t t, int c, int size) {
    float w[SIZE > 0 ? SIZE : 16];
    const int n = SIZE > 0 ? SIZE : size;
    const int lo = n / 2;
#pragma unroll
    for (int j = 0; j < (SIZE > 0 ? SIZE : 16); ++j) {
        if (j < n) {
            int q = static_cast<int>(t) - lo + j;  // scipy mode='reflect': d c b a | a b c d | d c b a
            const int Li = static_cast<int>(L);
            if (q < 0 || q >= Li) {               // rare: only within size/2 samples of the section ends
                const int P = 2 * Li;
                q %= P; if (q < 0) q += P;
                if (q >= Li) q = P - 1 - q;
            }
            w[j] = src[q * C + c];
        }
    }
    if (SIZE == 3 || SIZE == 5 || SIZE == 7 || SIZE == 9) {
        float chk = 0.f;
#pragma unroll
        for (int j = 0; j < (SIZE > 0 ? SIZE : 1); ++j) chk += w[j];
        if (chk == chk) return median_network<SIZE>(w);  // no NaN (inf - inf also lands in the rank path)
    }
    if (SIZE > 0) {
        const int want = SIZE / 2;
        float med = w[0];
#pragma unroll
        for (int i = 0; i < SIZE; ++i) {
            int rank = 0;
#pragma unroll
            for (int j = 0; j < SIZE; ++j) rank += (w[j] < w[i]) || (w[j] == w[i] && j < i);
            if (rank == want) med = w[i];
        }
        return med;
    }
    return median_rank(w, n);
}

// K4_RUN consecutive samples of one channel per thread: the SIZE + K4_RUN - 1 samples the windows share are
// loaded once (10 loads for four 7-point medians instead of 28) and are all in flight together.
constexpr int K4_RUN = 4;
template <int SIZE>
__device__ __forceinline__ void median_run(const float *src, int Li, int C, int t0, int c, float *med, int ostride,
                                           int ooff) {
    constexpr int NW = SIZE + K4_RUN - 1, lo = SIZE / 2;
    float w[NW];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        int q = t0 - lo + j;  // scipy mode='reflect': d c b a | a b c d | d c b a
        if (q < 0 || q >= Li) {
            const int P = 2 * Li;
            q %= P; if (q < 0) q += P;
            if (q >= Li) q = P - 1 - q;
        }
        w[j] = src[q * C + c];
    }
    float chk = 0.f;
#pragma unroll
    for (int j = 0; j < NW; ++j) chk += w[j];
    const bool finite = chk == chk;  // a NaN anywhere in the run (or inf - inf): rank counting, as in median_window
#pragma unroll
    for (int k = 0; k < K4_RUN; ++k) {
        if (t0 + k >= Li) break;
        float v[SIZE];
#pragma unroll
        for (int j = 0; j < SIZE; ++j) v[j] = w[k + j];
        med[(t0 + k) * ostride + ooff] = finite ? median_network<SIZE>(v) : median_rank(v, SIZE);
    }
}

__device__ __forceinline__ double block_sum(double v, double *scratch) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < K4_THREADS / 32; ++i) r += scratch[i];
    return r;
}

// two sums / two maxima with one pair of barriers (same per-value summation order as block_sum / block_max)
__device__ __forceinline__ void block_sum2(double &a, double &b, double *scratch) {
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) { scratch[warp] = a; scratch[K4_THREADS / 32 + warp] = b; }
    __syncthreads();
    double ra = 0.0, rb = 0.0;
    for (int i = 0; i < K4_THREADS / 32; ++i) { ra += scratch[i]; rb += scratch[K4_THREADS / 32 + i]; }
    a = ra; b = rb;
}
__device__ __forceinline__ void block_max2(float &a, float &b, float *scratch) {
    for (int o = 16; o > 0; o >>= 1) {
        a = fmaxf(a, __shfl_down_sync(0xffffffffu, a, o));
        b = fmaxf(b, __shfl_down_sync(0xffffffffu, b, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) { scratch[warp] = a; scratch[K4_THREADS / 32 + warp] = b; }
    __syncthreads();
    float ra = scratch[0], rb = scratch[K4_THREADS / 32];
    for (int i = 1; i < K4_THREADS / 32; ++i) { ra = fmaxf(ra, scratch[i]); rb = fmaxf(rb, scratch[K4_THREADS / 32 + i]); }
    a = ra; b = rb;
}

__device__ __forceinline__ float block_max(float v, float *scratch) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = scratch[0];
    for (int i = 1; i < K4_THREADS / 32; ++i) r = fmaxf(r, scratch[i]);
    return r;
}

// Normalised cross-correlation over the window [ws, ws + nl) of np.correlate(x, y, "full") and its
// argmax (detection.py:244-268).  xd, yd: the two signals as doubles in shared memory, length L.
// Result (adj - argmax) is left in *s_lag after a __syncthreads().
__device__ __forceinline__ void cc_argmax(const double *xd, const double *yd, int64_t L, int64_t ws, int64_t nl,
                                          int64_t adj, int cutoff, float *best_v, int *best_w, int *s_lag) {
    const int tid = threadIdx.x;
        // ---- bounded-lag cross-correlation: LPT consecutive lags per thread ----
        float bv = -INFINITY; int bw = INT32_MAX;
        for (int64_t w0 = static_cast<int64_t>(tid) * LPT; w0 < nl; w0 += K4_THREADS * LPT) {
            const int64_t m0 = ws + w0 - (L - 1);  // lag of the first of the LPT windows
            double acc[LPT];
            int64_t i0[LPT], i1[LPT];
#pragma unroll
            for (int u = 0; u < LPT; ++u) {
                const int64_t m = m0 + u;
                acc[u] = 0.0;
                i0[u] = m < 0 ? -m : 0;
                i1[u] = m > 0 ? L - m : L;
                if (w0 + u >= nl) { i0[u] = 0; i1[u] = 0; }
            }
            // common body range [lo, hi): every active lag is valid there
            int64_t lo = 0, hi = L;
#pragma unroll
            for (int u = 0; u < LPT; ++u) if (i1[u] > i0[u]) { lo = max(lo, i0[u]); hi = min(hi, i1[u]); }
            if (hi < lo) hi = lo;
            // heads (ascending i keeps the oracle's summation order)
#pragma unroll
            for (int u = 0; u < LPT; ++u)
                for (int64_t i = i0[u]; i < min(lo, i1[u]); ++i) acc[u] = __fma_rn(xd[i + m0 + u], yd[i], acc[u]);
            {
                const bool full = (w0 + LPT <= nl);
                if (full && lo < hi) {
                    // sliding register window over x: per CC_UN time steps, CC_UN loads of y (broadcast) and
                    // CC_UN of x feed CC_UN * LPT DFMA.  32-bit indices inside the loop.
                    const double *xp = xd + (lo + m0), *yp = yd + lo;
                    const int nbody = static_cast<int>(hi - lo);
                    double xw[LPT - 1 + CC_UN];
#pragma unroll
                    for (int k = 0; k < LPT - 1; ++k) xw[k] = xp[k];
                    int i = 0;
                    for (; i + CC_UN <= nbody; i += CC_UN) {
                        double yv[CC_UN];
#pragma unroll
                        for (int k = 0; k < CC_UN; ++k) { yv[k] = yp[i + k]; xw[LPT - 1 + k] = xp[i + LPT - 1 + k]; }
#pragma unroll
                        for (int k = 0; k < CC_UN; ++k)
#pragma unroll
                            for (int u = 0; u < LPT; ++u) acc[u] = __fma_rn(xw[k + u], yv[k], acc[u]);
#pragma unroll
                        for (int k = 0; k < LPT - 1; ++k) xw[k] = xw[k + CC_UN];
                    }
                    for (; i < nbody; ++i) {
                        const double yv = yp[i];
                        xw[LPT - 1] = xp[i + LPT - 1];
#pragma unroll
                        for (int u = 0; u < LPT; ++u) acc[u] = __fma_rn(xw[u], yv, acc[u]);
#pragma unroll
                        for (int k = 0; k < LPT - 1; ++k) xw[k] = xw[k + 1];
                    }
                } else {
                    for (int64_t i = lo; i < hi; ++i) {
                        const double yv = yd[i];
#pragma unroll
                        for (int u = 0; u < LPT; ++u)
                            if (w0 + u < nl) acc[u] = __fma_rn(xd[i + m0 + u], yv, acc[u]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < LPT; ++u)
                for (int64_t i = max(hi, i0[u]); i < i1[u]; ++i) acc[u] = __fma_rn(xd[i + m0 + u], yd[i], acc[u]);
#pragma unroll
            for (int u = 0; u < LPT; ++u) {
                if (w0 + u < nl) {
                    const int64_t m = m0 + u;
                    int64_t cnt = L - (m < 0 ? -m : m);  // detection.py:247-250
                    if (cnt < cutoff) cnt = cutoff;
                    const float v = __fdiv_rn(__double2float_rn(acc[u]), static_cast<float>(cnt));
                    if (v > bv) { bv = v; bw = static_cast<int>(w0 + u); }  // first maximum wins
                }
            }
        }
        // block argmax: larger value, ties -> smaller window index (np.argmax)
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_down_sync(0xffffffffu, bv, o);
            const int ow = __shfl_down_sync(0xffffffffu, bw, o);
            if (ov > bv || (ov == bv && ow < bw)) { bv = ov; bw = ow; }
        }
        if ((tid & 31) == 0) { best_v[tid >> 5] = bv; best_w[tid >> 5] = bw; }
        __syncthreads();
        if (tid == 0) {
            float v = best_v[0]; int w = best_w[0];
            for (int k = 1; k < K4_THREADS / 32; ++k)
                if (best_v[k] > v || (best_v[k] == v && best_w[k] < w)) { v = best_v[k]; w = best_w[k]; }
            if (w == INT32_MAX) w = 0;  // all NaN / -inf: np.argmax returns 0
            *s_lag = static_cast<int>(adj - w);
        }
        __syncthreads();
}


__device__ __forceinline__ double cc_exact_one(const double *xd, const double *yd, int64_t L, int64_t m) {
    const int64_t i0 = m < 0 ? -m : 0, i1 = m > 0 ? L - m : L;
    double acc = 0.0;
    for (int64_t i = i0; i < i1; ++i) acc = __fma_rn(xd[i + m], yd[i], acc);
    return acc;
}

// Returns false when the screening is inconclusive (caller runs cc_argmax).  Uniform across the block.
__device__ __forceinline__ bool cc_argmax_screen(const double *xd, const double *yd, int64_t L, int64_t ws, int64_t nl,
                                                 int64_t adj, int cutoff, const CcScratch &sc, float *best_v,
                                                 int *best_w, int *s_lag, bool &xf_valid, double &sx_cache) {
    const int tid = threadIdx.x;
    if (nl > K4_THREADS * K4_LPF || L < 64) return false;
    // float copies and the norms
    // (xf_valid: xd is still the signal of an earlier call, its float copy and norm are kept)
    const bool x_fresh = !xf_valid;
    xf_valid = true;
    double sx = 0.0, sy = 0.0;
    for (int64_t i = tid; i < L + 2 * XPAD; i += K4_THREADS) {
        if (x_fresh) {
            const int64_t k = i - XPAD;
            const double xv = (k >= 0 && k < L) ? xd[k] : 0.0;
            sc.xf[i] = static_cast<float>(xv);
            sx += xv * xv;
        }
        if (i < L) { const double yv = yd[i]; sc.yf[i] = static_cast<float>(yv); sy += yv * yv; }
    }
    block_sum2(sx, sy, sc.red_d);
    if (x_fresh) sx_cache = sx; else sx = sx_cache;
    // ||x|| ||y|| as ONE square root (two double square roots per thread per pair were 2.5 % of the kernel's
    // samples); it only enters the error bound, where the 1.0001 factor below covers its rounding
    if (!(sx < 1e30 && sy < 1e30)) return false;  // inf / NaN in the section: exact path
    // ||x|| ||y|| only enters the error bound: a float32 square root rounded up (of the product rounded up) is an
    // upper bound and costs a fraction of the double one (3.9 % of the kernel's samples sat on its result)
    const float S = __fsqrt_ru(__double2float_ru(sx * sy));
    if (sx == 0.0 || sy == 0.0) {   // one signal is all zero: every sum is 0, np.argmax returns index 0
        if (tid == 0) *s_lag = static_cast<int>(adj);
        __syncthreads();
        return true;
    }
    if (tid == 0) sc.cand[0] = 0;
    // pass 1: thread = (lag group g, time segment s)
    const int NG = static_cast<int>((nl + K4_LPF - 1) / K4_LPF);
    const int NS = K4_THREADS / NG;  // >= 1
    const int g = tid % NG, seg = tid / NG;
    float acc[K4_LPF];
#pragma unroll
    for (int u = 0; u < K4_LPF; ++u) acc[u] = 0.f;
    if (seg < NS) {
        const int64_t m0 = ws + static_cast<int64_t>(g) * K4_LPF - (L - 1);  // lag of the group's first window
        const int64_t m1 = m0 + K4_LPF - 1;
        // union of the valid index ranges of the group's lags; outside its own range a lag reads zeros
        int64_t lo = m1 < 0 ? -m1 : 0, hi = m0 > 0 ? L - m0 : L;
        if (lo < 0) lo = 0;
        if (hi > L) hi = L;
        if (hi < lo) hi = lo;
        // reads stay inside [-(K4_LPF-1), L + K4_LPF - 1) of x: covered by XPAD
        const int64_t len = hi - lo, per = (len + NS - 1) / NS;
        const int64_t i0 = lo + seg * per, i1 = min(hi, i0 + per);
        if (i1 > i0) {
            const float *xp = sc.xf + XPAD + i0 + m0, *yp = sc.yf + i0;
            const int n = static_cast<int>(i1 - i0);
            float xw[K4_LPF - 1 + CCF_UN];
#pragma unroll
            for (int k = 0; k < K4_LPF - 1; ++k) xw[k] = xp[k];
            int i = 0;
            for (; i + CCF_UN <= n; i += CCF_UN) {
                float yv[CCF_UN];
#pragma unroll
                for (int k = 0; k < CCF_UN; ++k) { yv[k] = yp[i + k]; xw[K4_LPF - 1 + k] = xp[i + K4_LPF - 1 + k]; }
#pragma unroll
                for (int k = 0; k < CCF_UN; ++k)
#pragma unroll
                    for (int u = 0; u < K4_LPF; ++u) acc[u] = fmaf(xw[k + u], yv[k], acc[u]);
#pragma unroll
                for (int k = 0; k < K4_LPF - 1; ++k) xw[k] = xw[k + CCF_UN];
            }
            for (; i < n; ++i) {
                const float yv = yp[i];
                xw[K4_LPF - 1] = xp[i + K4_LPF - 1];
#pragma unroll
                for (int u = 0; u < K4_LPF; ++u) acc[u] = fmaf(xw[u], yv, acc[u]);
#pragma unroll
                for (int k = 0; k < K4_LPF - 1; ++k) xw[k] = xw[k + 1];
            }
        }
    }
    __syncthreads();  // xf / yf reads done; part aliases nothing else
    if (seg < NS) {
#pragma unroll
        for (int u = 0; u < K4_LPF; ++u) sc.part[(seg * NG + g) * K4_LPF + u] = acc[u];
    }
    __syncthreads();
    // pass 2: v32, bounds, candidates
    const float Ef = S * (static_cast<float>(L + 8) * 5.9604644775390625e-08f) * 1.0002f;
    const int m_first = static_cast<int>(ws - (L - 1)), Lw = static_cast<int>(L), nlw = static_cast<int>(nl);
    auto eval = [&](int w, float &v, float &d) {
        const int gg = w / K4_LPF, u = w - gg * K4_LPF;
        float a = 0.f;
        for (int q = 0; q < NS; ++q) a += sc.part[(q * NG + gg) * K4_LPF + u];
        const int m = m_first + w;
        int cnt = Lw - (m < 0 ? -m : m);
        if (cnt < cutoff) cnt = cutoff;
        const float c = static_cast<float>(cnt);
        v = a / c;
        d = Ef / c * 1.0001f + fabsf(v) * 4.76837158203125e-07f;
    };
    float lowmax = -INFINITY;
    constexpr int EV = 2;  // values per thread kept in registers between the two sweeps (nl <= EV * threads: always
    float ev[EV], ed[EV];  // for the windows of fix_onsets; longer windows re-evaluate)
    const bool keep = nlw <= EV * K4_THREADS;
#pragma unroll
    for (int q = 0; q < EV; ++q) {
        const int w = tid + q * K4_THREADS;
        ev[q] = -INFINITY; ed[q] = 0.f;
        if (keep && w < nlw) { eval(w, ev[q], ed[q]); lowmax = fmaxf(lowmax, ev[q] - ed[q]); }
    }
    if (!keep) {
        for (int w = tid; w < nlw; w += K4_THREADS) {
            float v, d;
            eval(w, v, d);
            lowmax = fmaxf(lowmax, v - d);
        }
    }
    lowmax = block_max(lowmax, sc.red_f);
    __syncthreads();
    if (keep) {
#pragma unroll
        for (int q = 0; q < EV; ++q) {
            const int w = tid + q * K4_THREADS;
            if (w < nlw && ev[q] + ed[q] >= lowmax) {
                const int k = atomicAdd(&sc.cand[0], 1);
                if (k < CAND_CAP) sc.cand[1 + k] = w;
            }
        }
    } else {
        for (int w = tid; w < nlw; w += K4_THREADS) {
            float v, d;
            eval(w, v, d);
            if (v + d >= lowmax) {
                const int k = atomicAdd(&sc.cand[0], 1);
                if (k < CAND_CAP) sc.cand[1 + k] = static_cast<int>(w);
            }
        }
    }
    __syncthreads();
    const int nc = sc.cand[0];
    if (tid == 0) {
        atomicAdd(&g_cc_stats[0], 1ull);
        atomicAdd(&g_cc_stats[nc == 1 ? 1 : (nc >= 1 && nc <= CAND_CAP ? 2 : 3)], 1ull);
    }
    if (nc < 1 || nc > CAND_CAP) return false;  // NaNs (no candidate) or a flat window: exact path
    if (nc == 1) {
        if (tid == 0) *s_lag = static_cast<int>(adj - sc.cand[1]);
        __syncthreads();
        return true;
    }
    // exact values of the survivors, one thread each, then np.argmax's rule (first maximum)
    float bv = -INFINITY;
    int bw = INT32_MAX;
    if (tid < nc) {
        const int w = sc.cand[1 + tid];
        const int64_t m = ws + w - (L - 1);
        int64_t cnt = L - (m < 0 ? -m : m);
        if (cnt < cutoff) cnt = cutoff;
        bv = __fdiv_rn(__double2float_rn(cc_exact_one(xd, yd, L, m)), static_cast<float>(cnt));
        bw = w;
        if (!(bv == bv)) { bv = -INFINITY; bw = INT32_MAX; }  // unreachable for finite data
    }
    if (tid < 32) {
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_down_sync(0xffffffffu, bv, o);
            const int ow = __shfl_down_sync(0xffffffffu, bw, o);
            if (ov > bv || (ov == bv && ow < bw)) { bv = ov; bw = ow; }
        }
        if (tid == 0) *s_lag = static_cast<int>(adj - (bw == INT32_MAX ? 0 : bw));
    }
    __syncthreads();
    return true;
}

// The two exponentially weighted sums of adjust_onset (detection.py:310-342).  Returns false when the
// reference would raise "operands could not be broadcast" (SURVEY Q10).  Uniform across the block.
__device__ __forceinline__ bool adjust_sums(const double *xd, const double *yd, int64_t L, int64_t o0, int64_t o1,
                                            int lag, float xmax, float ymax, double *red_d, double &da, double &db,
                                            int64_t &ld_out) {
    const int tid = threadIdx.x;
        const int64_t ld = (o1 - o0) - lag, k = ld < 0 ? -ld : ld;
        int64_t xs, xe, ys, ye;
        if (ld < 0) {
            xs = o0 + ld > 0 ? o0 + ld : 0; xe = o0 < L ? o0 : L;
            ys = o1 < L ? o1 : L;           ye = o1 - ld < L ? o1 - ld : L;
        } else {
            xs = o0;                        xe = o0 + ld < L ? o0 + ld : L;
            ys = o1 - ld > 0 ? o1 - ld : 0; ye = o1 < L ? o1 : L;
        }
        const int64_t lx = xe - xs, ly = ye - ys;
        {   // Q10: the reference raises "operands could not be broadcast" here
            const int64_t nx = lx > 0 ? lx : 0, ex = lx > 0 ? lx : (lx == 0 ? k : (k + lx > 0 ? k + lx : 0));
            bool bad = nx != ex && nx != 1 && ex != 1;
            if (ly != 0) {
                const int64_t ny = ly > 0 ? ly : 0, ey = ly > 0 ? ly : (k + ly > 0 ? k + ly : 0);
                bad = bad || (ny != ey && ny != 1 && ey != 1);
            }
            if (bad) return false;
        }
        const double stop = -2.718281828459045, step = k > 1 ? stop / static_cast<double>(k - 1) : 0.0;
        double pa = 0.0, pb = 0.0;
        for (int64_t i = tid; i < lx; i += K4_THREADS) {
            const int64_t wi = k - lx + i;
            pa += xd[xs + i] * exp((wi == k - 1 && k > 1) ? stop : static_cast<double>(wi) * step);
        }
        for (int64_t i = tid; i < ly; i += K4_THREADS) {
            const int64_t wi = k - 1 - i;
            pb += yd[ys + i] * exp((wi == k - 1 && k > 1) ? stop : static_cast<double>(wi) * step);
        }
        block_sum2(pa, pb, red_d);
        da = pa; db = pb;
        da = da / static_cast<double>(xmax);
        db = ly != 0 ? db / static_cast<double>(ymax) : 0.0;
        __syncthreads();
        ld_out = ld;
        return true;
}

// Median-filtered samples of ONE channel of the section into col[0 .. Li) (column mode, see k4_fix).
__device__ __forceinline__ void median_column(const float *src, int Li, int C, int c, int size, float *col) {
    const int tid = threadIdx.x;
    if (size == 3 || size == 5 || size == 7 || size == 9) {
        const int n_runs = (Li + K4_RUN - 1) / K4_RUN;
        for (int r = tid; r < n_runs; r += K4_THREADS) {
            const int t0 = r * K4_RUN;
            switch (size) {
                case 3: median_run<3>(src, Li, C, t0, c, col, 1, 0); break;
                case 5: median_run<5>(src, Li, C, t0, c, col, 1, 0); break;
                case 7: median_run<7>(src, Li, C, t0, c, col, 1, 0); break;
                default: median_run<9>(src, Li, C, t0, c, col, 1, 0); break;
            }
        }
    } else {
        for (int t = tid; t < Li; t += K4_THREADS)
            col[t] = size == 1 ? src[t * C + c] : median_window<0>(src, Li, C, t, c, size);
    }
}

__global__ void __launch_bounds__(K4_THREADS, K4_MINCTA) k4_fix(const K4Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = a.C, tid = threadIdx.x, h = blockIdx.x;
    const FixParams fp = a.fp;
    double *xd = reinterpret_cast<double *>(smem_raw);
    double *yd = xd + a.Lmax + 16;
    float *bufA = reinterpret_cast<float *>(yd + a.Lmax + 16);
    CcScratch sc;
    // section storage: every channel's median-filtered samples [Lmax, C] -- or, in column mode (a.columns:
    // many-channel hits, where that buffer alone would cap an SM at 3 hits), only the reference channel's and
    // the current later channel's columns, the latter re-filtered from the L1-resident section once per pair
    const bool columns = a.columns != 0;
    sc.xf = bufA + static_cast<size_t>(a.Lmax | 1) * (columns ? 2 : C);
    sc.yf = sc.xf + a.Lmax + 2 * XPAD;
    sc.part = sc.yf + a.Lmax + 16;
    sc.cand = reinterpret_cast<int *>(sc.part + K4_THREADS * K4_LPF);
    __shared__ int64_t og[32], so[32], zl[32];
    __shared__ int idx[32];
    __shared__ int64_t s_s0, s_L0;
    __shared__ int s_status;
    __shared__ double red_d[2 * (K4_THREADS / 32)];
    __shared__ float red_f[2 * (K4_THREADS / 32)];
    __shared__ float best_v[K4_THREADS / 32];
    __shared__ int best_w[K4_THREADS / 32];
    __shared__ int s_lag;
    sc.red_d = red_d; sc.red_f = red_f;

    if (tid < C) og[tid] = static_cast<int64_t>(a.onsets[static_cast<int64_t>(h) * C + tid]);
    __syncthreads();
    const int look = fp.cutoff + fp.tol;  // detection.py:413
    if (tid == 0) {
        int st = FIX_OK;
        for (int c = 0; c < C; ++c) {
            if (og[c] < 0) st = FIX_INCOMPLETE;  // -1 = channel missing in this group
            og[c] += fp.shift;                   // detection.py:414
            idx[c] = c;
        }
        for (int i = 1; i < C; ++i) {            // np.argsort (stable for <= 16 elements)
            const int v = idx[i];
            int j = i - 1;
            while (j >= 0 && og[idx[j]] > og[v]) { idx[j + 1] = idx[j]; --j; }
            idx[j + 1] = v;
        }
        const int64_t s0 = og[idx[0]] - look;    // detection.py:419
        int64_t s1 = og[idx[C - 1]] + look;
        if (s1 > a.n_samples || fp.to_end) s1 = a.n_samples;
        const int64_t L0 = s1 - s0;
        if (st == FIX_OK && (s0 < 0 || L0 - fp.d < 1)) st = FIX_DEGENERATE;
        if (st == FIX_OK && L0 > a.Lmax) st = FIX_TOO_LONG;
        s_s0 = s0; s_L0 = L0; s_status = st;
    }
    __syncthreads();
    const int64_t s0 = s_s0, L0 = s_L0;
    if (s_status != FIX_OK) {
        if (tid < C) {
            a.out_onsets[static_cast<int64_t>(h) * C + tid] = static_cast<int32_t>(og[tid]);
            if (a.out_lags) a.out_lags[static_cast<int64_t>(h) * C + tid] = LAG_NONE;
        }
        if (tid == 0) a.out_status[h] = s_status;
        return;
    }
    const int64_t rec = a.hit_rec ? a.hit_rec[h] : h;
    const float *src = a.audio + rec * a.rec_stride + s0 * C;
    if (tid < C && a.out_lags) a.out_lags[static_cast<int64_t>(h) * C + tid] = LAG_NONE;
    // median filter along time (detection.py:420-422), read straight from the recording (the 7 rows
    // around a sample are coalesced and L1-resident) into the only section buffer kept in shared memory
    float *med = bufA;  // [C][MS] channel-major with an odd row stride: conflict-free for lanes across channels
    const int MS = a.Lmax | 1;  // (median stores) and for lanes across time (section reads)
    float *colx = bufA, *coly = bufA + a.Lmax;  // column mode
    const int n_el = static_cast<int>(L0) * C;
    if (columns) {
        median_column(src, static_cast<int>(L0), C, idx[0], fp.filter_size, colx);  // the reference channel, once
    } else if (fp.filter_size == 3 || fp.filter_size == 5 || fp.filter_size == 7 || fp.filter_size == 9) {
        // runs of K4_RUN samples of one channel per thread (lanes across channels, then across runs)
        const int Li = static_cast<int>(L0), n_runs = (Li + K4_RUN - 1) / K4_RUN * C;
        for (int r = tid; r < n_runs; r += K4_THREADS) {
            const int tr = r / C, c = r - tr * C, t0 = tr * K4_RUN;
            switch (fp.filter_size) {
                case 3: median_run<3>(src, Li, C, t0, c, med, 1, c * MS); break;
                case 5: median_run<5>(src, Li, C, t0, c, med, 1, c * MS); break;
                case 7: median_run<7>(src, Li, C, t0, c, med, 1, c * MS); break;
                default: median_run<9>(src, Li, C, t0, c, med, 1, c * MS); break;
            }
        }
    } else {
        const int dt_el = K4_THREADS / C, dc_el = K4_THREADS % C;
        int t_el = tid / C, c_el = tid % C;  // (sample, channel) of element e, advanced without a division
        for (int e = tid; e < n_el; e += K4_THREADS) {
            const int64_t t = t_el;
            const int c = c_el;
            t_el += dt_el; c_el += dc_el;
            if (c_el >= C) { c_el -= C; ++t_el; }
            med[c * MS + static_cast<int>(t)] = fp.filter_size == 1 ? src[e] : median_window<0>(src, L0, C, t, c, fp.filter_size);
        }
    }
    const int64_t L = L0 - fp.d;
    if (tid < C) { so[tid] = og[tid] - s0; zl[tid] = 0; }  // detection.py:429
    __syncthreads();
    // section value after np.diff(., d), direction mask, abs (detection.py:420-428) and the in-place
    // zero_left prefixes (435-437), evaluated on the fly from the median-filtered samples
    const int ref_ch = idx[0];
    auto secval = [&](int64_t t, int c) -> float {
        if (t < zl[c]) return 0.0f;
        float w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            w[k] = k <= fp.d ? (columns ? (c == ref_ch ? colx : coly)[t + k] : med[c * MS + t + k]) : 0.0f;
        for (int rr = 0; rr < fp.d; ++rr)
#pragma unroll
            for (int k = 0; k < 3; ++k) w[k] = __fsub_rn(w[k + 1], w[k]);
        float v = w[0];
        if (fp.direction == 1 && v < 0.f) v = 0.f;
        if (fp.direction == 2 && v > 0.f) v = 0.f;
        if (fp.take_abs) v = fabsf(v);
        return v;
    };

    const int r = idx[0];
    int status = FIX_OK;
    int64_t x_zl = -1;
    float xmax_c = 0.f;
    double sx_c = 0.0;
    bool xf_valid = false;  // the screening's float copy / norm of xd
    for (int j = 1; j < C; ++j) {
        const int ci = idx[j];
        const int64_t o0 = so[r], o1 = so[ci];
        if (fp.zero_left) {  // detection.py:435-437 (python slice x[:o] = 0): prefixes only ever grow
            int64_t b0 = 0, z0 = o0, b1 = 0, z1 = o1;
            py_slice(b0, z0, L); py_slice(b1, z1, L);
            __syncthreads();
            if (tid == 0) { zl[r] = max(zl[r], z0); zl[ci] = max(zl[ci], z1); }
            __syncthreads();
        }
        // the reference channel's signal only changes when its zero_left prefix grew: keep xd, its float copy,
        // norm and maximum across pairs otherwise (so[r] moves, but it only selects the window)
        const bool x_fresh = zl[r] != x_zl;
        x_zl = zl[r];
        if (columns) {  // this pair's later channel, median filtered straight from the (L1-resident) section
            median_column(src, static_cast<int>(L0), C, ci, fp.filter_size, coly);
            __syncthreads();
        }
        float xm = -INFINITY, ym = -INFINITY;
        for (int64_t t = tid; t < L; t += K4_THREADS) {
            const float yv = secval(t, ci);
            yd[t] = static_cast<double>(yv);
            ym = fmaxf(ym, yv);
            if (x_fresh) {
                const float xv = secval(t, r);
                xd[t] = static_cast<double>(xv);
                xm = fmaxf(xm, xv);
            }
        }
        block_max2(xm, ym, red_f);
        if (x_fresh) { xmax_c = xm; xf_valid = false; }
        const float xmax = xmax_c, ymax = ym;
        __syncthreads();
        // window of the full CC, detection.py:259-264
        const int64_t cur = o1 - o0;
        int64_t ws = L - cur - fp.tol, we = L - cur + fp.tol;
        const int64_t adj = cur + fp.tol;
        py_slice(ws, we, 2 * L - 1);
        const int64_t nl = we - ws;
        if (nl <= 0) continue;  // detection.py:265-266 -> None, no adjustment
        if (!a.screen || !cc_argmax_screen(xd, yd, L, ws, nl, adj, fp.cutoff, sc, best_v, best_w, &s_lag, xf_valid, sx_c))
            cc_argmax(xd, yd, L, ws, nl, adj, fp.cutoff, best_v, best_w, &s_lag);
        const int lag = s_lag;
        if (tid == 0 && a.out_lags) a.out_lags[static_cast<int64_t>(h) * C + ci] = lag;
        // ---- adjust_onset, detection.py:299-352 ----
        double da, db;
        int64_t ld;
        if (!adjust_sums(xd, yd, L, o0, o1, lag, xmax, ymax, red_d, da, db, ld)) { status = FIX_REF_CRASH; break; }
        if (tid == 0) {
            int64_t ca, cb;
            if (da > db && !(o0 + ld < 0)) { ca = ld; cb = 0; }
            else { ca = 0; cb = -ld; }
            og[r] += ca; og[ci] += cb;  // detection.py:447-450
            so[r] += ca; so[ci] += cb;
        }
        __syncthreads();
    }
    if (tid < C) a.out_onsets[static_cast<int64_t>(h) * C + tid] = static_cast<int32_t>(og[tid]);
    if (tid == 0) a.out_status[h] = status;
}


__device__ __forceinline__ int64_t load_pair(const PairArgs &a, int p, double *xd, double *yd, float *tmp,
                                             float *red_f, float &xmax, float &ymax) {
    // np.diff(., d) then optional abs (detection.py:238-242); tmp: 2*n floats of scratch
    const int tid = threadIdx.x;
    float *tx = tmp, *ty = tmp + a.n;
    for (int t = tid; t < a.n; t += K4_THREADS) {
        tx[t] = a.x[static_cast<int64_t>(p) * a.n + t];
        ty[t] = a.y[static_cast<int64_t>(p) * a.n + t];
    }
    __syncthreads();
    int64_t L = a.n;
    for (int r = 0; r < a.d; ++r) {
        float nx[8], ny[8];  // n <= 8 * K4_THREADS
        int cnt = 0;
        for (int t = tid; t < L - 1; t += K4_THREADS, ++cnt) {
            nx[cnt] = __fsub_rn(tx[t + 1], tx[t]);
            ny[cnt] = __fsub_rn(ty[t + 1], ty[t]);
        }
        __syncthreads();
        cnt = 0;
        for (int t = tid; t < L - 1; t += K4_THREADS, ++cnt) { tx[t] = nx[cnt]; ty[t] = ny[cnt]; }
        __syncthreads();
        --L;
    }
    float xm = -INFINITY, ym = -INFINITY;
    for (int t = tid; t < L; t += K4_THREADS) {
        float xv = tx[t], yv = ty[t];
        if (a.take_abs) { xv = fabsf(xv); yv = fabsf(yv); }
        xd[t] = xv; yd[t] = yv;
        xm = fmaxf(xm, xv); ym = fmaxf(ym, yv);
    }
    xmax = block_max(xm, red_f);
    ymax = block_max(ym, red_f);
    __syncthreads();
    return L;
}

__global__ void __launch_bounds__(K4_THREADS) k4_cc_pairs(const PairArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *xd = reinterpret_cast<double *>(smem_raw);
    double *yd = xd + a.n + 16;
    float *tmp = reinterpret_cast<float *>(yd + a.n + 16);
    __shared__ float red_f[K4_THREADS / 32], best_v[K4_THREADS / 32];
    __shared__ int best_w[K4_THREADS / 32];
    __shared__ int s_lag;
    const int p = blockIdx.x;
    float xmax, ymax;
    const int64_t L = load_pair(a, p, xd, yd, tmp, red_f, xmax, ymax);
    int64_t ws, we, adj;
    if (a.use_legal) {  // detection.py:256-258
        const int64_t l0 = a.onsets[2 * p], l1 = a.onsets[2 * p + 1];
        ws = L - l1; we = L - l0; adj = l1;
    } else {
        const int64_t cur = static_cast<int64_t>(a.onsets[2 * p + 1]) - a.onsets[2 * p];
        ws = L - cur - a.tol; we = L - cur + a.tol; adj = cur + a.tol;
    }
    py_slice(ws, we, 2 * L - 1);
    if (we - ws <= 0 || L <= 0) { if (threadIdx.x == 0) a.out[p] = LAG_NONE; return; }
    __shared__ double red_d[2 * (K4_THREADS / 32)];
    CcScratch sc;
    sc.xf = tmp + 2 * a.n;
    sc.yf = sc.xf + a.n + 2 * XPAD;
    sc.part = sc.yf + a.n + 16;
    sc.cand = reinterpret_cast<int *>(sc.part + K4_THREADS * K4_LPF);
    sc.red_d = red_d; sc.red_f = red_f;
    double sx_unused = 0.0;
    bool xf_valid = false;
    if (!a.screen || !cc_argmax_screen(xd, yd, L, ws, we - ws, adj, a.cutoff, sc, best_v, best_w, &s_lag, xf_valid, sx_unused))
        cc_argmax(xd, yd, L, ws, we - ws, adj, a.cutoff, best_v, best_w, &s_lag);
    if (threadIdx.x == 0) a.out[p] = s_lag;
}

__global__ void __launch_bounds__(K4_THREADS) k4_adjust_pairs(const PairArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *xd = reinterpret_cast<double *>(smem_raw);
    double *yd = xd + a.n + 16;
    float *tmp = reinterpret_cast<float *>(yd + a.n + 16);
    __shared__ float red_f[K4_THREADS / 32];
    __shared__ double red_d[2 * (K4_THREADS / 32)];
    const int p = blockIdx.x;
    float xmax, ymax;
    const int64_t L = load_pair(a, p, xd, yd, tmp, red_f, xmax, ymax);
    const int64_t o0 = a.onsets[2 * p], o1 = a.onsets[2 * p + 1];
    double da, db;
    int64_t ld;
    const bool ok = adjust_sums(xd, yd, L, o0, o1, a.new_lag[p], xmax, ymax, red_d, da, db, ld);
    if (threadIdx.x == 0) {
        if (!ok) { a.out[2 * p] = LAG_NONE; a.out[2 * p + 1] = LAG_NONE; }  // reference raises ValueError
        else if (da > db && !(o0 + ld < 0)) { a.out[2 * p] = static_cast<int32_t>(ld); a.out[2 * p + 1] = 0; }
        else { a.out[2 * p] = 0; a.out[2 * p + 1] = static_cast<int32_t>(-ld); }
    }
}

