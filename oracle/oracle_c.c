/*
 * oracle_c.c -- CPU restatement (plain C) of the reference's hot path.
 *
 * TEST INFRASTRUCTURE.  This file is the checker, never the thing measured or shipped:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load the library built from it.  The product path (onset_fingerprinting_b200/)
 * never links or calls it.
 *
 * Parity pin: the restatement is checked against outputs of the UNMODIFIED reference
 * (imported from /root/reference through oracle/ref_harness.py) -- see
 * oracle/make_golden.py and tests/test_oracle_vs_golden.py.  The reference has no tests or
 * golden vectors of its own.
 *
 * Arithmetic contract: IEEE float32, one rounding per
 * written operation, no FMA contraction (compiled with -ffp-contract=off), except
 *   - the single double add inside the attack/release follower,
 *   - log10 / 10**x: the reference uses numpy's float32 ufuncs whose results are
 *     platform dependent (SVML vs glibc, <= 3 ulp apart).  The oracle pins them to the
 *     correctly rounded value: (float)log10((double)x), (float)pow(10.0,(double)q).
 *   - cross-correlation sums: double accumulation in index order, rounded once to float
 *     (np.correlate's blocked float32 summation cannot be bit-matched; H2).
 *
 * All citations are file:line under /root/reference/onset_fingerprinting/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------
 * K1: amplitude onset detector
 * ---------------------------------------------------------------------------------- */

typedef struct {
    int32_t n_channels;  /* C */
    int32_t block_size;  /* B */
    int32_t use_hp;      /* hipass_freq != 0 (detection.py:692-696) */
    int32_t manual;      /* on_threshold > 1 (detection.py:687) */
    int32_t cooldown;    /* detection.py:689 */
    float b[5], a[5];    /* float32(butter(4, f, 'high', fs=sr)) (detection.py:493-496) */
    float floor_db;      /* detection.py:685 */
    float fast_att, fast_rel, slow_att, slow_rel; /* float32(1/x) (detection.py:514-515) */
    float on_thr, off_thr;                        /* detection.py:686,688 */
    float alpha_min, alpha_max, minmin;           /* detection.py:703-708: 1e-4, 1e-5, 2 */
} orc_params;

typedef struct {
    float z[4];     /* lfilter DF2T delay line, detection.py:497 */
    float yf, ys;   /* last row of the followers' y buffers, envelope_follower.c:13-16 */
    float mn, mx;   /* detection.py:556-557 */
    float prev;     /* detection.py:711,792 */
    int32_t state;  /* detection.py:710 */
    int32_t deb;    /* detection.py:712 */
} orc_chan;

ORC_API void orc_state_init(orc_chan *st, const orc_params *p) {
    for (int c = 0; c < p->n_channels; ++c) {
        memset(&st[c], 0, sizeof(orc_chan));
        st[c].yf = p->floor_db;  /* detection.py:697-702 */
        st[c].ys = p->floor_db;
        st[c].mn = 0.0f;         /* detection.py:704: x0 = [[0..],[10..]] */
        st[c].mx = 10.0f;
    }
}

/* scipy.signal.lfilter, order 4, direct form II transposed, float32 (detection.py:499-501).
 * Unfused op order verified bit-exact against scipy. */
static inline float orc_hp(orc_chan *s, const orc_params *p, float x) {
    float y = s->z[0] + p->b[0] * x;
    s->z[0] = (s->z[1] + x * p->b[1]) - y * p->a[1];
    s->z[1] = (s->z[2] + x * p->b[2]) - y * p->a[2];
    s->z[2] = (s->z[3] + x * p->b[3]) - y * p->a[3];
    s->z[3] = x * p->b[4] - y * p->a[4];
    return y;
}

/* envelope_follower.c:15-22 */
static inline float orc_ar(float y, float x, float att, float rel) {
    float t = x - y;
    float d = (float)((double)t + 1e-10);
    return d > 0 ? y + att * d : y + rel * d;
}

/* detection.py:747-748: 20*log10(|x + 1e-10|) clipped at floor, all float32 */
static inline float orc_db(float h, float floor_db) {
    float v = fabsf(h + 1e-10f);
    float l = (float)log10((double)v);
    float db = 20.0f * l;
    /* np.clip(min) -> np.maximum semantics; a NaN input would propagate, -inf -> floor */
    return db < floor_db ? floor_db : db;
}

/* detection.py:751-754 */
static inline float orc_rel(float yf, float ys, float floor_db) {
    float r = yf - ys;
    float q = r / 20.0f;
    float a = (float)pow(10.0, (double)q);
    a = a - 1e-10f;
    if (a < 0.0f) a = 0.0f;
    if (a > -floor_db) a = -floor_db;
    return a;
}

/* envelope_follower.c:38-52 */
static inline void orc_minmax(orc_chan *s, const orc_params *p, float r) {
    float ia_min = (float)(1.0 - (double)p->alpha_min);
    float ia_max = (float)(1.0 - (double)p->alpha_max);
    if (r < p->minmin) s->mn = p->minmin;
    else if (r < s->mn) s->mn = r;
    else s->mn = s->mn * ia_min + r * p->alpha_min;
    if (r > s->mx) s->mx = r;
    else s->mx = s->mx * ia_max + r * p->alpha_max;
}

static inline float orc_front(orc_chan *s, const orc_params *p, float x, int run_hp) {
    float h = (p->use_hp && run_hp) ? orc_hp(s, p, x) : x;
    float db = orc_db(h, p->floor_db);
    s->yf = orc_ar(s->yf, db, p->fast_att, p->fast_rel);
    s->ys = orc_ar(s->ys, db, p->slow_att, p->slow_rel);
    return orc_rel(s->yf, s->ys, p->floor_db);
}

/* AmplitudeOnsetDetector.init_minmax_tracker (detection.py:827-840): the high-pass runs over
 * ALL n samples in one lfilter call, followers and min/max only over the full blocks. */
ORC_API void orc_warmup(const float *x, int64_t n, const orc_params *p, orc_chan *st) {
    const int C = p->n_channels, B = p->block_size;
    int64_t n_full = (n / B) * B;
    for (int64_t i = 0; i < n; ++i) {
        for (int c = 0; c < C; ++c) {
            orc_chan *s = &st[c];
            float v = x[i * C + c];
            float h = p->use_hp ? orc_hp(s, p, v) : v;
            if (i < n_full) {
                float db = orc_db(h, p->floor_db);
                s->yf = orc_ar(s->yf, db, p->fast_att, p->fast_rel);
                s->ys = orc_ar(s->ys, db, p->slow_att, p->slow_rel);
                float r = orc_rel(s->yf, s->ys, p->floor_db);
                orc_minmax(s, p, r);
            }
        }
    }
}

/* AmplitudeOnsetDetector.__call__ (detection.py:727-798) on one [B, C] block.
 * rel: [B, C] out (required, used as scratch).  ch/delta: out, capacity C.
 * Returns the number of onsets (channels ascending). */
ORC_API int orc_block(const float *x, const orc_params *p, orc_chan *st, float *rel,
                      int32_t *ch, int32_t *delta) {
    const int C = p->n_channels, B = p->block_size;
    for (int k = 0; k < B; ++k)
        for (int c = 0; c < C; ++c) rel[k * C + c] = orc_front(&st[c], p, x[k * C + c], 1);
    /* detection.py:759-763 */
    if (!p->manual)
        for (int c = 0; c < C; ++c)
            for (int k = 0; k < B; ++k) orc_minmax(&st[c], p, rel[k * C + c]);
    int n_on = 0, M = 0;
    int32_t oi_all[64];
    int32_t *oi = C <= 64 ? oi_all : (int32_t *)malloc(sizeof(int32_t) * C);
    for (int c = 0; c < C; ++c) {
        orc_chan *s = &st[c];
        float thr_on = p->manual ? p->on_thr : s->mx * p->on_thr + s->mn;
        int first = 0, crossed0 = 0, found = 0;
        if (!s->state && s->deb < 1) { /* detection.py:764-770 */
            for (int k = 0; k < B; ++k) {
                float before = k == 0 ? s->prev : rel[(k - 1) * C + c];
                if (rel[k * C + c] > thr_on && before < thr_on) {
                    first = k; found = 1; crossed0 = (k == 0);
                    break;
                }
            }
        }
        oi[c] = found ? first : 0;                 /* detection.py:774 */
        int hit = (oi[c] > 0) || crossed0;         /* detection.py:775 */
        if (hit) {
            s->state = 1; s->deb = p->cooldown;    /* detection.py:778-779 */
            ch[n_on] = c; delta[n_on] = oi[c]; ++n_on;
        }
        if (s->deb > 0) s->deb -= B;               /* detection.py:780 */
        if (oi[c] > M) M = oi[c];
    }
    for (int c = 0; c < C; ++c) {                  /* detection.py:784-792 */
        orc_chan *s = &st[c];
        float thr_off = p->manual ? p->off_thr : s->mx * p->off_thr + s->mn;
        for (int k = M; k < B; ++k)
            if (rel[k * C + c] < thr_off) { s->state = 0; break; }
        s->prev = rel[(B - 1) * C + c];
    }
    if (oi != oi_all) free(oi);
    return n_on;
}

/* detect_onsets_amplitude (detection.py:19-86): warm-up on x[:warm_n], then the block loop
 * from sample 0; the trailing partial block is dropped.  rel may be NULL.
 * Returns number of onsets found (may exceed cap; only cap are stored). */
ORC_API int64_t orc_detect_offline(const float *x, int64_t n, const orc_params *p, int64_t warm_n,
                                   float *rel, int32_t *out_ch, int64_t *out_idx, int64_t cap,
                                   orc_chan *st_io) {
    const int C = p->n_channels, B = p->block_size;
    orc_chan *st = st_io ? st_io : (orc_chan *)malloc(sizeof(orc_chan) * C);
    if (!st_io) orc_state_init(st, p);
    if (warm_n > n) warm_n = n;
    if (warm_n > 0) orc_warmup(x, warm_n, p, st);
    float *scratch = (float *)malloc(sizeof(float) * B * C);
    int32_t *bch = (int32_t *)malloc(sizeof(int32_t) * C * 2), *bdl = bch + C;
    int64_t n_out = 0;
    for (int64_t i = 0; i + B <= n; i += B) {
        float *r = rel ? rel + i * C : scratch;
        int k = orc_block(x + i * C, p, st, r, bch, bdl);
        for (int j = 0; j < k; ++j) {
            if (n_out < cap) { out_ch[n_out] = bch[j]; out_idx[n_out] = i + bdl[j]; }
            ++n_out;
        }
    }
    free(scratch); free(bch);
    if (!st_io) free(st);
    return n_out;
}

/* Standalone pieces of the ctypes DLL (envelope_follower.c:6-57), same signatures. */
ORC_API void orc_ar_envelope(const float *x, float *y, float attack, float release, int size,
                             int num_samples) {
    for (int j = 0; j < num_samples; ++j)
        for (int i = 0; i < size; ++i) {
            float prev = j > 0 ? y[(j - 1) * size + i] : y[(num_samples - 1) * size + i];
            y[j * size + i] = orc_ar(prev, x[j * size + i], attack, release);
        }
}

ORC_API void orc_minmax_envelope(const float *x, float *mn, float *mx, float a_min, float a_max,
                                 float minmin, int n_samples, int n_channels) {
    orc_params p; memset(&p, 0, sizeof p);
    p.alpha_min = a_min; p.alpha_max = a_max; p.minmin = minmin;
    for (int c = 0; c < n_channels; ++c) {
        orc_chan s; memset(&s, 0, sizeof s);
        s.mn = mn[c]; s.mx = mx[c];
        for (int i = 0; i < n_samples; ++i) orc_minmax(&s, &p, x[i * n_channels + c]);
        mn[c] = s.mn; mx[c] = s.mx;
    }
}

/* backtrack_onsets (detection.py:800-825 == envelope_follower.c:59-85).  buf is the last N
 * rows of rel in time order ([N, C]); deltas are relative to the start of the last block. */
ORC_API void orc_backtrack(const float *buf, const int32_t *ch, int32_t *deltas, float alpha,
                           float tol, int64_t N, int n_onsets, int C, int B) {
    float omba = (float)(1.0 - (double)alpha);  /* np.float32(1 - b_alpha) */
    for (int j = 0; j < n_onsets; ++j) {
        int c = ch[j];
        int64_t i = B - deltas[j];
        float cur = buf[(N - i) * C + c];
        i += 1;
        float prev = buf[(N - i) * C + c];
        float ps = alpha * prev + omba * cur;
        while (cur > ps && fabsf(ps - prev) > tol && i + 1 < N) {
            deltas[j] -= 1;
            i += 1;
            cur = ps;
            prev = buf[(N - i) * C + c];
            ps = alpha * prev + omba * cur;
        }
    }
}

/* ------------------------------------------------------------------------------------
 * K4: bounded-lag cross-correlation, adjust_onset, fix_onsets
 * ---------------------------------------------------------------------------------- */

#define ORC_LAG_NONE INT32_MIN

/* Margin audit (SURVEY.md H2 iii): when non-NULL, orc_cc_lag stores {top-1 value, best other value}
 * in orc_audit[0..1] and orc_adjust_onset stores {da, db} in orc_audit[2..3].  Set per pair by
 * orc_fix_group_audit; never set on the plain paths. */
static __thread double *orc_audit = NULL;

/* Python slice [s:e) on a sequence of length len */
static inline void py_slice(int64_t *s, int64_t *e, int64_t len) {
    if (*s < 0) { *s += len; if (*s < 0) *s = 0; } else if (*s > len) *s = len;
    if (*e < 0) { *e += len; if (*e < 0) *e = 0; } else if (*e > len) *e = len;
}

/* cross_correlation_lag (detection.py:195-268) on already differenced / rectified inputs
 * (stride = element stride).  use_legal: window cc[n-l1 : n-l0]; else centred on onsets.
 * Returns ORC_LAG_NONE for an empty window (reference returns None). */
ORC_API int32_t orc_cc_lag(const float *x, const float *y, int64_t stride, int64_t n, int32_t o0,
                           int32_t o1, int use_legal, int32_t l0, int32_t l1, int32_t cutoff,
                           int32_t tol) {
    int64_t s, e, adj;
    if (use_legal) { s = n - l1; e = n - l0; adj = l1; }            /* detection.py:256-258 */
    else { int64_t cur = (int64_t)o1 - o0; s = n - cur - tol; e = n - cur + tol; adj = cur + tol; }
    py_slice(&s, &e, 2 * n - 1);
    if (e <= s) return ORC_LAG_NONE;                                /* detection.py:265-266 */
    float best = 0; int64_t best_w = -1;
    for (int64_t k = s; k < e; ++k) {
        int64_t m = k - (n - 1); /* np.correlate(x,y,'full')[k] = sum_i x[i+m]*y[i] */
        int64_t i0 = m < 0 ? -m : 0, i1 = m > 0 ? n - m : n;
        double acc = 0.0;
        for (int64_t i = i0; i < i1; ++i) acc += (double)x[(i + m) * stride] * (double)y[i * stride];
        int64_t cnt = n - (m < 0 ? -m : m);                         /* detection.py:247-250 */
        if (cnt < cutoff) cnt = cutoff;
        float v = (float)acc / (float)cnt;
        if (best_w < 0 || v > best) { best = v; best_w = k - s; } /* np.argmax: first max wins */
    }
    if (orc_audit) { /* second pass: the largest value at any other lag */
        double second = -INFINITY;
        for (int64_t k = s; k < e; ++k) {
            if (k - s == best_w) continue;
            int64_t m = k - (n - 1);
            int64_t i0 = m < 0 ? -m : 0, i1 = m > 0 ? n - m : n;
            double acc = 0.0;
            for (int64_t i = i0; i < i1; ++i) acc += (double)x[(i + m) * stride] * (double)y[i * stride];
            int64_t cnt = n - (m < 0 ? -m : m);
            if (cnt < cutoff) cnt = cutoff;
            float v = (float)acc / (float)cnt;
            if (v > second) second = v;
        }
        orc_audit[0] = best; orc_audit[1] = second;
    }
    return (int32_t)(adj - best_w);                                 /* detection.py:268 */
}

/* Window bounds of adjust_onset (detection.py:319-330) */
static inline void adj_bounds(int64_t oa, int64_t ob, int64_t n, int64_t ld, int64_t *xs, int64_t *xe,
                              int64_t *ys, int64_t *ye) {
    if (ld < 0) {
        *xs = oa + ld > 0 ? oa + ld : 0; *xe = oa < n ? oa : n;
        *ys = ob < n ? ob : n;           *ye = ob - ld < n ? ob - ld : n;
    } else {
        *xs = oa;                        *xe = oa + ld < n ? oa + ld : n;
        *ys = ob - ld > 0 ? ob - ld : 0; *ye = ob < n ? ob : n;
    }
}

/* Q10: would the reference raise "operands could not be broadcast" in adjust_onset?
 * x[xs:xe] has max(l,0) elements, exp[-l:] has l (l>0), all k (l==0: exp[-0:]) or
 * max(k+l,0) (l<0) elements; numpy broadcasting accepts equal sizes or a size of 1.
 * The y side is guarded for l == 0 only (detection.py:335-342). */
ORC_API int orc_adjust_would_raise(int32_t oa, int32_t ob, int64_t n, int32_t new_lag) {
    int64_t ld = (int64_t)(ob - oa) - new_lag, k = ld < 0 ? -ld : ld, xs, xe, ys, ye;
    adj_bounds(oa, ob, n, ld, &xs, &xe, &ys, &ye);
    int64_t lx = xe - xs, ly = ye - ys;
    int64_t nx = lx > 0 ? lx : 0, ex = lx > 0 ? lx : (lx == 0 ? k : (k + lx > 0 ? k + lx : 0));
    if (nx != ex && nx != 1 && ex != 1) return 1;
    if (ly != 0) {
        int64_t ny = ly > 0 ? ly : 0, ey = ly > 0 ? ly : (k + ly > 0 ? k + ly : 0);
        if (ny != ey && ny != 1 && ey != 1) return 1;
    }
    return 0;
}

/* adjust_onset (detection.py:299-352).  Returns (ca, cb) through out[2].
 * Sums are accumulated in double in index order (np.sum is pairwise; only the sign of
 * da - db matters and exact ties are 0 == 0 in either order). */
ORC_API void orc_adjust_onset(int32_t oa, int32_t ob, const float *x, const float *y,
                              int64_t stride, int64_t n, int32_t new_lag, int32_t *out) {
    int64_t ld = (int64_t)(ob - oa) - new_lag;
    int64_t k = ld < 0 ? -ld : ld;
    int64_t xs, xe, ys, ye;
    adj_bounds(oa, ob, n, ld, &xs, &xe, &ys, &ye);
    /* np.linspace(0, -e, k): i*step with step = -e/(k-1), last element forced to stop */
    double stop = -M_E, step = k > 1 ? stop / (double)(k - 1) : 0.0;
#define EXPW(i) exp(((i) == k - 1 && k > 1) ? stop : (double)(i) * step)
    float xmax = -INFINITY, ymax = -INFINITY;
    for (int64_t i = 0; i < n; ++i) {
        if (x[i * stride] > xmax) xmax = x[i * stride];
        if (y[i * stride] > ymax) ymax = y[i * stride];
    }
    int64_t lx = xe - xs, ly = ye - ys;
    double da = 0.0, db = 0.0;
    for (int64_t i = 0; i < lx; ++i) da += (double)x[(xs + i) * stride] * EXPW(k - lx + i);
    da = da / (double)xmax;                                   /* 0/0 -> NaN -> "else" branch */
    if (ly != 0) {
        for (int64_t i = 0; i < ly; ++i) db += (double)y[(ys + i) * stride] * EXPW(k - 1 - i);
        db = db / (double)ymax;
    }
#undef EXPW
    if (orc_audit) { orc_audit[2] = da; orc_audit[3] = db; }
    if (da > db && !(oa + ld < 0)) { out[0] = (int32_t)ld; out[1] = 0; }
    else { out[0] = 0; out[1] = (int32_t)(-ld); }             /* Q7: both else branches */
}

static int cmp_float(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* scipy.ndimage.median_filter(sec, size, axes=0) with mode='reflect' (d c b a | a b c d | d c b a),
 * window offsets [-(size/2), size - size/2 - 1] (detection.py:420-422). in/out: [L, C]. */
ORC_API void orc_median_axis0(const float *in, float *out, int64_t L, int C, int size) {
    float w[64];
    int lo = size / 2;
    for (int64_t t = 0; t < L; ++t)
        for (int c = 0; c < C; ++c) {
            for (int j = 0; j < size; ++j) {
                int64_t q = t - lo + j;
                /* reflect, period 2L */
                if (L == 1) q = 0;
                else {
                    int64_t P = 2 * L;
                    q %= P; if (q < 0) q += P;
                    if (q >= L) q = P - 1 - q;
                }
                w[j] = in[q * C + c];
            }
            qsort(w, size, sizeof(float), cmp_float);
            out[t * C + c] = w[size / 2];
        }
}

/* status codes per hit */
#define ORC_FIX_OK 0
#define ORC_FIX_DEGENERATE 1 /* a - look < 0 or section too short: reference raises / wraps (Q6) */
#define ORC_FIX_REF_CRASH 2  /* reference raises ValueError in adjust_onset (Q10) */

static __thread double *orc_audit_base = NULL;

/* fix_onsets (detection.py:373-451) for ONE onset group.  audio [N, C]; og [C] in/out.
 * direction: 0 none, 1 "up", 2 "down".  lags_out [C] (optional): the lag returned by
 * cross_correlation_lag for each later channel (ORC_LAG_NONE where none / reference channel). */
ORC_API int orc_fix_group(const float *audio, int64_t N, int C, int64_t *og, int filter_size, int d,
                          int direction, int take_abs, int zero_left, int cutoff, int tol,
                          int32_t *lags_out) {
    int look = cutoff + tol;                                         /* detection.py:413 */
    int idx[64];
    for (int c = 0; c < C; ++c) idx[c] = c;
    /* np.argsort default (quicksort, not stable for equal keys in general; for n<=16 numpy uses
     * insertion sort which IS stable) -> stable insertion sort */
    for (int i = 1; i < C; ++i) {
        int v = idx[i], j = i - 1;
        while (j >= 0 && og[idx[j]] > og[v]) { idx[j + 1] = idx[j]; --j; }
        idx[j + 1] = v;
    }
    if (lags_out) for (int c = 0; c < C; ++c) lags_out[c] = ORC_LAG_NONE;
    int64_t a = og[idx[0]], b = og[idx[C - 1]];
    int64_t s0 = a - look, s1 = b + look;                            /* detection.py:419 */
    if (s0 < 0) return ORC_FIX_DEGENERATE;
    if (s1 > N) s1 = N;
    int64_t L0 = s1 - s0;
    if (L0 - d < 1) return ORC_FIX_DEGENERATE;
    float *m0 = (float *)malloc(sizeof(float) * L0 * C), *m1 = (float *)malloc(sizeof(float) * L0 * C);
    orc_median_axis0(audio + s0 * C, m0, L0, C, filter_size);
    int64_t L = L0;
    for (int r = 0; r < d; ++r) {                                    /* np.diff(., d, axis=0) */
        for (int64_t t = 0; t + 1 < L; ++t)
            for (int c = 0; c < C; ++c) m1[t * C + c] = m0[(t + 1) * C + c] - m0[t * C + c];
        float *tmp = m0; m0 = m1; m1 = tmp; --L;
    }
    for (int64_t i = 0; i < L * C; ++i) {                            /* detection.py:423-428 */
        if (direction == 1 && m0[i] < 0) m0[i] = 0;
        if (direction == 2 && m0[i] > 0) m0[i] = 0;
        if (take_abs) m0[i] = fabsf(m0[i]);
    }
    int64_t so[64];
    for (int c = 0; c < C; ++c) so[c] = og[c] - s0;                  /* detection.py:429 */
    int status = ORC_FIX_OK;
    int r = idx[0];
    for (int j = 1; j < C; ++j) {
        int i = idx[j];
        int64_t o0 = so[r], o1 = so[i];
        float *x = m0 + r, *y = m0 + i;
        if (zero_left) {                                             /* detection.py:435-437 */
            /* Python slice x[:o0] = 0 (negative o0 counts from the end) */
            int64_t e0 = 0, z0 = o0, e1 = 0, z1 = o1;
            py_slice(&e0, &z0, L); py_slice(&e1, &z1, L);
            for (int64_t t = 0; t < z0; ++t) x[t * C] = 0.0f;
            for (int64_t t = 0; t < z1; ++t) y[t * C] = 0.0f;
        }
        if (orc_audit_base) orc_audit = orc_audit_base + 4 * i;
        int32_t lag = orc_cc_lag(x, y, C, L, (int32_t)o0, (int32_t)o1, 0, 0, 0, cutoff, tol);
        if (lags_out) lags_out[i] = lag;
        if (lag == ORC_LAG_NONE) continue;
        if (orc_adjust_would_raise((int32_t)o0, (int32_t)o1, L, lag)) { status = ORC_FIX_REF_CRASH; break; } /* Q10 */
        int32_t cab[2];
        orc_adjust_onset((int32_t)o0, (int32_t)o1, x, y, C, L, lag, cab);
        og[r] += cab[0]; og[i] += cab[1];                            /* detection.py:447-450 */
        so[r] += cab[0]; so[i] += cab[1];
    }
    free(m0); free(m1);
    orc_audit = NULL;
    return status;
}

/* orc_fix_group + the margin audit: audit [C][4] = {cc top-1, cc best other lag, da, db} per later
 * channel (NaN where the pair was not evaluated). */
ORC_API int orc_fix_group_audit(const float *audio, int64_t N, int C, int64_t *og, int filter_size, int d,
                                int direction, int take_abs, int zero_left, int cutoff, int tol,
                                int32_t *lags_out, double *audit) {
    for (int i = 0; i < 4 * C; ++i) audit[i] = NAN;
    orc_audit_base = audit;
    int st = orc_fix_group(audio, N, C, og, filter_size, d, direction, take_abs, zero_left, cutoff, tol, lags_out);
    orc_audit_base = NULL; orc_audit = NULL;
    return st;
}

/* ------------------------------------------------------------------------------------
 * K5: MINPACK hybrj for n = 2 exactly as fsolve(xtol=0.01, maxfev=20, fprime=...) runs it
 * (multilateration.py:230-316; algorithm: MINPACK hybrj/dogleg/qrfac/qform/r1updt/r1mpyq,
 * scipy 1.18.1's vendored copy).  SURVEY.md Appendix C.
 * ---------------------------------------------------------------------------------- */

typedef struct { double xa, ya, za, xb, yb, zb, xo, yo, zo, da, db; } tri_problem;

static inline double enorm2(double a, double b) {
    /* MINPACK enorm rescales to avoid overflow; for the magnitudes met here (1e-12..1e4) it
     * reduces to sqrt of the plain sum accumulated in order. */
    return sqrt(a * a + b * b);
}

static void tri_f(const tri_problem *q, const double *x, double *f) {
    /* multilateration.py:259-273, z = 0 */
    double X = x[0], Y = x[1];
    double dA = sqrt((X - q->xa) * (X - q->xa) + (Y - q->ya) * (Y - q->ya) + (0.0 - q->za) * (0.0 - q->za));
    double dB = sqrt((X - q->xb) * (X - q->xb) + (Y - q->yb) * (Y - q->yb) + (0.0 - q->zb) * (0.0 - q->zb));
    double dO = sqrt((X - q->xo) * (X - q->xo) + (Y - q->yo) * (Y - q->yo) + (0.0 - q->zo) * (0.0 - q->zo));
    f[0] = dA - dO - q->da;
    f[1] = dB - dO - q->db;
}

static void tri_j(const tri_problem *q, const double *x, double J[2][2]) {
    /* multilateration.py:275-302 */
    double X = x[0], Y = x[1];
    double dA = sqrt((X - q->xa) * (X - q->xa) + (Y - q->ya) * (Y - q->ya) + (0.0 - q->za) * (0.0 - q->za));
    double dB = sqrt((X - q->xb) * (X - q->xb) + (Y - q->yb) * (Y - q->yb) + (0.0 - q->zb) * (0.0 - q->zb));
    double dO = sqrt((X - q->xo) * (X - q->xo) + (Y - q->yo) * (Y - q->yo) + (0.0 - q->zo) * (0.0 - q->zo));
    J[0][0] = (X - q->xa) / dA - (X - q->xo) / dO;
    J[0][1] = (Y - q->ya) / dA - (Y - q->yo) / dO;
    J[1][0] = (X - q->xb) / dB - (X - q->xo) / dO;
    J[1][1] = (Y - q->yb) / dB - (Y - q->yo) / dO;
}

#define HYB_EPS 2.220446049250313e-16
#define HYB_GIANT 1.79769313486231570815e308

double orc_trace[64][2]; int orc_trace_n = 0;
/* returns ier (1 = converged); x in/out */
ORC_API int orc_hybrj2(const tri_problem *q, double *x, double xtol, int maxfev, int *nfev_out) {
    const double p1 = .1, p5 = .5, p001 = 1e-3, p0001 = 1e-4, factor = 100.0;
    double fvec[2], a[2][2], diag[2] = {0, 0}, qtf[2], r11, r12, r22, Q[2][2];
    double wa1[2], wa2[2], wa3[2], wa4[2];
    double delta = 0, xnorm = 0, fnorm, pnorm, fnorm1, actred, prered, ratio, temp, sum;
    int nfev, iter = 1, ncsuc = 0, ncfail = 0, nslow1 = 0, nslow2 = 0, jeval;
    tri_f(q, x, fvec); nfev = 1;
    fnorm = enorm2(fvec[0], fvec[1]);
    int info = 0;
    for (;;) { /* outer loop */
        jeval = 1;
        tri_j(q, x, a);
        /* qrfac, no pivoting; column norms */
        double acnorm[2], rdiag[2];
        acnorm[0] = enorm2(a[0][0], a[1][0]);
        acnorm[1] = enorm2(a[0][1], a[1][1]);
        rdiag[0] = acnorm[0]; rdiag[1] = acnorm[1];
        {   /* j = 0 */
            double ajn = enorm2(a[0][0], a[1][0]);
            if (ajn != 0.0) {
                if (a[0][0] < 0.0) ajn = -ajn;
                a[0][0] /= ajn; a[1][0] /= ajn;
                a[0][0] += 1.0;
                sum = a[0][0] * a[0][1] + a[1][0] * a[1][1];
                temp = sum / a[0][0];
                a[0][1] -= temp * a[0][0];
                a[1][1] -= temp * a[1][0];
                /* (rdiag[1] downdate is only used when pivoting) */
            }
            rdiag[0] = -ajn;
            /* j = 1 */
            ajn = fabs(a[1][1]); /* enorm of a single element */
            if (ajn != 0.0) {
                if (a[1][1] < 0.0) ajn = -ajn;
                a[1][1] /= ajn;
                a[1][1] += 1.0;
            }
            rdiag[1] = -ajn;
        }
        if (iter == 1) {
            for (int j = 0; j < 2; ++j) { diag[j] = acnorm[j]; if (acnorm[j] == 0.0) diag[j] = 1.0; }
            wa3[0] = diag[0] * x[0]; wa3[1] = diag[1] * x[1];
            xnorm = enorm2(wa3[0], wa3[1]);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        /* qtf = Q^T fvec */
        qtf[0] = fvec[0]; qtf[1] = fvec[1];
        if (a[0][0] != 0.0) {
            sum = a[0][0] * qtf[0] + a[1][0] * qtf[1];
            temp = -sum / a[0][0];
            qtf[0] += a[0][0] * temp; qtf[1] += a[1][0] * temp;
        }
        if (a[1][1] != 0.0) {
            sum = a[1][1] * qtf[1];
            temp = -sum / a[1][1];
            qtf[1] += a[1][1] * temp;
        }
        /* copy R (upper triangle by rows) */
        r11 = rdiag[0]; r12 = a[0][1]; r22 = rdiag[1];
        /* qform: accumulate Q from the Householder vectors stored in the lower trapezoid */
        Q[0][0] = a[0][0]; Q[1][0] = a[1][0]; Q[0][1] = 0.0; Q[1][1] = a[1][1];
        for (int k = 1; k >= 0; --k) {
            double w[2] = {0, 0};
            for (int i = k; i < 2; ++i) { w[i] = Q[i][k]; Q[i][k] = 0.0; }
            Q[k][k] = 1.0;
            if (w[k] != 0.0) {
                for (int j = k; j < 2; ++j) {
                    sum = 0.0;
                    for (int i = k; i < 2; ++i) sum += Q[i][j] * w[i];
                    temp = sum / w[k];
                    for (int i = k; i < 2; ++i) Q[i][j] -= temp * w[i];
                }
            }
        }
        for (int j = 0; j < 2; ++j) if (acnorm[j] > diag[j]) diag[j] = acnorm[j]; /* mode 1 rescale */

        for (;;) { /* inner loop */
            /* dogleg */
            double px[2];
            {
                double t2 = r22, t1 = r11;
                if (t2 == 0.0) { double l = fabs(r12) > fabs(r22) ? fabs(r12) : fabs(r22); t2 = HYB_EPS * l; if (t2 == 0.0) t2 = HYB_EPS; }
                px[1] = (qtf[1] - 0.0) / t2;
                if (t1 == 0.0) { double l = fabs(r11); t1 = HYB_EPS * l; if (t1 == 0.0) t1 = HYB_EPS; }
                px[0] = (qtf[0] - r12 * px[1]) / t1;
                double w2[2] = {diag[0] * px[0], diag[1] * px[1]};
                double qnorm = enorm2(w2[0], w2[1]);
                if (qnorm > delta) {
                    double g[2];
                    g[0] = (0.0 + r11 * qtf[0]) / diag[0];
                    g[1] = ((0.0 + r12 * qtf[0]) + r22 * qtf[1]) / diag[1];
                    double gnorm = enorm2(g[0], g[1]);
                    double sgnorm = 0.0, alpha = delta / qnorm;
                    if (gnorm != 0.0) {
                        g[0] = (g[0] / gnorm) / diag[0];
                        g[1] = (g[1] / gnorm) / diag[1];
                        double s0 = (0.0 + r11 * g[0]) + r12 * g[1];
                        double s1 = 0.0 + r22 * g[1];
                        temp = enorm2(s0, s1);
                        sgnorm = gnorm / temp / temp;
                        alpha = 0.0;
                        if (sgnorm < delta) {
                            double bnorm = enorm2(qtf[0], qtf[1]);
                            temp = bnorm / gnorm * (bnorm / qnorm) * (sgnorm / delta);
                            double d1 = sgnorm / delta, d2 = temp - delta / qnorm, d3 = delta / qnorm,
                                   d4 = sgnorm / delta;
                            temp = temp - delta / qnorm * (d1 * d1) +
                                   sqrt(d2 * d2 + (1.0 - d3 * d3) * (1.0 - d4 * d4));
                            double d5 = sgnorm / delta;
                            alpha = delta / qnorm * (1.0 - d5 * d5) / temp;
                        }
                    }
                    temp = (1.0 - alpha) * (sgnorm < delta ? sgnorm : delta);
                    px[0] = temp * g[0] + alpha * px[0];
                    px[1] = temp * g[1] + alpha * px[1];
                }
            }
            for (int j = 0; j < 2; ++j) {
                wa1[j] = -px[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            pnorm = enorm2(wa3[0], wa3[1]);
            if (iter == 1 && pnorm < delta) delta = pnorm;
            if (orc_trace_n < 64) { orc_trace[orc_trace_n][0] = wa2[0]; orc_trace[orc_trace_n][1] = wa2[1]; ++orc_trace_n; }
            tri_f(q, wa2, wa4); ++nfev;
            fnorm1 = enorm2(wa4[0], wa4[1]);
            actred = -1.0;
            if (fnorm1 < fnorm) { double d = fnorm1 / fnorm; actred = 1.0 - d * d; }
            /* predicted reduction: || qtf + R wa1 || */
            wa3[0] = qtf[0] + ((0.0 + r11 * wa1[0]) + r12 * wa1[1]);
            wa3[1] = qtf[1] + (0.0 + r22 * wa1[1]);
            temp = enorm2(wa3[0], wa3[1]);
            prered = 0.0;
            if (temp < fnorm) { double d = temp / fnorm; prered = 1.0 - d * d; }
            ratio = prered > 0.0 ? actred / prered : 0.0;
            if (ratio < p1) { ncsuc = 0; ++ncfail; delta = p5 * delta; }
            else {
                ncfail = 0; ++ncsuc;
                if (ratio >= p5 || ncsuc > 1) { double t = pnorm / p5; if (t > delta) delta = t; }
                if (fabs(ratio - 1.0) <= p1) delta = pnorm / p5;
            }
            if (ratio >= p0001) {
                for (int j = 0; j < 2; ++j) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; fvec[j] = wa4[j]; }
                xnorm = enorm2(wa2[0], wa2[1]);
                fnorm = fnorm1;
                ++iter;
            }
            ++nslow1; if (actred >= p001) nslow1 = 0;
            if (jeval) ++nslow2;
            if (actred >= p1) nslow2 = 0;
            if (delta <= xtol * xnorm || fnorm == 0.0) info = 1;
            if (info != 0) goto done;
            if (nfev >= maxfev) info = 2;
            { double t = p1 * delta; if (pnorm > t) t = pnorm; if (p1 * t <= HYB_EPS * xnorm) info = 3; }
            if (nslow2 == 5) info = 4;
            if (nslow1 == 10) info = 5;
            if (info != 0) goto done;
            if (ncfail == 2) break; /* re-evaluate the Jacobian */
            /* rank-one (Broyden) update */
            for (int j = 0; j < 2; ++j) {
                sum = 0.0;
                for (int i = 0; i < 2; ++i) sum += Q[i][j] * wa4[i];
                wa2[j] = (sum - wa3[j]) / pnorm;
                wa1[j] = diag[j] * (diag[j] * wa1[j] / pnorm);
                if (ratio >= p0001) qtf[j] = sum;
            }
            /* r1updt(m=2,n=2): R + u v^T -> (Q1-rotations) upper-tri; u = wa1, v = wa2, w = wa3 */
            {
                double u0 = wa1[0], u1 = wa1[1], v0 = wa2[0], v1 = wa2[1], w0, w1, cs, sn, tau, cot, tn;
                /* w starts as the last column of s: for n=2 jj points at r22 */
                w1 = r22; w0 = 0.0;
                /* rotate v into a multiple of e_n: j = n-1 = 0 */
                if (v0 != 0.0) {
                    if (fabs(v1) < fabs(v0)) {
                        cot = v1 / v0; sn = p5 / sqrt(0.25 + 0.25 * (cot * cot)); cs = sn * cot;
                        tau = 1.0; if (fabs(cs) * HYB_GIANT > 1.0) tau = 1.0 / cs;
                    } else {
                        tn = v0 / v1; cs = p5 / sqrt(0.25 + 0.25 * (tn * tn)); sn = cs * tn; tau = sn;
                    }
                    v1 = sn * v0 + cs * v1;
                    v0 = tau;
                    /* apply to s (row 0: r11, r12) and w */
                    temp = cs * r11 - sn * w0; w0 = sn * r11 + cs * w0; r11 = temp;
                    temp = cs * r12 - sn * w1; w1 = sn * r12 + cs * w1; r12 = temp;
                }
                /* add the spike from the rank-1 update to w */
                w0 += v1 * u0; w1 += v1 * u1;
                /* eliminate the spike: j = 0 */
                int sing = 0;
                if (w0 != 0.0) {
                    if (fabs(r11) < fabs(w0)) {
                        cot = r11 / w0; sn = p5 / sqrt(0.25 + 0.25 * (cot * cot)); cs = sn * cot;
                        tau = 1.0; if (fabs(cs) * HYB_GIANT > 1.0) tau = 1.0 / cs;
                    } else {
                        tn = w0 / r11; cs = p5 / sqrt(0.25 + 0.25 * (tn * tn)); sn = cs * tn; tau = sn;
                    }
                    temp = cs * r11 + sn * w0; w0 = -sn * r11 + cs * w0; r11 = temp;
                    temp = cs * r12 + sn * w1; w1 = -sn * r12 + cs * w1; r12 = temp;
                    w0 = tau;
                }
                if (r11 == 0.0) sing = 1;
                r22 = w1;
                if (r22 == 0.0) sing = 1;
                (void)sing;
                /* r1mpyq on Q (2x2) and on qtf (1x2) with the stored rotations (v0, w0) */
                double *rows[3] = {Q[0], Q[1], qtf};
                for (int rr = 0; rr < 3; ++rr) {
                    double *A = rows[rr];
                    /* first set: j = n-2 = 0, from v */
                    if (fabs(v0) > 1.0) { cs = 1.0 / v0; sn = sqrt(1.0 - cs * cs); }
                    else { sn = v0; cs = sqrt(1.0 - sn * sn); }
                    temp = cs * A[0] - sn * A[1]; A[1] = sn * A[0] + cs * A[1]; A[0] = temp;
                    /* second set: from w */
                    if (fabs(w0) > 1.0) { cs = 1.0 / w0; sn = sqrt(1.0 - cs * cs); }
                    else { sn = w0; cs = sqrt(1.0 - sn * sn); }
                    temp = cs * A[0] + sn * A[1]; A[1] = -sn * A[0] + cs * A[1]; A[0] = temp;
                }
            }
            jeval = 0;
        }
    }
done:
    if (nfev_out) *nfev_out = nfev;
    return info;
}

/* solve_trilateration_3d (multilateration.py:230-316): returns 1 and writes xy on ier == 1 */
ORC_API int orc_solve_tri3d(const double *sa, const double *sb, const double *so, double da, double db,
                            const double *guess, double *xy) {
    tri_problem q = {sa[0], sa[1], sa[2], sb[0], sb[1], sb[2], so[0], so[1], so[2], da, db};
    double x[2] = {guess[0], guess[1]};
    int ier = orc_hybrj2(&q, x, 0.01, 20, 0);
    xy[0] = x[0]; xy[1] = x[1];
    return ier == 1;
}

/* Per-hit batch equivalent of Multilaterate3D.locate on the three onsets of a hit
 * (multilateration.py:397-426, 428-534, 536-566).
 *   locs [S,3] cm; maps [S,S,H,H] float32 lag maps (NaN outside), max_lags/min_lags [S,S] f32,
 *   max_max [S] f32; sensors[3], onsets[3] in detection order (stable-sorted here).
 * status: 0 located, 1 lag > max_max, 2 is_legal failed, 3 no seed cell, 4 solver failed,
 *         5 degenerate sensor pair (Q8). */
ORC_API int orc_locate3(const double *locs, int S, const float *maps, int H, const float *max_lags,
                        const float *min_lags, const float *max_max, double radius, double samples_per_cm,
                        double sr, double c_cm, const int32_t *sensors, const int64_t *onsets, double *xy) {
    int ord[3] = {0, 1, 2};
    for (int i = 1; i < 3; ++i) { int v = ord[i], j = i - 1; while (j >= 0 && onsets[ord[j]] > onsets[v]) { ord[j + 1] = ord[j]; --j; } ord[j + 1] = v; }
    int s0 = sensors[ord[0]], s1 = sensors[ord[1]], s2 = sensors[ord[2]];
    int64_t o0 = onsets[ord[0]], o1 = onsets[ord[1]], o2 = onsets[ord[2]];
    int64_t lag1 = o1 - o0, lag2 = o2 - o0;
    xy[0] = xy[1] = NAN;
    if ((double)lag1 > (double)max_max[s0] || (double)lag2 > (double)max_max[s0]) return 1;
    if (!((double)min_lags[s0 * S + s1] < (double)lag1 && (double)lag1 < (double)max_lags[s0 * S + s1])) return 2;
    if (!((double)min_lags[s0 * S + s2] < (double)lag2 && (double)lag2 < (double)max_lags[s0 * S + s2])) return 2;
    /* is_legal_3d (multilateration.py:413-426): float32 map vs python float (lag +- tol) -> the
     * comparison promotes the float32 map value to double */
    double tol = 1 * samples_per_cm;
    const float *lm1 = maps + ((int64_t)(s0 * S + s1)) * H * H, *lm2 = maps + ((int64_t)(s0 * S + s2)) * H * H;
    int kfound = 0;
    for (int k = 0; k < H * H; ++k) {
        double m1 = lm1[k], m2 = lm2[k];
        if (m1 < lag1 + tol && m1 > lag1 - tol && m2 < lag2 + tol && m2 > lag2 - tol) { kfound = k; break; }
    }
    int ci = kfound % H, cj = kfound / H; /* np.unravel_index(k, shape, "F") */
    if (ci == 0 && cj == 0) return 3;
    double guess[2] = {ci - radius, cj - radius};
    if (s1 == 1) { s1 = 0; s2 = 1; int64_t t = o1; o1 = o2; o2 = t; } /* Q8, multilateration.py:542-544 */
    if (s1 == s0 || s2 == s0 || s1 == s2) { /* reference still calls fsolve; residual degenerate */ }
    double da = (double)(o1 - o0) / sr * c_cm, db = (double)(o2 - o0) / sr * c_cm;
    return orc_solve_tri3d(locs + 3 * s1, locs + 3 * s2, locs + 3 * s0, da, db, guess, xy) ? 0 : 4;
}
