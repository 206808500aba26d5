// K5: per-hit TDOA multilateration (sm_100a), one thread per hit, IEEE double, no FMA contraction
// (this file is compiled with -fmad=false).
//
// Replaces Multilaterate3D.is_legal / is_legal_3d / trilaterate and solve_trilateration_3d
// (reference multilateration.py:230-316, 397-426, 536-566) in the per-hit batch form of
// SURVEY.md Appendix D.  The solver is MINPACK hybrj for n = 2 exactly as
// scipy.optimize.fsolve(xtol=0.01, maxfev=20, fprime=...) runs it (Appendix C): fsolve stops long
// before convergence, so a generic Gauss-Newton cannot reproduce its output to 1e-4; replaying the
// algorithm step for step does (bit-identical to scipy when squares are formed by multiplication).
// The same restatement, written for the CPU, is oracle/oracle_c.c:orc_hybrj2.
#include "ofp_common.cuh"

#include <algorithm>

namespace ofp {

typedef struct { double xa, ya, za, xb, yb, zb, xo, yo, zo, da, db; } tri_problem;

__device__ __forceinline__ double enorm2(double a, double b) {
    /* MINPACK enorm rescales to avoid overflow; for the magnitudes met here (1e-12..1e4) it
     * reduces to sqrt of the plain sum accumulated in order. */
    return sqrt(a * a + b * b);
}

__device__ __forceinline__ void tri_f(const tri_problem *q, const double *x, double *f) {
    /* multilateration.py:259-273, z = 0 */
    double X = x[0], Y = x[1];
    double dA = sqrt((X - q->xa) * (X - q->xa) + (Y - q->ya) * (Y - q->ya) + (0.0 - q->za) * (0.0 - q->za));
    double dB = sqrt((X - q->xb) * (X - q->xb) + (Y - q->yb) * (Y - q->yb) + (0.0 - q->zb) * (0.0 - q->zb));
    double dO = sqrt((X - q->xo) * (X - q->xo) + (Y - q->yo) * (Y - q->yo) + (0.0 - q->zo) * (0.0 - q->zo));
    f[0] = dA - dO - q->da;
    f[1] = dB - dO - q->db;
}

__device__ __forceinline__ void tri_j(const tri_problem *q, const double *x, double J[2][2]) {
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

/* returns ier (1 = converged); x in/out */
__device__ int hybrj2(const tri_problem *q, double *x, double xtol, int maxfev, int *nfev_out) {
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


struct K5Args {
    const double *locs;      // [S, 3] cm
    const float *maps;       // [S, S, Hm, Hm] lag maps (NaN = illegal)
    const float *max_lags, *min_lags, *max_max;  // [S,S], [S,S], [S]
    int32_t S, Hm, H, n_per_hit;
    double radius, samples_per_cm, sr, c_cm;
    const int32_t *sensors;  // [H, 3] sensor index of each onset, or null: (0, 1, 2)
    const int32_t *onsets;   // [H, onset_stride]; the first three entries are used
    int32_t onset_stride;
    double *xy;              // [H, 2], NaN when not located
    int32_t *status;         // [H]
    float *pair_lags = nullptr;  // [H, 2] or null.  Non-null: no solve; the two lags trilaterate() would hand to
                                 // Multilaterate3D.model (multilateration.py:553-557) are written instead
};

#ifndef OFP_K5_MINCTA
#define OFP_K5_MINCTA 8
#endif
__global__ void __launch_bounds__(128, OFP_K5_MINCTA) k5_locate(const K5Args a) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= a.H) return;
    const int S = a.S, Hm = a.Hm;
    int sen[3]; long long on[3];
    for (int i = 0; i < 3; ++i) {
        sen[i] = a.sensors ? a.sensors[3 * h + i] : i;
        on[i] = a.onsets[static_cast<int64_t>(h) * a.onset_stride + i];
    }
    // stable argsort of the three onsets
    int ord[3] = {0, 1, 2};
    for (int i = 1; i < 3; ++i) {
        const int v = ord[i];
        int j = i - 1;
        while (j >= 0 && on[ord[j]] > on[v]) { ord[j + 1] = ord[j]; --j; }
        ord[j + 1] = v;
    }
    int s0 = sen[ord[0]], s1 = sen[ord[1]], s2 = sen[ord[2]];
    long long o0 = on[ord[0]], o1 = on[ord[1]], o2 = on[ord[2]];
    const long long lag1 = o1 - o0, lag2 = o2 - o0;
    double out[2] = {__longlong_as_double(0x7ff8000000000000ll), __longlong_as_double(0x7ff8000000000000ll)};
    int st = 0;
    const bool valid = s0 >= 0 && s0 < S && s1 >= 0 && s1 < S && s2 >= 0 && s2 < S && s0 != s1 && s0 != s2;
    if (!valid) st = 5;
    // multilateration.py:439-441, 397-411 (float32 bounds compared with python ints -> exact in double)
    if (st == 0 && (static_cast<double>(lag1) > static_cast<double>(a.max_max[s0]) ||
                    static_cast<double>(lag2) > static_cast<double>(a.max_max[s0]))) st = 1;
    if (st == 0 && !(static_cast<double>(a.min_lags[s0 * S + s1]) < static_cast<double>(lag1) &&
                     static_cast<double>(lag1) < static_cast<double>(a.max_lags[s0 * S + s1]))) st = 2;
    if (st == 0 && !(static_cast<double>(a.min_lags[s0 * S + s2]) < static_cast<double>(lag2) &&
                     static_cast<double>(lag2) < static_cast<double>(a.max_lags[s0 * S + s2]))) st = 2;
    if (st == 0) {
        // is_legal_3d (multilateration.py:413-426): first cell, C-order flatten, unravelled "F"
        const double tol = 1 * a.samples_per_cm;
        const float *lm1 = a.maps + static_cast<int64_t>(s0 * S + s1) * Hm * Hm;
        const float *lm2 = a.maps + static_cast<int64_t>(s0 * S + s2) * Hm * Hm;
        const double l1lo = lag1 - tol, l1hi = lag1 + tol, l2lo = lag2 - tol, l2hi = lag2 + tol;
        int kfound = 0;
        for (int k = 0; k < Hm * Hm; ++k) {
            const double m1 = lm1[k], m2 = lm2[k];
            if (m1 < l1hi && m1 > l1lo && m2 < l2hi && m2 > l2lo) { kfound = k; break; }
        }
        const int ci = kfound % Hm, cj = kfound / Hm;
        if (ci == 0 && cj == 0) st = 3;
        else {
            double x[2] = {ci - a.radius, cj - a.radius};
            if (s1 == 1) { s1 = 0; s2 = 1; const long long t = o1; o1 = o2; o2 = t; }  // Q8, multilateration.py:542-544
            if (a.pair_lags != nullptr) {  // model bypass: the network sees (d_a1, d_b1) as float32
                a.pair_lags[2 * static_cast<int64_t>(h)] = static_cast<float>(o1 - o0);
                a.pair_lags[2 * static_cast<int64_t>(h) + 1] = static_cast<float>(o2 - o0);
                a.xy[2 * static_cast<int64_t>(h)] = out[0];
                a.xy[2 * static_cast<int64_t>(h) + 1] = out[1];
                a.status[h] = 0;
                return;
            }
            tri_problem q;
            q.xa = a.locs[3 * s1]; q.ya = a.locs[3 * s1 + 1]; q.za = a.locs[3 * s1 + 2];
            q.xb = a.locs[3 * s2]; q.yb = a.locs[3 * s2 + 1]; q.zb = a.locs[3 * s2 + 2];
            q.xo = a.locs[3 * s0]; q.yo = a.locs[3 * s0 + 1]; q.zo = a.locs[3 * s0 + 2];
            q.da = static_cast<double>(o1 - o0) / a.sr * a.c_cm;  // multilateration.py:563-564
            q.db = static_cast<double>(o2 - o0) / a.sr * a.c_cm;
            const int ier = hybrj2(&q, x, 0.01, 20, nullptr);
            if (ier == 1) { out[0] = x[0]; out[1] = x[1]; }
            else st = 4;
        }
    }
    a.xy[2 * static_cast<int64_t>(h)] = out[0];
    a.xy[2 * static_cast<int64_t>(h) + 1] = out[1];
    a.status[h] = st;
}

// ---------------------------------------------------------------------------------------------
// Streaming locate for many concurrent streams: PlayRec.detect_hits + Multilaterate3D.locate
// (realtime/audio.py:62-74, multilateration.py:428-534 with rec_audio = None), one thread per stream.
// Each stream keeps its `ongoing` list of onset groups in device memory; a block's detections (K1's
// per-block output) are taken in sample order and fed through the reference's group state machine
// statement by statement -- including its quirks: the in-place swap when an onset lies before a
// group's first one (which also replaces the detection for the rest of the loop), the extended
// group being appended twice, `break` on a repeated first sensor, and every group after a solve
// being dropped.  A block stops at its first located hit like detect_hits does.
// ---------------------------------------------------------------------------------------------
constexpr int SL_GMAX = 16;  // groups kept per stream
constexpr int SL_LEN = 4;    // members kept per group (the reference solves at exactly three)

struct SlArgs {
    K5Args geo;               // lag maps, bounds, sensor positions (hit fields unused)
    int32_t n_streams, C;
    const int32_t *det_ch, *det_delta, *det_cnt;  // [S, C], [S, C], [S]: K1's per-block output
    int64_t current_index;    // sample index of the block start
    const int64_t *current_index_dev = nullptr;  // when set: read from device memory instead (graph replay)
    int32_t *g_count;         // [S]
    int32_t *g_len;           // [S, GMAX]
    int32_t *g_sensor;        // [S, GMAX, LEN]
    int64_t *g_onset;         // [S, GMAX, LEN]
    double *xy;               // [S, 2]
    int32_t *found;           // [S]: 1 located, 0 nothing, -1 state overflow (groups or members dropped),
                              //      -2 the reference would have raised inside adjust_onset
    // ring-buffer refinement (multilateration.py:457-501), RING instantiation only
    const float *ring = nullptr;  // [S, ring_rows, C] most recent audio rows of every stream (row = sample % ring_rows)
    int32_t ring_rows = 0, block = 0;
    int32_t tol = 50, cutoff = 10;  // ONSET_TOL, NORM_CUTOFF (multilateration.py:18-19); lookaround = their sum
};

struct SlGroup { int len; int s[SL_LEN]; long long o[SL_LEN]; };

__device__ __forceinline__ bool sl_is_legal(const K5Args &a, int first, int later, long long lag) {
    const double l = static_cast<double>(lag);
    return static_cast<double>(a.min_lags[first * a.S + later]) < l && l < static_cast<double>(a.max_lags[first * a.S + later]);
}

// ---- ring-buffer refinement of one (group, detection) pair by ONE WARP (multilateration.py:457-501) ----
// section = ring rows [last_onset - lookaround - 1, counter) of the two sensors' columns -> median 5 (scipy
// 'reflect') -> first difference -> rising flanks zeroed, abs -> cross_correlation_lag(onsets, d = 0, tol 50,
// cutoff 10) (detection.py:195-268) -> adjust_onset (299-352).  Arithmetic as in oracle_c.c: products exact in
// double, sums in double in index order per lag, one rounding to float32, float32 division by the count,
// first maximum wins.
constexpr int SL_LMAX = 1024;  // longest refinement section (samples); longer ones raise the overflow flag
struct SlRefine { int has, err; long long lag, co, cn; };

__device__ __forceinline__ void sl_py_slice(long long &s, long long &e, long long len) {
    if (s < 0) { s += len; if (s < 0) s = 0; } else if (s > len) s = len;
    if (e < 0) { e += len; if (e < 0) e = 0; } else if (e > len) e = len;
}

__device__ __forceinline__ float sl_med5(float a0, float a1, float a2, float a3, float a4) {
    // median of five by an optimal 9-exchange network (values only; inputs are finite audio samples)
#define SL_CS(x, y) { const float lo_ = fminf(x, y), hi_ = fmaxf(x, y); x = lo_; y = hi_; }
    SL_CS(a0, a3); SL_CS(a1, a4); SL_CS(a0, a2); SL_CS(a1, a3); SL_CS(a0, a1); SL_CS(a2, a4); SL_CS(a1, a2);
    SL_CS(a3, a4); SL_CS(a2, a3);
#undef SL_CS
    return a2;
}

__device__ SlRefine sl_refine_pair(const SlArgs &a, int st, int lane, double *xd, double *yd, int first_sensor,
                                   int sensor, long long last_onset, long long onset, long long counter) {
    SlRefine r;
    r.has = 0; r.err = 0; r.lag = 0; r.co = 0; r.cn = 0;
    const int C = a.C, NR = a.ring_rows, tol = a.tol, cutoff = a.cutoff, look = a.tol + a.cutoff;
    long long rows = counter - last_onset + look + 1;  // rec_audio[-i - 1:], i = counter - last_onset + lookaround
    if (rows > NR) rows = NR;                          // python slice past the start: the whole ring
    if (rows < 2) return r;
    if (rows - 1 > SL_LMAX) { r.err = 1; return r; }
    const long long start = counter - rows;            // absolute sample index of section row 0
    const int L0 = static_cast<int>(rows), n = L0 - 1;
    const float *ring = a.ring + static_cast<int64_t>(st) * NR * C;
    auto sample = [&](int t, int ch) -> float {         // section row t (reflected at the ends), channel ch
        if (t < 0) t = -t - 1;
        if (t >= L0) t = 2 * L0 - 1 - t;
        t = min(max(t, 0), L0 - 1);
        const long long sidx = start + t;
        if (sidx < 0) return 0.0f;                      // before the stream started: the ring is zero-initialised
        return ring[static_cast<int64_t>(sidx % NR) * C + ch];
    };
    auto med = [&](int t, int ch) -> float {
        return sl_med5(sample(t - 2, ch), sample(t - 1, ch), sample(t, ch), sample(t + 1, ch), sample(t + 2, ch));
    };
    float xm = -INFINITY, ym = -INFINITY;
    for (int t = lane; t < n; t += 32) {
        float dx = __fsub_rn(med(t + 1, first_sensor), med(t, first_sensor));
        float dy = __fsub_rn(med(t + 1, sensor), med(t, sensor));
        dx = dx >= 0.0f ? 0.0f : fabsf(dx);             // section[section >= 0] = 0; abs
        dy = dy >= 0.0f ? 0.0f : fabsf(dy);
        xd[t] = static_cast<double>(dx);
        yd[t] = static_cast<double>(dy);
        xm = fmaxf(xm, dx); ym = fmaxf(ym, dy);
    }
    for (int o = 16; o > 0; o >>= 1) {
        xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, o));
        ym = fmaxf(ym, __shfl_xor_sync(0xffffffffu, ym, o));
    }
    __syncwarp();
    // ---- cross_correlation_lag, window centred on the current lag (detection.py:259-268) ----
    const long long cur = onset - last_onset;
    long long ws = n - cur - tol, we = n - cur + tol;
    sl_py_slice(ws, we, 2ll * n - 1);
    if (we - ws <= 0) return r;                         // None: no adjustment
    float best = 0.0f;
    long long best_k = -1;
    for (long long k = ws + lane; k < we; k += 32) {
        const long long m = k - (n - 1);                // np.correlate(x, y, 'full')[k] = sum_i x[i + m] * y[i]
        const int i0 = m < 0 ? static_cast<int>(-m) : 0, i1 = m > 0 ? static_cast<int>(n - m) : n;
        double acc = 0.0;
        for (int i = i0; i < i1; ++i) acc = fma(xd[i + m], yd[i], acc);  // products of two float32 are exact in double
        long long cnt = n - (m < 0 ? -m : m);
        if (cnt < cutoff) cnt = cutoff;
        const float v = __fdiv_rn(static_cast<float>(acc), static_cast<float>(cnt));
        if (best_k < 0 || v > best) { best = v; best_k = k; }
    }
    for (int o = 16; o > 0; o >>= 1) {                  // first maximum over the lanes
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const long long ok = __shfl_xor_sync(0xffffffffu, best_k, o);
        if (ok >= 0 && (best_k < 0 || ov > best || (ov == best && ok < best_k))) { best = ov; best_k = ok; }
    }
    const long long lag = cur + tol - (best_k - ws);
    // ---- adjust_onset on section-relative onsets (lookaround, lookaround + cur) ----
    const long long oa = look, ob = look + cur, ld = (ob - oa) - lag, kk = ld < 0 ? -ld : ld;
    long long xs, xe, ys, ye;
    if (ld < 0) { xs = oa + ld > 0 ? oa + ld : 0; xe = oa < n ? oa : n; ys = ob < n ? ob : n; ye = ob - ld < n ? ob - ld : n; }
    else { xs = oa; xe = oa + ld < n ? oa + ld : n; ys = ob - ld > 0 ? ob - ld : 0; ye = ob < n ? ob : n; }
    const long long lx = xe - xs, ly = ye - ys;
    {   // would the reference raise "operands could not be broadcast" (oracle_c.c:orc_adjust_would_raise)?
        const long long nx = lx > 0 ? lx : 0, ex = lx > 0 ? lx : (lx == 0 ? kk : (kk + lx > 0 ? kk + lx : 0));
        bool bad = nx != ex && nx != 1 && ex != 1;
        if (ly != 0) {
            const long long ny = ly > 0 ? ly : 0, ey = ly > 0 ? ly : (kk + ly > 0 ? kk + ly : 0);
            bad = bad || (ny != ey && ny != 1 && ey != 1);
        }
        if (bad) { r.err = 2; return r; }
    }
    const double stop = -2.718281828459045, step = kk > 1 ? stop / static_cast<double>(kk - 1) : 0.0;
    auto expw = [&](long long i) { return exp((i == kk - 1 && kk > 1) ? stop : static_cast<double>(i) * step); };
    double da = 0.0, db = 0.0;
    for (long long i = lane; i < lx; i += 32) da += xd[xs + i] * expw(kk - lx + i);
    for (long long i = lane; i < ly; i += 32) db += yd[ys + i] * expw(kk - 1 - i);
    for (int o = 16; o > 0; o >>= 1) {
        da += __shfl_xor_sync(0xffffffffu, da, o);
        db += __shfl_xor_sync(0xffffffffu, db, o);
    }
    da = da / static_cast<double>(xm);
    if (ly != 0) db = db / static_cast<double>(ym); else db = 0.0;
    r.has = 1; r.lag = lag;
    if (da > db && !(oa + ld < 0)) { r.co = ld; r.cn = 0; }
    else { r.co = 0; r.cn = -ld; }
    __syncwarp();
    return r;
}

// RING = false: one thread per stream.  RING = true: one WARP per stream -- every lane runs the (uniform) group
// state machine, the refinement of a pair is shared by the 32 lanes, lane 0 writes the state back.
// Group identity: the reference appends the SAME tuple twice when a pair is legal, and mutates group lists in
// place (the swap at 443-449, `group[1][0] += co` at 497) -- the twin sees the mutation.  Stored groups carry a
// uid (upper bits of g_len); an entry with the uid of the entry processed just before it inherits that entry's
// (possibly mutated) first sensor / onset.
template <bool RING>
__global__ void k5_stream_locate(const SlArgs a) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int st = RING ? gtid >> 5 : gtid;
    const int lane = RING ? (threadIdx.x & 31) : 0;
    if (st >= a.n_streams) return;
    extern __shared__ __align__(16) double sl_smem[];
    double *xd = RING ? sl_smem + static_cast<size_t>(threadIdx.x >> 5) * 2 * SL_LMAX : nullptr;
    double *yd = RING ? xd + SL_LMAX : nullptr;
    const K5Args &geo = a.geo;
    const int S = geo.S, Hm = geo.Hm;
    const long long cur_index = a.current_index_dev ? *a.current_index_dev : a.current_index;
    const long long counter = cur_index + a.block;  // rows written to the ring so far (incl. this block)
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double out[2] = {nan, nan};
    int found = 0;
    bool overflow = false, raised = false;
    // detections of this block in sample order (np.argsort on <= C values: stable insertion sort)
    const int nd = min(a.det_cnt[st], a.C);
    int ord[32];
    for (int i = 0; i < nd; ++i) {
        int j = i - 1;
        const int v = a.det_delta[static_cast<int64_t>(st) * a.C + i];
        while (j >= 0 && a.det_delta[static_cast<int64_t>(st) * a.C + ord[j]] > v) { ord[j + 1] = ord[j]; --j; }
        ord[j + 1] = i;
    }
    int32_t *gl = a.g_len + static_cast<int64_t>(st) * SL_GMAX;
    int32_t *gs = a.g_sensor + static_cast<int64_t>(st) * SL_GMAX * SL_LEN;
    int64_t *go = a.g_onset + static_cast<int64_t>(st) * SL_GMAX * SL_LEN;
    int ng = a.g_count[st] & 0xff;
    int next_uid = a.g_count[st] >> 8;
    // the previous contents are read by every lane before lane 0 overwrites them below: work on a register copy
    SlGroup cur_g[SL_GMAX];
    int cur_uid[SL_GMAX];
    for (int gi = 0; gi < ng; ++gi) {
        cur_g[gi].len = gl[gi] & 0xff;
        cur_uid[gi] = gl[gi] >> 8;
        for (int k = 0; k < SL_LEN; ++k) { cur_g[gi].s[k] = gs[gi * SL_LEN + k]; cur_g[gi].o[k] = go[gi * SL_LEN + k]; }
    }
    for (int di = 0; di < nd && !found; ++di) {
        int sensor = a.det_ch[static_cast<int64_t>(st) * a.C + ord[di]];
        long long onset = cur_index + a.det_delta[static_cast<int64_t>(st) * a.C + ord[di]];
        // ---- Multilaterate3D.locate(sensor, onset) ----
        SlGroup ngp[SL_GMAX];  // new_groups
        int nuid[SL_GMAX];
        int nn = 0;
        auto push = [&](const SlGroup &g, int uid) {
            if (nn < SL_GMAX) { ngp[nn] = g; nuid[nn] = uid; ++nn; } else overflow = true;
        };
        bool returned = false, broke = false;
        int prev_uid = -1, prev_s0 = 0;
        long long prev_o0 = 0;
        for (int gi = 0; gi < ng && !returned && !broke; ++gi) {
            SlGroup g = cur_g[gi];
            int uid = cur_uid[gi];
            if (uid == prev_uid) { g.s[0] = prev_s0; g.o[0] = prev_o0; }  // the same tuple, already mutated
            long long lag = onset - g.o[0];
            if (static_cast<double>(lag) > static_cast<double>(geo.max_max[g.s[0]])) { prev_uid = uid; prev_s0 = g.s[0]; prev_o0 = g.o[0]; continue; }
            if (lag < 0) {  // multilateration.py:443-449
                const int ts = g.s[0]; const long long to = g.o[0];
                g.s[0] = sensor; g.o[0] = onset;
                sensor = ts; onset = to;
                lag = -lag;
            }
            bool member = false;
            for (int k = 0; k < g.len; ++k) member = member || g.s[k] == sensor;
            if (!member) {
                if (RING) {  // multilateration.py:457-501
                    const SlRefine rf = sl_refine_pair(a, st, lane, xd, yd, g.s[0], sensor, g.o[0], onset, counter);
                    if (rf.err == 1) overflow = true;
                    if (rf.err == 2) raised = true;
                    if (rf.has) { lag = rf.lag; g.o[0] += rf.co; onset += rf.cn; }
                }
                prev_uid = uid; prev_s0 = g.s[0]; prev_o0 = g.o[0];  // what the shared lists hold from here on
                if (sl_is_legal(geo, g.s[0], sensor, lag)) {
                    if (g.len < SL_LEN) { g.s[g.len] = sensor; g.o[g.len] = onset; ++g.len; }
                    else overflow = true;
                    uid = next_uid++;  // group[0] + [..] builds a NEW tuple
                    if (g.len == 3) {
                        if (g.s[0] == g.s[1]) { broke = true; break; }
                        // is_legal_3d (multilateration.py:413-426)
                        const double tol = 1 * geo.samples_per_cm;
                        const long long lag1 = g.o[1] - g.o[0], lag2 = g.o[2] - g.o[0];
                        const float *lm1 = geo.maps + static_cast<int64_t>(g.s[0] * S + g.s[1]) * Hm * Hm;
                        const float *lm2 = geo.maps + static_cast<int64_t>(g.s[0] * S + g.s[2]) * Hm * Hm;
                        const double l1lo = lag1 - tol, l1hi = lag1 + tol, l2lo = lag2 - tol, l2hi = lag2 + tol;
                        int kf = 0;
                        for (int k = 0; k < Hm * Hm; ++k) {
                            const double m1 = lm1[k], m2 = lm2[k];
                            if (m1 < l1hi && m1 > l1lo && m2 < l2hi && m2 > l2lo) { kf = k; break; }
                        }
                        const int ci = kf % Hm, cj = kf / Hm;
                        if (!(ci == 0 && cj == 0)) {
                            // trilaterate (multilateration.py:536-575) incl. the sensor rewrite (Q8)
                            int s0 = g.s[0], s1 = g.s[1], s2 = g.s[2];
                            long long o0 = g.o[0], o1 = g.o[1], o2 = g.o[2];
                            if (s1 == 1) { s1 = 0; s2 = 1; const long long t = o1; o1 = o2; o2 = t; }
                            double x[2] = {ci - geo.radius, cj - geo.radius};
                            tri_problem q;
                            q.xa = geo.locs[3 * s1]; q.ya = geo.locs[3 * s1 + 1]; q.za = geo.locs[3 * s1 + 2];
                            q.xb = geo.locs[3 * s2]; q.yb = geo.locs[3 * s2 + 1]; q.zb = geo.locs[3 * s2 + 2];
                            q.xo = geo.locs[3 * s0]; q.yo = geo.locs[3 * s0 + 1]; q.zo = geo.locs[3 * s0 + 2];
                            q.da = static_cast<double>(o1 - o0) / geo.sr * geo.c_cm;
                            q.db = static_cast<double>(o2 - o0) / geo.sr * geo.c_cm;
                            const int ier = hybrj2(&q, x, 0.01, 20, nullptr);
                            if (ier == 1) {
                                out[0] = x[0]; out[1] = x[1]; found = 1;
                                // remove_seed (multilateration.py:160-167): same first sensor and onset
                                int w = 0;
                                for (int r = 0; r < nn; ++r)
                                    if (!(ngp[r].s[0] == g.s[0] && ngp[r].o[0] == g.o[0])) { ngp[w] = ngp[r]; nuid[w] = nuid[r]; ++w; }
                                nn = w;
                            }
                            returned = true;  // self.ongoing = new_groups; return res
                            break;
                        }
                    }
                    push(g, uid);
                }
            } else {
                prev_uid = uid; prev_s0 = g.s[0]; prev_o0 = g.o[0];
            }
            if (static_cast<double>(lag) <= static_cast<double>(geo.max_max[g.s[0]])) push(g, uid);
        }
        if (!returned) {
            SlGroup single;
            single.len = 1;
            for (int k = 0; k < SL_LEN; ++k) { single.s[k] = -1; single.o[k] = 0; }
            single.s[0] = sensor; single.o[0] = onset;
            push(single, next_uid++);
        }
        ng = nn;
        for (int gi = 0; gi < ng; ++gi) { cur_g[gi] = ngp[gi]; cur_uid[gi] = nuid[gi]; }
    }
    if (lane == 0) {
        for (int gi = 0; gi < ng; ++gi) {
            gl[gi] = cur_g[gi].len | (cur_uid[gi] << 8);
            for (int k = 0; k < SL_LEN; ++k) { gs[gi * SL_LEN + k] = cur_g[gi].s[k]; go[gi * SL_LEN + k] = cur_g[gi].o[k]; }
        }
        a.g_count[st] = ng | ((next_uid & 0x7fffff) << 8);
        a.xy[2 * static_cast<int64_t>(st)] = out[0];
        a.xy[2 * static_cast<int64_t>(st) + 1] = out[1];
        a.found[st] = found ? 1 : (raised ? -2 : (overflow ? -1 : 0));
    }
}

// One block of every stream into its ring (row = sample index % ring_rows); blocks [S, B, C] at stream stride.
__global__ void k_ring_write(float *ring, int32_t ring_rows, const float *blocks, int64_t stream_stride, int32_t S,
                             int32_t B, int32_t C, const int64_t *current_index_dev, int64_t current_index) {
    const int64_t base = current_index_dev ? *current_index_dev : current_index;
    const int64_t per = static_cast<int64_t>(B) * C, total = per * S;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t s = i / per, e = i - s * per, row = e / C, c = e - row * C;
        ring[(s * ring_rows + (base + row) % ring_rows) * C + c] = blocks[s * stream_stride + e];
    }
}

// solve_trilateration / solve_trilateration_3d (multilateration.py:170-316) with an explicit seed, one
// thread per problem: prob [P, 11] = (sensor_a xyz, sensor_b xyz, sensor_origin xyz, delta_d_a, delta_d_b).
__global__ void k5_solve(const double *prob, const double *seed, int P, double xtol, int maxfev, double *xy,
                         int32_t *ier_out, int32_t *nfev_out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const double *v = prob + 11 * static_cast<int64_t>(p);
    tri_problem q;
    q.xa = v[0]; q.ya = v[1]; q.za = v[2]; q.xb = v[3]; q.yb = v[4]; q.zb = v[5];
    q.xo = v[6]; q.yo = v[7]; q.zo = v[8]; q.da = v[9]; q.db = v[10];
    double x[2] = {seed[2 * p], seed[2 * p + 1]};
    int nfev = 0;
    const int ier = hybrj2(&q, x, xtol, maxfev, &nfev);
    xy[2 * static_cast<int64_t>(p)] = x[0];
    xy[2 * static_cast<int64_t>(p) + 1] = x[1];
    ier_out[p] = ier;
    if (nfev_out) nfev_out[p] = nfev;
}

}  // namespace ofp

using namespace ofp;

extern "C" int ofp_stream_locate_state_bytes(int32_t n_streams, int64_t *count_bytes, int64_t *len_bytes,
                                             int64_t *sensor_bytes, int64_t *onset_bytes) {
    OFP_REQUIRE(count_bytes && len_bytes && sensor_bytes && onset_bytes, "null argument");
    *count_bytes = 4ll * n_streams;
    *len_bytes = 4ll * n_streams * SL_GMAX;
    *sensor_bytes = 4ll * n_streams * SL_GMAX * SL_LEN;
    *onset_bytes = 8ll * n_streams * SL_GMAX * SL_LEN;
    return OFP_OK;
}

extern "C" int ofp_stream_locate(const double *sensor_xyz_dev, int32_t n_sensors, const float *lag_maps_dev,
                                 int32_t map_size, const float *max_lags_dev, const float *min_lags_dev,
                                 const float *max_max_dev, double radius_cm, double samples_per_cm, double sr,
                                 double c_cm_s, int32_t n_streams, int32_t n_channels, const int32_t *det_channel_dev,
                                 const int32_t *det_delta_dev, const int32_t *det_count_dev, int64_t current_index,
                                 int32_t *state_count_dev, int32_t *state_len_dev, int32_t *state_sensor_dev,
                                 int64_t *state_onset_dev, double *xy_dev, int32_t *found_dev, void *stream) {
    OFP_REQUIRE(sensor_xyz_dev && lag_maps_dev && max_lags_dev && min_lags_dev && max_max_dev && det_channel_dev &&
                    det_delta_dev && det_count_dev && state_count_dev && state_len_dev && state_sensor_dev &&
                    state_onset_dev && xy_dev && found_dev, "null argument");
    OFP_REQUIRE(n_sensors >= 3 && n_channels >= 1 && n_channels <= 32, "bad sensor / channel count");
    if (n_streams == 0) return OFP_OK;
    SlArgs a;
    a.geo.locs = sensor_xyz_dev; a.geo.maps = lag_maps_dev; a.geo.max_lags = max_lags_dev;
    a.geo.min_lags = min_lags_dev; a.geo.max_max = max_max_dev; a.geo.S = n_sensors; a.geo.Hm = map_size;
    a.geo.H = 0; a.geo.n_per_hit = 3; a.geo.radius = radius_cm; a.geo.samples_per_cm = samples_per_cm;
    a.geo.sr = sr; a.geo.c_cm = c_cm_s; a.geo.sensors = nullptr; a.geo.onsets = nullptr; a.geo.onset_stride = 0;
    a.geo.xy = nullptr; a.geo.status = nullptr;
    a.n_streams = n_streams; a.C = n_channels; a.det_ch = det_channel_dev; a.det_delta = det_delta_dev;
    a.det_cnt = det_count_dev; a.current_index = current_index; a.g_count = state_count_dev;
    a.g_len = state_len_dev; a.g_sensor = state_sensor_dev; a.g_onset = state_onset_dev; a.xy = xy_dev;
    a.found = found_dev;
    k5_stream_locate<false><<<(n_streams + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(a);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

__global__ void k_advance_index(int64_t *idx, int64_t by) { *idx += by; }

// ofp_stream_locate with the block's start index read from device memory and advanced by `advance` samples
// afterwards: the form a CUDA graph can replay (kernel arguments are baked into a captured graph).
extern "C" int ofp_stream_locate_dev(const double *sensor_xyz_dev, int32_t n_sensors, const float *lag_maps_dev,
                                     int32_t map_size, const float *max_lags_dev, const float *min_lags_dev,
                                     const float *max_max_dev, double radius_cm, double samples_per_cm, double sr,
                                     double c_cm_s, int32_t n_streams, int32_t n_channels,
                                     const int32_t *det_channel_dev, const int32_t *det_delta_dev,
                                     const int32_t *det_count_dev, int64_t *current_index_dev, int32_t advance,
                                     int32_t *state_count_dev, int32_t *state_len_dev, int32_t *state_sensor_dev,
                                     int64_t *state_onset_dev, double *xy_dev, int32_t *found_dev, void *stream) {
    OFP_REQUIRE(sensor_xyz_dev && lag_maps_dev && max_lags_dev && min_lags_dev && max_max_dev && det_channel_dev &&
                    det_delta_dev && det_count_dev && state_count_dev && state_len_dev && state_sensor_dev &&
                    state_onset_dev && xy_dev && found_dev && current_index_dev, "null argument");
    OFP_REQUIRE(n_sensors >= 3 && n_channels >= 1 && n_channels <= 32, "bad sensor / channel count");
    if (n_streams == 0) return OFP_OK;
    SlArgs a;
    a.geo.locs = sensor_xyz_dev; a.geo.maps = lag_maps_dev; a.geo.max_lags = max_lags_dev;
    a.geo.min_lags = min_lags_dev; a.geo.max_max = max_max_dev; a.geo.S = n_sensors; a.geo.Hm = map_size;
    a.geo.H = 0; a.geo.n_per_hit = 3; a.geo.radius = radius_cm; a.geo.samples_per_cm = samples_per_cm;
    a.geo.sr = sr; a.geo.c_cm = c_cm_s; a.geo.sensors = nullptr; a.geo.onsets = nullptr; a.geo.onset_stride = 0;
    a.geo.xy = nullptr; a.geo.status = nullptr;
    a.n_streams = n_streams; a.C = n_channels; a.det_ch = det_channel_dev; a.det_delta = det_delta_dev;
    a.det_cnt = det_count_dev; a.current_index = 0; a.current_index_dev = current_index_dev;
    a.g_count = state_count_dev; a.g_len = state_len_dev; a.g_sensor = state_sensor_dev; a.g_onset = state_onset_dev;
    a.xy = xy_dev; a.found = found_dev;
    k5_stream_locate<false><<<(n_streams + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(a);
    k_advance_index<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(current_index_dev, advance);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

// The same with the ring-buffer refinement of every new pair (multilateration.py:457-501, what PlayRec's callback
// runs: realtime/audio.py:69 passes self.rec_audio): ring_dev [S, ring_rows, C] holds the most recent rows of every
// stream INCLUDING the current block (ofp_ring_write first).  One warp per stream.
extern "C" int ofp_stream_locate_ring_dev(const double *sensor_xyz_dev, int32_t n_sensors, const float *lag_maps_dev,
                                          int32_t map_size, const float *max_lags_dev, const float *min_lags_dev,
                                          const float *max_max_dev, double radius_cm, double samples_per_cm,
                                          double sr, double c_cm_s, int32_t n_streams, int32_t n_channels,
                                          const int32_t *det_channel_dev, const int32_t *det_delta_dev,
                                          const int32_t *det_count_dev, int64_t *current_index_dev, int32_t advance,
                                          const float *ring_dev, int32_t ring_rows, int32_t block_size,
                                          int32_t onset_tolerance, int32_t normalization_cutoff,
                                          int32_t *state_count_dev, int32_t *state_len_dev,
                                          int32_t *state_sensor_dev, int64_t *state_onset_dev, double *xy_dev,
                                          int32_t *found_dev, void *stream) {
    OFP_REQUIRE(sensor_xyz_dev && lag_maps_dev && max_lags_dev && min_lags_dev && max_max_dev && det_channel_dev &&
                    det_delta_dev && det_count_dev && state_count_dev && state_len_dev && state_sensor_dev &&
                    state_onset_dev && xy_dev && found_dev && current_index_dev && ring_dev, "null argument");
    OFP_REQUIRE(n_sensors >= 3 && n_channels >= 1 && n_channels <= 32, "bad sensor / channel count");
    OFP_REQUIRE(ring_rows >= block_size && block_size >= 1 && onset_tolerance >= 1 && normalization_cutoff >= 1,
                "bad ring / refinement parameters");
    if (n_streams == 0) return OFP_OK;
    SlArgs a;
    a.geo.locs = sensor_xyz_dev; a.geo.maps = lag_maps_dev; a.geo.max_lags = max_lags_dev;
    a.geo.min_lags = min_lags_dev; a.geo.max_max = max_max_dev; a.geo.S = n_sensors; a.geo.Hm = map_size;
    a.geo.H = 0; a.geo.n_per_hit = 3; a.geo.radius = radius_cm; a.geo.samples_per_cm = samples_per_cm;
    a.geo.sr = sr; a.geo.c_cm = c_cm_s; a.geo.sensors = nullptr; a.geo.onsets = nullptr; a.geo.onset_stride = 0;
    a.geo.xy = nullptr; a.geo.status = nullptr;
    a.n_streams = n_streams; a.C = n_channels; a.det_ch = det_channel_dev; a.det_delta = det_delta_dev;
    a.det_cnt = det_count_dev; a.current_index = 0; a.current_index_dev = current_index_dev;
    a.g_count = state_count_dev; a.g_len = state_len_dev; a.g_sensor = state_sensor_dev; a.g_onset = state_onset_dev;
    a.xy = xy_dev; a.found = found_dev;
    a.ring = ring_dev; a.ring_rows = ring_rows; a.block = block_size; a.tol = onset_tolerance; a.cutoff = normalization_cutoff;
    constexpr int WARPS = 4;
    const size_t smem = static_cast<size_t>(WARPS) * 2 * SL_LMAX * sizeof(double);
    // (set on every launch: the attribute belongs to the function ON THE CURRENT DEVICE, and a process may drive several)
    OFP_CUDA_CHECK(cudaFuncSetAttribute(k5_stream_locate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    k5_stream_locate<true><<<(n_streams + WARPS - 1) / WARPS, 32 * WARPS, smem, static_cast<cudaStream_t>(stream)>>>(a);
    k_advance_index<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(current_index_dev, advance);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

extern "C" int ofp_ring_write(float *ring_dev, int32_t ring_rows, const float *blocks_dev, int64_t stream_stride,
                              int32_t n_streams, int32_t block_size, int32_t n_channels,
                              const int64_t *current_index_dev, void *stream) {
    OFP_REQUIRE(ring_dev && blocks_dev && current_index_dev, "null argument");
    OFP_REQUIRE(ring_rows >= block_size && block_size >= 1 && n_channels >= 1, "bad ring shape");
    if (n_streams == 0) return OFP_OK;
    if (stream_stride <= 0) stream_stride = static_cast<int64_t>(block_size) * n_channels;
    const int64_t total = static_cast<int64_t>(n_streams) * block_size * n_channels;
    const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 8));
    k_ring_write<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(ring_dev, ring_rows, blocks_dev, stream_stride,
                                                                    n_streams, block_size, n_channels,
                                                                    current_index_dev, 0);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

extern "C" int ofp_solve_trilateration(const double *problems_dev, const double *seeds_dev, int32_t n_problems,
                                       double xtol, int32_t maxfev, double *xy_dev, int32_t *ier_dev,
                                       int32_t *nfev_dev, void *stream) {
    if (n_problems == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(problems_dev && seeds_dev && xy_dev && ier_dev, "null argument");
    k5_solve<<<(n_problems + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        problems_dev, seeds_dev, n_problems, xtol, maxfev, xy_dev, ier_dev, nfev_dev);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

extern "C" int ofp_locate_hits(const double *sensor_xyz_dev, int32_t n_sensors, const float *lag_maps_dev,
                               int32_t map_size, const float *max_lags_dev, const float *min_lags_dev,
                               const float *max_max_dev, double radius_cm, double samples_per_cm, double sr,
                               double c_cm_s, const int32_t *hit_sensors_dev, const int32_t *hit_onsets_dev,
                               int32_t onset_stride, int32_t n_hits, double *xy_dev, int32_t *status_dev,
                               void *stream) {
    if (n_hits == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(sensor_xyz_dev && lag_maps_dev && max_lags_dev && min_lags_dev && max_max_dev && hit_onsets_dev &&
                    xy_dev && status_dev, "null argument");
    OFP_REQUIRE(n_sensors >= 3 && onset_stride >= 3, "need at least three sensors / onsets per hit");
    K5Args a;
    a.locs = sensor_xyz_dev; a.maps = lag_maps_dev; a.max_lags = max_lags_dev; a.min_lags = min_lags_dev;
    a.max_max = max_max_dev; a.S = n_sensors; a.Hm = map_size; a.H = n_hits; a.n_per_hit = 3;
    a.radius = radius_cm; a.samples_per_cm = samples_per_cm; a.sr = sr; a.c_cm = c_cm_s;
    a.sensors = hit_sensors_dev; a.onsets = hit_onsets_dev; a.onset_stride = onset_stride;
    a.xy = xy_dev; a.status = status_dev;
    k5_locate<<<(n_hits + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

// ---------------------------------------------------------------------------------------------
// Model bypass of trilaterate (multilateration.py:553-557): res = model.call_np((d_a1, d_b1)) * 100 with
// model = calibration.FCNN (calibration.py:463-560) in eval mode: [Linear -> BatchNorm1d (running
// statistics) -> activation] x n_hidden -> Linear.  One thread per row, activations in registers,
// parameters in shared memory.  BatchNorm is applied as the per-feature affine (scale, shift) it is
// at inference; packed per layer: W [out][in], b [out], scale [out], shift [out].
// ---------------------------------------------------------------------------------------------
namespace ofp {
constexpr int FC_MAXW = 32;   // widest layer
constexpr int FC_MAXL = 8;    // layers incl. the output layer
struct FcArgs {
    const float *x;  // [n, in]
    int64_t n;
    int32_t n_layers, act;
    int32_t width[FC_MAXL + 1];  // in, hidden..., out
    int32_t off[FC_MAXL];        // float offset of layer l's block
    int32_t n_params;
    const float *params;
    const int32_t *status;       // optional: rows with status != 0 are skipped
    float out_scale;
    float *out_f32;              // [n, out] or null
    double *out_f64;             // [n, out] or null (rows skipped keep their content)
};

__global__ void __launch_bounds__(128) k5_fcnn(const FcArgs a) {
    extern __shared__ float fc_prm[];
    for (int i = threadIdx.x; i < a.n_params; i += blockDim.x) fc_prm[i] = a.params[i];
    __syncthreads();
    const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    if (a.status != nullptr && a.status[r] != 0) return;
    float h[FC_MAXW], g[FC_MAXW];
    const int win = a.width[0];
    for (int i = 0; i < FC_MAXW; ++i) h[i] = i < win ? a.x[r * win + i] : 0.f;
    for (int l = 0; l < a.n_layers; ++l) {
        const int ni = a.width[l], no = a.width[l + 1];
        const float *W = fc_prm + a.off[l], *b = W + no * ni, *sc = b + no, *sh = sc + no;
        const bool last = l == a.n_layers - 1;
#pragma unroll
        for (int o = 0; o < FC_MAXW; ++o) {
            float v = 0.f;
            if (o < no) {
#pragma unroll
                for (int i = 0; i < FC_MAXW; ++i)
                    if (i < ni) v = fmaf(W[o * ni + i], h[i], v);
                v += b[o];
                if (!last) {
                    v = fmaf(v, sc[o], sh[o]);
                    if (a.act == 0) v = fmaxf(v, 0.f);
                    else if (a.act == 1) v = tanhf(v);
                    else if (a.act == 2) v = 1.0f / (1.0f + expf(-v));
                    else if (a.act == 3) v = v / (1.0f + expf(-v));
                }
            }
            g[o] = v;
        }
#pragma unroll
        for (int o = 0; o < FC_MAXW; ++o) h[o] = g[o];
    }
    const int nout = a.width[a.n_layers];
#pragma unroll
    for (int o = 0; o < FC_MAXW; ++o) {
        if (o < nout) {
            if (a.out_f32) a.out_f32[r * nout + o] = h[o] * a.out_scale;
            if (a.out_f64) a.out_f64[r * nout + o] = static_cast<double>(h[o] * a.out_scale);
        }
    }
}
}  // namespace ofp

extern "C" int ofp_fcnn_forward(const float *x_dev, int64_t n_rows, int32_t n_layers, const int32_t *widths_host,
                                int32_t activation, const float *params_dev, const int32_t *status_dev,
                                float out_scale, float *out_f32_dev, double *out_f64_dev, void *stream) {
    if (n_rows == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(x_dev && widths_host && params_dev && (out_f32_dev || out_f64_dev), "null argument");
    OFP_REQUIRE(n_layers >= 1 && n_layers <= FC_MAXL, "1..%d layers supported", FC_MAXL);
    OFP_REQUIRE(activation >= 0 && activation <= 4, "activation: 0 ReLU, 1 tanh, 2 sigmoid, 3 SiLU, 4 identity");
    FcArgs a{};
    a.x = x_dev; a.n = n_rows; a.n_layers = n_layers; a.act = activation; a.params = params_dev;
    a.status = status_dev; a.out_scale = out_scale; a.out_f32 = out_f32_dev; a.out_f64 = out_f64_dev;
    int off = 0;
    for (int l = 0; l <= n_layers; ++l) {
        OFP_REQUIRE(widths_host[l] >= 1 && widths_host[l] <= FC_MAXW, "layer widths 1..%d supported", FC_MAXW);
        a.width[l] = widths_host[l];
        if (l < n_layers) { a.off[l] = off; off += widths_host[l + 1] * widths_host[l] + 3 * widths_host[l + 1]; }
    }
    a.n_params = off;
    k5_fcnn<<<static_cast<unsigned>((n_rows + 127) / 128), 128, off * sizeof(float), static_cast<cudaStream_t>(stream)>>>(a);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

extern "C" int ofp_locate_hits_lags(const double *sensor_xyz_dev, int32_t n_sensors, const float *lag_maps_dev,
                                    int32_t map_size, const float *max_lags_dev, const float *min_lags_dev,
                                    const float *max_max_dev, double radius_cm, double samples_per_cm, double sr,
                                    double c_cm_s, const int32_t *hit_sensors_dev, const int32_t *hit_onsets_dev,
                                    int32_t onset_stride, int32_t n_hits, float *pair_lags_dev, double *xy_dev,
                                    int32_t *status_dev, void *stream) {
    if (n_hits == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(sensor_xyz_dev && lag_maps_dev && max_lags_dev && min_lags_dev && max_max_dev && hit_onsets_dev &&
                    xy_dev && status_dev && pair_lags_dev, "null argument");
    OFP_REQUIRE(n_sensors >= 3 && onset_stride >= 3, "need at least three sensors / onsets per hit");
    K5Args a;
    a.locs = sensor_xyz_dev; a.maps = lag_maps_dev; a.max_lags = max_lags_dev; a.min_lags = min_lags_dev;
    a.max_max = max_max_dev; a.S = n_sensors; a.Hm = map_size; a.H = n_hits; a.n_per_hit = 3;
    a.radius = radius_cm; a.samples_per_cm = samples_per_cm; a.sr = sr; a.c_cm = c_cm_s;
    a.sensors = hit_sensors_dev; a.onsets = hit_onsets_dev; a.onset_stride = onset_stride;
    a.xy = xy_dev; a.status = status_dev; a.pair_lags = pair_lags_dev;
    k5_locate<<<(n_hits + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}
