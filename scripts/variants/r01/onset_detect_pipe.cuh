// K1, software-pipelined form (included by onset_detect.cu).
//
// The detector is five stages per sample: high-pass (serial) -> dB (pointwise) -> followers (serial)
// -> 10**x (pointwise) -> min/max trackers (serial).  A lane cannot split time, and there are only
// one or two warps per scheduler, so the latency of the three serial recurrences has to be hidden
// inside the warp.  The loop below works on chunks of 8 samples and skews the five stages by one
// chunk each: iteration i runs
//     S1 high-pass of chunk i      P1 dB of chunk i-1      S2 followers of chunk i-2
//     P2 10**x of chunk i-3        S3 min/max of chunk i-4
// which are mutually independent, so ptxas can fill the dependent-issue gaps of each recurrence with
// the other four strands (one straight-line block of ~900 instructions per iteration).  The values
// in flight between strands live in registers (h, db, dr, amp: 4 x 8 floats).
//
// Exactness is unchanged from k1_detect: the pointwise strands use the table-driven double
// evaluation with a rounding test and the followers the float32 shortcut; whatever they flag
// (k1_math.cuh, ar_step) is redone exactly *for that strand only* after a warp vote -- pointwise
// results are recomputed from their inputs, the followers from the state saved at the top of the
// iteration.  The block FSM runs four chunks behind the loads, on a ring of B + 8 envelope rows.
#pragma once

namespace ofp {

constexpr int PU = 8;       // samples per chunk
constexpr int PIPE_LAG = 4; // S3 runs this many chunks behind S1

struct P1Regs {  // dB front end of PU samples, stage-major (to_db_fast op for op)
    uint32_t ix[PU], tmp[PU];
    int32_t kexp[PU];
    double invc[PU], logc[PU], r[PU], r2[PU], p01[PU], p23[PU], pp[PU], base[PU];
};
struct P2Regs {  // 10**x of PU samples, stage-major (to_amp_fast op for op)
    float q[PU];
    int32_t ki[PU];
    double t[PU], kd[PU], rr[PU], r2[PU], p01[PU], p23[PU], pp[PU], sc[PU];
};

template <int S>
__device__ __forceinline__ void p1_stage(const float (&h)[PU], P1Regs &R, float floor_db, uint32_t logtab,
                                         const MathConst &mc, float (&db)[PU], bool &flag) {
#pragma unroll
    for (int u = 0; u < PU; ++u) {
        if (S == 0) R.ix[u] = __float_as_uint(fabsf(__fadd_rn(h[u], 1e-10f)));
        if (S == 1) R.tmp[u] = R.ix[u] - OFP_LOG_OFF;
        if (S == 2)
            lds_2f64(logtab + ((R.tmp[u] >> (23 - OFP_LOG_N - 4)) & (((1u << OFP_LOG_N) - 1u) << 4)), R.invc[u],
                     R.logc[u]);
        if (S == 3) R.kexp[u] = static_cast<int32_t>(R.tmp[u]) >> 23;
        if (S == 4) {
            const uint32_t iz = R.ix[u] - (R.tmp[u] & 0xff800000u);
            const double z = __hiloint2double(static_cast<int>((iz >> 3) + 0x38000000u), static_cast<int>(iz << 29));
            R.r[u] = __fma_rn(z, R.invc[u], -1.0);
        }
        if (S == 5) R.base[u] = __fma_rn(static_cast<double>(R.kexp[u]), mc.log10_2, R.logc[u]);
        if (S == 6) R.r2[u] = __dmul_rn(R.r[u], R.r[u]);
        if (S == 7) R.p01[u] = __fma_rn(R.r[u], mc.a2, mc.a1);
        if (S == 8) R.p23[u] = __fma_rn(R.r[u], mc.a4, mc.a3);
        if (S == 9) R.pp[u] = __fma_rn(R.r2[u], mc.a5, R.p23[u]);
        if (S == 10) R.pp[u] = __fma_rn(R.r2[u], R.pp[u], R.p01[u]);
        if (S == 11) R.base[u] = __fma_rn(R.r[u], R.pp[u], R.base[u]);  // log10(v)
        if (S == 12)
            flag |= ((R.ix[u] - 0x00800000u) >= 0x7f000000u) | near_f32_midpoint(R.base[u], 1u << 13);
        if (S == 13) db[u] = db_of(R.base[u], floor_db);
    }
}
constexpr int P1_STAGES = 14;

template <int S>
__device__ __forceinline__ void p2_stage(const float (&dr)[PU], P2Regs &R, float ceil_amp, uint32_t exptab,
                                         const MathConst &mc, float (&amp)[PU], bool &flag) {
#pragma unroll
    for (int u = 0; u < PU; ++u) {
        if (S == 0) R.q[u] = __fmul_rn(dr[u], 0.05f);
        if (S == 1) R.q[u] = __fmaf_rn(__fmaf_rn(-20.0f, R.q[u], dr[u]), 0.05f, R.q[u]);  // dr / 20, exact
        if (S == 2) R.t[u] = __dmul_rn(static_cast<double>(R.q[u]), mc.log2_10);
        if (S == 3) R.kd[u] = __fma_rn(R.t[u], 32.0, mc.shift);
        if (S == 4) { R.ki[u] = __double2loint(R.kd[u]); R.kd[u] = __dsub_rn(R.kd[u], mc.shift); }
        if (S == 5) R.sc[u] = lds_f64(exptab + ((R.ki[u] & 31) << 3));
        if (S == 6) R.rr[u] = __fma_rn(R.kd[u], -0.03125, R.t[u]);
        if (S == 7) R.r2[u] = __dmul_rn(R.rr[u], R.rr[u]);
        if (S == 8) R.p01[u] = __fma_rn(R.rr[u], mc.e2, mc.e1);
        if (S == 9) R.p23[u] = __fma_rn(R.rr[u], mc.e4, mc.e3);
        if (S == 10) R.pp[u] = __fma_rn(R.r2[u], mc.e5, R.p23[u]);
        if (S == 11) R.pp[u] = __fma_rn(R.r2[u], R.pp[u], R.p01[u]);
        if (S == 12) R.t[u] = __fma_rn(__dmul_rn(R.sc[u], R.rr[u]), R.pp[u], R.sc[u]);
        if (S == 13)
            R.t[u] = __hiloint2double(__double2hiint(R.t[u]) + ((R.ki[u] >> 5) << 20), __double2loint(R.t[u]));
        if (S == 14) flag |= !(fabsf(R.q[u]) < 30.0f) | near_f32_midpoint(R.t[u], 1u << 8);
        if (S == 15) amp[u] = amp_of(R.t[u], ceil_amp);
    }
}
constexpr int P2_STAGES = 16;

struct PipeVecs {
    float h[PU], db[PU], dr[PU], amp[PU];
};

// One slot = one sample of each serial strand plus a slice of the pointwise strands.
template <int U, bool USE_HP, bool HP_SYM, bool DO_MM>
__device__ __forceinline__ void pipe_slot(Lane &L, const Coef &k, const MathConst &mc, const float (&x)[PU],
                                          uint32_t logtab, uint32_t exptab, const PipeVecs &in, PipeVecs &out,
                                          P1Regs &A, P2Regs &E, bool &f1, bool &f2, bool &f3) {
    // pointwise slices: P1 occupies slots 0-3, P2 slots 4-7 (their temporaries are never live together)
    if (U < 4) {
        p1_stage<4 * U + 0>(in.h, A, k.floor_db, logtab, mc, out.db, f1);
        p1_stage<4 * U + 1>(in.h, A, k.floor_db, logtab, mc, out.db, f1);
    } else {
        p2_stage<4 * (U - 4) + 0>(in.dr, E, k.ceil_amp, exptab, mc, out.amp, f3);
        p2_stage<4 * (U - 4) + 1>(in.dr, E, k.ceil_amp, exptab, mc, out.amp, f3);
    }
    // S1: high-pass of sample U of the newest chunk
    out.h[U] = USE_HP ? (HP_SYM ? hp_step_sym(L, k, x[U]) : hp_step(L, k, x[U])) : x[U];
    // S2: followers (float32 shortcut of ar_step; the sliver where it is not proven exact is flagged
    // conservatively on the operands, see chunk_fast)
    {
        const float d = in.db[U];
        const float t1 = __fsub_rn(d, L.yf), t2 = __fsub_rn(d, L.ys);
        f2 |= (fabsf(d) < 2.0f) | (fabsf(L.yf) < 2.0f) | (fabsf(L.ys) < 2.0f);
        const float d1 = __fadd_rn(t1, 1e-10f), d2 = __fadd_rn(t2, 1e-10f);
        L.yf = __fadd_rn(L.yf, __fmul_rn(d1 > 0.0f ? k.fa : k.fr, d1));
        L.ys = __fadd_rn(L.ys, __fmul_rn(d2 > 0.0f ? k.sa : k.sr, d2));
        out.dr[U] = __fsub_rn(L.yf, L.ys);
    }
    if (U < 4) {
        if (4 * U + 2 < P1_STAGES) p1_stage<4 * U + 2>(in.h, A, k.floor_db, logtab, mc, out.db, f1);
        if (4 * U + 3 < P1_STAGES) p1_stage<4 * U + 3>(in.h, A, k.floor_db, logtab, mc, out.db, f1);
    } else {
        p2_stage<4 * (U - 4) + 2>(in.dr, E, k.ceil_amp, exptab, mc, out.amp, f3);
        p2_stage<4 * (U - 4) + 3>(in.dr, E, k.ceil_amp, exptab, mc, out.amp, f3);
    }
    // S3: min/max trackers and block extrema of the oldest chunk
    {
        const float r = in.amp[U];
        if (DO_MM) minmax_step(L, k, r);
        L.bmax = fmaxf(L.bmax, r);
        L.bmin = fminf(L.bmin, r);
    }
}

template <bool USE_HP, bool HP_SYM, bool DO_MM>
__device__ __forceinline__ uint32_t pipe_body(Lane &L, const Coef &k, const MathConst &mc, uint32_t xs, uint32_t step,
                                              uint32_t logtab, uint32_t exptab, const PipeVecs &in, PipeVecs &out) {
    float x[PU];
#pragma unroll
    for (int u = 0; u < PU; ++u) x[u] = lds_f32(xs + u * step);
    P1Regs A;
    P2Regs E;
    bool f1 = false, f2 = false, f3 = false;
    pipe_slot<0, USE_HP, HP_SYM, DO_MM>(L, k, mc, x, logtab, exptab, in, out, A, E, f1, f2, f3);
    pipe_slot<1, USE_HP, HP_SYM, DO_MM>(L, k, mc, x, logtab, exptab, in, out, A, E, f1, f2, f3);
    pipe_slot<2, USE_HP, HP_SYM, DO_MM>(L, k, mc, x, logtab, exptab, in, out, A, E, f1, f2, f3);
    pipe_slot<3, USE_HP, HP_SYM, DO_MM>(L, k, mc, x, logtab, exptab, in, out, A, E, f1, f2, f3);
    pipe_slot<4, USE_HP, HP_SYM, DO_MM>(L, k, mc, x, logtab, exptab, in, out, A, E, f1, f2, f3);
    pipe_slot<5, USE_HP, HP_SYM, DO_MM>(L, k, mc, x, logtab, exptab, in, out, A, E, f1, f2, f3);
    pipe_slot<6, USE_HP, HP_SYM, DO_MM>(L, k, mc, x, logtab, exptab, in, out, A, E, f1, f2, f3);
    pipe_slot<7, USE_HP, HP_SYM, DO_MM>(L, k, mc, x, logtab, exptab, in, out, A, E, f1, f2, f3);
    return (f1 ? 1u : 0u) | (f2 ? 2u : 0u) | (f3 ? 4u : 0u);
}

// Exact recomputation of whatever the fast strands flagged (rare; after a warp vote).  Fully unrolled
// with constant indices so that the in-flight vectors stay in registers.
__device__ __forceinline__ void pipe_fix_db(const float (&h)[PU], float (&db)[PU], float floor_db, uint32_t logtab,
                                            const MathConst &mc) {
#pragma unroll
    for (int u = 0; u < PU; ++u) {
        float v;
        bool redo;
        float d = to_db_fast(h[u], floor_db, logtab, mc, v, redo);
        if (redo) d = db_of(slow_log10(v), floor_db);
        db[u] = d;
    }
}
__device__ __forceinline__ void pipe_fix_amp(const float (&dr)[PU], float (&amp)[PU], float ceil_amp, uint32_t exptab,
                                             const MathConst &mc) {
#pragma unroll
    for (int u = 0; u < PU; ++u) {
        float q;
        bool redo;
        float am = to_amp_fast(dr[u], ceil_amp, exptab, mc, q, redo);
        if (redo) {
            const float qq = fabsf(q) < 30.0f ? q : __fdiv_rn(dr[u], 20.0f);
            am = amp_of(slow_exp10(qq), ceil_amp);
        }
        amp[u] = am;
    }
}

template <bool USE_HP, bool HP_SYM, bool MANUAL>
__global__ void __launch_bounds__(32, 1) k1_pipe(const __grid_constant__ CUtensorMap tmap, const K1Args a, const int NR) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    double *logtab = reinterpret_cast<double *>(smem + 128);
    double *exptab = logtab + (2 << OFP_LOG_N);
    float *stages = reinterpret_cast<float *>(smem + K1_SMEM_HEADER);
    float *relbuf = stages + static_cast<size_t>(a.nst) * a.stage_floats;

    const int lane = threadIdx.x;
    const int C = a.p.n_channels, B = a.p.block_size, G = a.G, T = a.T, TC = a.TC;
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
    const uint32_t logtab_s = smem_u32(logtab), exptab_s = smem_u32(exptab);
    const uint32_t step = 4u * C;
    MathConst mc = math_const();
    launder(kf, mc, smem_u32(relbuf));
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

    for (int i = lane; i < (2 << OFP_LOG_N); i += 32) logtab[i] = g_logtab[i];
    exptab[lane] = g_exptab[lane];
    __syncwarp();
    if (lane == 0) {
        for (int s = 0; s < a.nst; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmap);
    }
    __syncwarp();
    const uint32_t box_bytes = static_cast<uint32_t>(G) * TC * 4u;
    uint32_t it = 0;  // tiles consumed so far (ring position / parity)
    int32_t cnt = 0;  // onsets emitted for this lane's recording
    int64_t blk = 0;  // main-phase block index

    float *rcol = relbuf + g * a.stride_rel + c;  // this lane's column of the envelope ring
    const uint32_t rcol_s = smem_u32(rcol);
    const uint32_t stage0_s = smem_u32(stages) + 4u * (g * TC + c);
    const int CPB = B / PU;  // chunks per block

    for (int phase = 0; phase < 2; ++phase) {
        const int64_t len = phase == 0 ? a.warm_n : a.n_main;
        if (len <= 0) continue;
        const int64_t env_len = phase == 0 ? (len / B) * B : len;
        const int64_t nch = env_len / PU;
        const int64_t ntiles = (len + T - 1) / T;
        if (lane == 0) {
            for (int p = 0; p < a.nst - 1 && p < ntiles; ++p) {
                const int s = (it + p) % a.nst;
                mbar_expect_tx(&bars[s], box_bytes);
                tma_load_2d(stages + static_cast<size_t>(s) * a.stage_floats, &tmap, &bars[s], p * TC, rec0);
            }
        }
        // input tile walk: tile `tidx` of this phase is entered lazily, left when all T samples are read
        int64_t tidx = 0;
        int tpos = 0;
        bool entered = false;
        uint32_t sp = stage0_s;
        auto enter_tile = [&]() {
            const int64_t nx = tidx + a.nst - 1;
            if (lane == 0 && nx < ntiles) {
                const int sn = (it + a.nst - 1) % a.nst;
                mbar_expect_tx(&bars[sn], box_bytes);
                tma_load_2d(stages + static_cast<size_t>(sn) * a.stage_floats, &tmap, &bars[sn],
                            static_cast<int32_t>(nx * TC), rec0);
            }
            const int s = it % a.nst;
            mbar_wait(&bars[s], (it / a.nst) & 1u);
            sp = stage0_s + 4u * static_cast<uint32_t>(s) * a.stage_floats;
            entered = true;
        };
        auto leave_tile = [&]() {
            __syncwarp();  // every lane is done with this stage before the producer refills it
            ++it; ++tidx; tpos = 0; entered = false;
        };

        PipeVecs v;
#pragma unroll
        for (int u = 0; u < PU; ++u) { v.h[u] = 0.f; v.db[u] = a.p.floor_db; v.dr[u] = 0.f; v.amp[u] = 0.f; }
        int wrow = 0;   // ring row of the next chunk P2 stores
        int brow = 0;   // ring row where the block S3 is working on starts
        int s3pos = 0;  // chunks of that block S3 has consumed
        const int64_t total = nch > 0 ? nch + PIPE_LAG : 0;
        for (int64_t i = 0; i < total; ++i) {
            uint32_t xs = sp;
            if (i < nch) {
                if (!entered) enter_tile();
                xs = sp + static_cast<uint32_t>(tpos) * step;
            }
            const Lane sv = L;
            PipeVecs w;
            uint32_t flags;
            if (MANUAL && phase == 1)
                flags = pipe_body<USE_HP, HP_SYM, false>(L, kf, mc, xs, step, logtab_s, exptab_s, v, w);
            else
                flags = pipe_body<USE_HP, HP_SYM, true>(L, kf, mc, xs, step, logtab_s, exptab_s, v, w);
            if (__any_sync(0xffffffffu, flags != 0)) {
                if (__any_sync(0xffffffffu, flags & 1u)) pipe_fix_db(v.h, w.db, kf.floor_db, logtab_s, mc);
                if (__any_sync(0xffffffffu, flags & 2u)) {
                    L.yf = sv.yf; L.ys = sv.ys;
#pragma unroll
                    for (int u = 0; u < PU; ++u) {
                        L.yf = ar_step(L.yf, v.db[u], kf.fa, kf.fr);
                        L.ys = ar_step(L.ys, v.db[u], kf.sa, kf.sr);
                        w.dr[u] = __fsub_rn(L.yf, L.ys);
                    }
                }
                if (__any_sync(0xffffffffu, flags & 4u)) pipe_fix_amp(v.dr, w.amp, kf.ceil_amp, exptab_s, mc);
            }
            // strands working on chunks outside [0, nch) ran on filler: undo their state changes
            if (i >= nch) { L.z0 = sv.z0; L.z1 = sv.z1; L.z2 = sv.z2; L.z3 = sv.z3; }
            else {
                tpos += PU;
                if (tpos == T) leave_tile();
            }
            if (i < 2 || i >= nch + 2) { L.yf = sv.yf; L.ys = sv.ys; }
            if (i >= 3 && i < nch + 3) {  // P2 finished chunk i-3: into the envelope ring
                if (in_group) {
                    const uint32_t rp = rcol_s + static_cast<uint32_t>(wrow) * step;
#pragma unroll
                    for (int u = 0; u < PU; ++u) sts_f32(rp + u * step, w.amp[u]);
                }
                wrow += PU;
                if (wrow == NR) wrow = 0;
            }
            if (i < PIPE_LAG) { L.mn = sv.mn; L.mx = sv.mx; L.bmax = sv.bmax; L.bmin = sv.bmin; }
            else if (++s3pos == CPB) {  // S3 consumed the last chunk of a block
                if (phase == 1) {
                    __syncwarp();
                    block_end(L, a, rcol, relbuf, lane, g, c, rec, rec0, active, rec_mask, lower_mask, cnt, blk, brow, NR);
                    ++blk;
                }
                brow += B;
                if (brow >= NR) brow -= NR;
                s3pos = 0;
                L.bmax = -INFINITY; L.bmin = INFINITY;
            }
            v = w;
        }
        // warm-up tail beyond the last full block: only the high-pass advances
        // (detection.py:828-829 filters the whole half second in one call)
        int64_t rest = len - env_len;
        while (rest > 0) {
            if (!entered) enter_tile();
            const int n = static_cast<int>(min(static_cast<int64_t>(T - tpos), rest));
            if (USE_HP) {
                for (int j = 0; j < n; ++j) hp_step(L, kf, lds_f32(sp + static_cast<uint32_t>(tpos + j) * step));
            }
            tpos += n;
            rest -= n;
            if (tpos == T) leave_tile();
        }
        if (entered) leave_tile();  // last tile of the phase only partly used
    }

    if (active) {
        a.st.z0[lid] = L.z0; a.st.z1[lid] = L.z1; a.st.z2[lid] = L.z2; a.st.z3[lid] = L.z3;
        a.st.yf[lid] = L.yf; a.st.ys[lid] = L.ys; a.st.mn[lid] = L.mn; a.st.mx[lid] = L.mx;
        a.st.prev[lid] = L.prev; a.st.state[lid] = L.state; a.st.deb[lid] = L.deb;
        if (c == 0 && a.on_cnt != nullptr) a.on_cnt[rec] = cnt;
    }
}

}  // namespace ofp
