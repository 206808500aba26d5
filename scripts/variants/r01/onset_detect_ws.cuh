// K1, warp-specialised form.  Included by onset_detect.cu (uses its helpers).
//
// Time cannot be split bit-exactly (non-linear recurrences), and only R*C lanes exist, so a single
// warp per recording group is bound by dependent-instruction latency (ncu: 1.7 warps per scheduler,
// 43 % issue utilisation).  The per-sample work is therefore cut into three pipeline stages that run
// as three warps of one CTA and hand samples over through shared-memory rings guarded by mbarriers:
//
//   warp A  TMA tile -> high-pass (recurrence) -> dB (pointwise, correctly rounded log10)  -> db ring
//   warp B  db ring  -> fast/slow followers (recurrences) -> 10**x (pointwise)             -> rel ring
//   warp C  rel ring -> min/max trackers (recurrences), block FSM, onset compaction, rel -> HBM
//
// Each lane keeps the same (recording, channel) in all three warps; ring rows are [sample][lane], so
// every shared-memory access of a warp hits 30 consecutive words (conflict free).  The rel ring holds
// one whole block plus one chunk: the thresholds of a block depend on its END-of-block min/max
// (SURVEY Q4), so warp C re-scans the block in the ring when a crossing is possible.
#pragma once

namespace ofp {

constexpr int WS_THREADS = 96;
constexpr int WS_MAX_NDB = 8, WS_MAX_NRS = 40;
// shared-memory header: mbarriers
constexpr int WS_OFF_XFULL = 0;                               // up to 4
constexpr int WS_OFF_DBFULL = 64;                             // WS_MAX_NDB
constexpr int WS_OFF_DBEMPTY = WS_OFF_DBFULL + 8 * WS_MAX_NDB;
constexpr int WS_OFF_RELFULL = WS_OFF_DBEMPTY + 8 * WS_MAX_NDB;
constexpr int WS_OFF_RELEMPTY = WS_OFF_RELFULL + 8 * WS_MAX_NRS;
constexpr int WS_OFF_TABLES = 1024;
constexpr int WS_OFF_DATA = WS_OFF_TABLES + (2 << OFP_LOG_N) * 8 + 32 * 8;  // 3328, 128-aligned

struct WsCfg {
    int32_t CH, NDB, NRS, row;      // chunk length, db ring chunks, rel ring chunks, floats per ring row
    uint32_t off_x, off_db, off_rel;  // byte offsets of the x stages / db ring / rel ring
};

__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(200000u)  // suspend-time hint (ns): sleep instead of spinning
            : "memory");
    } while (!done);
}

template <bool USE_HP>
__global__ void __launch_bounds__(WS_THREADS, 7) k1_detect_ws(const __grid_constant__ CUtensorMap tmap, const K1Args a,
                                                           const WsCfg w) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = a.p.n_channels, B = a.p.block_size, G = a.G, T = a.T, TC = a.TC;
    const int CH = w.CH, NDB = w.NDB, NRS = w.NRS;
    const int g_raw = lane / C;
    const bool in_group = g_raw < G;
    const int g = in_group ? g_raw : 0;
    const int c = in_group ? lane - g_raw * C : 0;
    const int li = g * C + c;  // column of this lane in the ring rows
    const int rec0 = blockIdx.x * G;
    const int rec = rec0 + g;
    const bool active = in_group && rec < a.R;
    const int64_t lid = static_cast<int64_t>(rec) * C + c;
    const uint32_t rowb = 4u * w.row;

    // ---- set-up: tables, barriers, constants ----
    {
        double *logtab = reinterpret_cast<double *>(smem + WS_OFF_TABLES);
        for (int i = threadIdx.x; i < (2 << OFP_LOG_N); i += WS_THREADS) logtab[i] = g_logtab[i];
        if (threadIdx.x < 32) logtab[(2 << OFP_LOG_N) + threadIdx.x] = g_exptab[threadIdx.x];
        if (threadIdx.x == 0) {
            for (int s = 0; s < a.nst; ++s) mbar_init(reinterpret_cast<uint64_t *>(smem + WS_OFF_XFULL) + s, 1);
            for (int s = 0; s < NDB; ++s) {
                mbar_init(reinterpret_cast<uint64_t *>(smem + WS_OFF_DBFULL) + s, 32);
                mbar_init(reinterpret_cast<uint64_t *>(smem + WS_OFF_DBEMPTY) + s, 32);
            }
            for (int s = 0; s < NRS; ++s) {
                mbar_init(reinterpret_cast<uint64_t *>(smem + WS_OFF_RELFULL) + s, 32);
                mbar_init(reinterpret_cast<uint64_t *>(smem + WS_OFF_RELEMPTY) + s, 32);
            }
            fence_mbar_init();
            tma_prefetch_desc(&tmap);
        }
    }
    Coef kf = load_coef(a);
    MathConst mc = math_const();
    launder(kf, mc, sbase + w.off_rel + 512u * warp);  // rel ring is idle until warp B produces
    __syncthreads();

    const uint32_t logtab_s = sbase + WS_OFF_TABLES, exptab_s = logtab_s + (2u << OFP_LOG_N) * 8u;
    const int64_t env_warm = (a.warm_n / B) * B;
    const int64_t Qw = env_warm / CH, Q = Qw + a.n_main / CH;

    if (warp == 0) {
        // =========================== warp A: TMA -> high-pass -> dB ===========================
        float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;
        if (active) { z0 = a.st.z0[lid]; z1 = a.st.z1[lid]; z2 = a.st.z2[lid]; z3 = a.st.z3[lid]; }
        Lane L; L.z0 = z0; L.z1 = z1; L.z2 = z2; L.z3 = z3;
        const uint32_t box_bytes = static_cast<uint32_t>(G) * TC * 4u;
        const uint32_t xfull = sbase + WS_OFF_XFULL;
        const uint32_t xlane = sbase + w.off_x + 4u * (g * TC + c);
        const uint32_t step = 4u * C;
        int xs = 0;                // x stage of the tile being consumed
        uint32_t xpar = 0;         // parity of xfull[xs]
        int kc = 0, slot = 0;      // position inside the chunk being produced, its ring slot
        uint32_t fix_mask = 0;     // samples of the current chunk whose dB needs the exact slow path
        uint32_t par = 1;          // parity to wait for on dbempty[slot] (fresh barrier passes parity 1)
        for (int phase = 0; phase < 2; ++phase) {
            const int64_t len = phase == 0 ? a.warm_n : a.n_main;
            if (len <= 0) continue;
            const int64_t env_len = phase == 0 ? env_warm : len;
            const int64_t ntiles = (len + T - 1) / T;
            if (lane == 0) {
                for (int p = 0; p < a.nst - 1 && p < ntiles; ++p) {
                    const int s = (xs + p) % a.nst;
                    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + WS_OFF_XFULL) + s;
                    mbar_expect_tx(bar, box_bytes);
                    tma_load_2d(smem + w.off_x + static_cast<size_t>(s) * a.stage_floats * 4, &tmap, bar, p * TC, rec0);
                }
            }
            for (int64_t ti = 0; ti < ntiles; ++ti) {
                const int s = xs;
                const int64_t nx = ti + a.nst - 1;
                if (lane == 0 && nx < ntiles) {
                    const int sn = xs == 0 ? a.nst - 1 : xs - 1;  // (xs + nst - 1) % nst
                    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + WS_OFF_XFULL) + sn;
                    mbar_expect_tx(bar, box_bytes);
                    tma_load_2d(smem + w.off_x + static_cast<size_t>(sn) * a.stage_floats * 4, &tmap, bar,
                                static_cast<int32_t>(nx * TC), rec0);
                }
                mbar_wait_s(xfull + 8u * s, xpar);
                if (++xs == a.nst) { xs = 0; xpar ^= 1u; }
                const uint32_t sp = xlane + 4u * static_cast<uint32_t>(s) * a.stage_floats;
                const int64_t t0 = ti * T;
                const int tl = static_cast<int>(min(static_cast<int64_t>(T), len - t0));
                int j = 0;
                // 4-sample units never straddle a chunk (4 | CH, env_len a multiple of CH)
                for (; j + 4 <= tl && t0 + j < env_len; j += 4) {
                    if (kc == 0) mbar_wait_s(sbase + WS_OFF_DBEMPTY + 8u * slot, par);
                    float h[4], db[4], aux[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float x = lds_f32(sp + (j + u) * step);
                        h[u] = USE_HP ? hp_step(L, kf, x) : x;
                    }
                    const uint32_t dst = sbase + w.off_db + (static_cast<uint32_t>(slot * CH + kc)) * rowb + 4u * li;
                    uint32_t redo4 = 0;
                    to_db_vec<4>(h, kf.floor_db, logtab_s, mc, db, aux, redo4, 0);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        // a flagged sample leaves |h + 1e-10| in the ring; it is fixed up when the chunk closes
                        if ((redo4 >> u) & 1u) db[u] = aux[u];
                        if (in_group) sts_f32(dst + u * rowb, db[u]);
                    }
                    fix_mask |= redo4 << kc;
                    kc += 4;
                    if (kc == CH) {
                        if (__any_sync(0xffffffffu, fix_mask != 0)) {  // rare: exact log10 out of line
                            const uint32_t base = sbase + w.off_db + static_cast<uint32_t>(slot * CH) * rowb + 4u * li;
                            while (fix_mask) {
                                const int k = __ffs(fix_mask) - 1;
                                fix_mask &= fix_mask - 1;
                                const float v = lds_f32(base + k * rowb);
                                if (in_group) sts_f32(base + k * rowb, db_of(slow_log10(v), kf.floor_db));
                            }
                            __syncwarp();
                        }
                        mbar_arrive_s(sbase + WS_OFF_DBFULL + 8u * slot);
                        kc = 0;
                        if (++slot == NDB) { slot = 0; par ^= 1u; }
                    }
                }
                // warm-up tail beyond the last full block: only the high-pass advances
                if (USE_HP)
                    for (; j < tl; ++j) hp_step(L, kf, lds_f32(sp + j * step));
                __syncwarp();
            }
        }
        if (active) { a.st.z0[lid] = L.z0; a.st.z1[lid] = L.z1; a.st.z2[lid] = L.z2; a.st.z3[lid] = L.z3; }
    } else if (warp == 1) {
        // =========================== warp B: followers -> 10**x ===========================
        float yf = kf.floor_db, ys = kf.floor_db;
        if (active) { yf = a.st.yf[lid]; ys = a.st.ys[lid]; }
        int sd = 0, sr = 0;
        uint32_t pd = 0, pr = 1;  // parities: dbfull[sd] (consumer), relempty[sr] (producer)
        for (int64_t q = 0; q < Q; ++q) {
            mbar_wait_s(sbase + WS_OFF_DBFULL + 8u * sd, pd);
            mbar_wait_s(sbase + WS_OFF_RELEMPTY + 8u * sr, pr);
            const uint32_t src = sbase + w.off_db + static_cast<uint32_t>(sd * CH) * rowb + 4u * li;
            const uint32_t dst = sbase + w.off_rel + static_cast<uint32_t>(sr * CH) * rowb + 4u * li;
            // followers: straight-line float32 steps; the (|t| tiny) sliver where the reference's double
            // add rounds differently is only FLAGGED here and the chunk is redone exactly if any lane hit it
            const float yf0 = yf, ys0 = ys;
            float dr[16];
            bool sliver = false;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (i < CH) {
                    const float db = lds_f32(src + i * rowb);
                    const float t1 = __fsub_rn(db, yf), t2 = __fsub_rn(db, ys);
                    sliver |= (fabsf(t1) < 0x1p-22f && t1 != 0.0f) | (fabsf(t2) < 0x1p-22f && t2 != 0.0f);
                    const float d1 = __fadd_rn(t1, 1e-10f), d2 = __fadd_rn(t2, 1e-10f);
                    yf = __fadd_rn(yf, __fmul_rn(d1 > 0.0f ? kf.fa : kf.fr, d1));
                    ys = __fadd_rn(ys, __fmul_rn(d2 > 0.0f ? kf.sa : kf.sr, d2));
                    dr[i] = __fsub_rn(yf, ys);
                }
            }
            if (__any_sync(0xffffffffu, sliver)) {
                yf = yf0; ys = ys0;
                for (int i = 0; i < CH; ++i) {
                    const float db = lds_f32(src + i * rowb);
                    yf = ar_step(yf, db, kf.fa, kf.fr);
                    ys = ar_step(ys, db, kf.sa, kf.sr);
                    const float v = __fsub_rn(yf, ys);
#pragma unroll
                    for (int k = 0; k < 16; ++k) if (k == i) dr[k] = v;
                }
            }
            uint32_t fix = 0;
#pragma unroll
            for (int i0 = 0; i0 < 16; i0 += 4) {
                if (i0 < CH) {
                    float d4[4] = {dr[i0], dr[i0 + 1], dr[i0 + 2], dr[i0 + 3]}, a4[4];
                    to_amp_vec<4>(d4, kf.ceil_amp, exptab_s, mc, a4, fix, i0);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        // a flagged sample leaves the follower difference in the ring; fixed up below
                        const float v = (fix >> (i0 + u)) & 1u ? d4[u] : a4[u];
                        if (in_group) sts_f32(dst + (i0 + u) * rowb, v);
                    }
                }
            }
            if (__any_sync(0xffffffffu, fix != 0)) {  // rare: exact 10**x out of line
                while (fix) {
                    const int k = __ffs(fix) - 1;
                    fix &= fix - 1;
                    const float r = lds_f32(dst + k * rowb);
                    if (in_group) sts_f32(dst + k * rowb, amp_of(slow_exp10(__fdiv_rn(r, 20.0f)), kf.ceil_amp));
                }
                __syncwarp();
            }
            mbar_arrive_s(sbase + WS_OFF_DBEMPTY + 8u * sd);
            mbar_arrive_s(sbase + WS_OFF_RELFULL + 8u * sr);
            if (++sd == NDB) { sd = 0; pd ^= 1u; }
            if (++sr == NRS) { sr = 0; pr ^= 1u; }
        }
        if (active) { a.st.yf[lid] = yf; a.st.ys[lid] = ys; }
    } else {
        // =========================== warp C: min/max, FSM, outputs ===========================
        Lane L;
        L.mn = 0.f; L.mx = 10.f; L.prev = 0.f; L.state = 0; L.deb = 0;
        if (active) { L.mn = a.st.mn[lid]; L.mx = a.st.mx[lid]; L.prev = a.st.prev[lid]; L.state = a.st.state[lid]; L.deb = a.st.deb[lid]; }
        L.bmax = -INFINITY; L.bmin = INFINITY;
        const unsigned rec_mask = (C == 32 ? 0xffffffffu : ((1u << C) - 1u)) << (g * C);
        const unsigned lower_mask = rec_mask & ((1u << lane) - 1u);
        const int cpb = B / CH;  // chunks per block
        int32_t cnt = 0;
        int64_t blk = 0;
        int kpos = 0;
        float *relp = a.rel ? a.rel + static_cast<int64_t>(rec) * a.rel_stride + c : nullptr;
        const bool wr = a.rel != nullptr && active;
        const uint32_t relcol = sbase + w.off_rel + 4u * li;
        const int lgch = CH == 16 ? 4 : (CH == 8 ? 3 : 2);
        int sr = 0, sr0 = 0;  // current chunk slot; slot of the first chunk of the current block
        uint32_t pr = 0;
        for (int64_t q = 0; q < Q; ++q) {
            const bool main_phase = q >= Qw;
            const bool do_minmax = !main_phase || !a.p.manual;
            mbar_wait_s(sbase + WS_OFF_RELFULL + 8u * sr, pr);
            const int sr_cur = sr;
            if (++sr == NRS) { sr = 0; pr ^= 1u; }
            const uint32_t src = relcol + static_cast<uint32_t>(sr_cur * CH) * rowb;
            for (int i = 0; i < CH; i += 4) {
                float r[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) r[u] = lds_f32(src + (i + u) * rowb);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (do_minmax) minmax_step(L, kf, r[u]);
                    L.bmax = fmaxf(L.bmax, r[u]);
                    L.bmin = fminf(L.bmin, r[u]);
                }
                if (main_phase && wr) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) relp[static_cast<int64_t>(i + u) * C] = r[u];
                }
            }
            if (main_phase && relp) relp += static_cast<int64_t>(CH) * C;
            kpos += CH;
            if (!main_phase) {
                mbar_arrive_s(sbase + WS_OFF_RELEMPTY + 8u * sr_cur);
                if (kpos == B) { kpos = 0; L.bmax = -INFINITY; L.bmin = INFINITY; sr0 = sr; }
                continue;
            }
            if (kpos < B) continue;
            kpos = 0;
            // ---- block FSM, detection.py:759-792; the block sits in ring chunks q-cpb+1 .. q ----
            auto rel_at = [&](int k) -> float {
                int slot = sr0 + (k >> lgch);
                if (slot >= NRS) slot -= NRS;
                return lds_f32(relcol + static_cast<uint32_t>((slot << lgch) + (k & (CH - 1))) * rowb);
            };
            const float last = rel_at(B - 1);
            const float thr_on = a.p.manual ? a.p.on_thr : __fadd_rn(__fmul_rn(L.mx, a.p.on_thr), L.mn);
            const float thr_off = a.p.manual ? a.p.off_thr : __fadd_rn(__fmul_rn(L.mx, a.p.off_thr), L.mn);
            int oi = 0;
            bool hit = false;
            if (!L.state && L.deb < 1 && L.bmax > thr_on) {
                float before = L.prev;
                for (int k = 0; k < B; ++k) {
                    const float r = rel_at(k);
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
                    if (rel_at(k) < thr_off) { off = true; break; }
            }
            if (off) L.state = 0;
            L.prev = last;
            if (hits) {
                const int pos = cnt + __popc(hits & lower_mask);
                if (hit && active && pos < a.cap) {
                    a.on_ch[static_cast<int64_t>(rec) * a.cap + pos] = c;
                    a.on_idx[static_cast<int64_t>(rec) * a.cap + pos] = static_cast<int32_t>(blk * B + oi);
                }
                cnt += __popc(hits & rec_mask);
            }
            ++blk;
            L.bmax = -INFINITY; L.bmin = INFINITY;
            __syncwarp();
            for (int k = 0, sl = sr0; k < cpb; ++k) {
                mbar_arrive_s(sbase + WS_OFF_RELEMPTY + 8u * static_cast<uint32_t>(sl));
                if (++sl == NRS) sl = 0;
            }
            sr0 = sr;
        }
        if (active) {
            a.st.mn[lid] = L.mn; a.st.mx[lid] = L.mx; a.st.prev[lid] = L.prev;
            a.st.state[lid] = L.state; a.st.deb[lid] = L.deb;
            if (c == 0 && a.on_cnt != nullptr) a.on_cnt[rec] = cnt;
        }
    }
}

}  // namespace ofp
