"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

TEST INFRASTRUCTURE.  The reference (timlod/onset-fingerprinting) has no tests and ships no
data, so the only way to pin the oracle is to run the reference itself on seeded synthetic
input (onset_fingerprinting_b200/synth.py) and commit what it returned.  The reference is
imported from /root/reference through oracle/ref_harness.py (librosa / loopmate stubbed, DLL
compiled by `make -C oracle ref`).  Inputs are NOT stored: tests regenerate them from the same
seeds (a sha1 of the input is stored to detect generator drift).

Recorded environment: see 'env' in each file (numpy/scipy versions, numpy SIMD dispatch --
float32 log10/power results depend on it).
"""
from __future__ import annotations

import hashlib
import io
import contextlib
import sys
from pathlib import Path

import numpy as np
import scipy

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_harness as rh  # noqa: E402
from onset_fingerprinting_b200 import synth  # noqa: E402

OUT = ROOT / "tests" / "golden"

DETECT_CASES = {
    # name: (synth kwargs, detect kwargs)
    "default3": (dict(seconds=3.0, seed=1), dict()),
    "realtime3": (dict(seconds=2.5, seed=2),
                  dict(hipass_freq=0, fast_ar=(0.3, 800), slow_ar=(8000, 8000), on_threshold=0.45,
                       off_threshold=0.45)),
    "manual_b100": (dict(seconds=2.0, seed=3), dict(block_size=100, on_threshold=1.8, off_threshold=1.2)),
    "b256_partial": (dict(seconds=2.0, seed=4), dict(block_size=256, hipass_freq=1000.0, cooldown=4000)),
    "mesh16": (dict(seconds=1.5, seed=5, sensors=synth.SENSORS_16MESH, medium="drumhead"), dict(block_size=64)),
}

FIX_OPTS = {
    "defaults": dict(),
    "d1_abs": dict(d=1, take_abs=True),
    "up_zero": dict(onset_direction="up", zero_left=True, normalization_cutoff=20, onset_tolerance=108,
                    filter_size=7, shift_onsets=40),
    "d1_down": dict(d=1, onset_direction="down", onset_tolerance=50),
    "abs150": dict(take_abs=True, onset_tolerance=150, d=1, filter_size=7),
}


def sha(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def env():
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        np.show_config()
    simd = [ln.strip() for ln in buf.getvalue().splitlines() if "AVX" in ln or "SSE" in ln]
    return f"numpy {np.__version__}; scipy {scipy.__version__}; simd {' '.join(simd)[:400]}"


def gen_detect(det):
    for name, (skw, dkw) in DETECT_CASES.items():
        x, hits = synth.drum_recording(**skw)
        ch, on, rel = det.detect_onsets_amplitude(x, sr=96000, **dkw)
        np.savez_compressed(
            OUT / f"detect_{name}.npz", channels=np.asarray(ch, np.int32), onsets=np.asarray(on, np.int64),
            rel_sub=rel[::16].astype(np.float32), rel_shape=np.asarray(rel.shape), x_sha=sha(x), env=env(),
            arrival=hits["arrival"])
        print(name, x.shape, len(ch), "onsets")


def gen_backtrack(det):
    """backtrack=True, offline and block by block (RingStub zero-initialised stands in for loopmate)."""
    x, _ = synth.drum_recording(seconds=2.0, seed=8)
    out = {"x_sha": sha(x), "env": env()}
    for tag, kw in (("b128", dict(backtrack_buffer_size=128, backtrack_smooth_size=5)),
                    ("b256s1", dict(backtrack_buffer_size=256, backtrack_smooth_size=1))):
        ch, on, _ = det.detect_onsets_amplitude(x, sr=96000, backtrack=True, **kw)
        out[f"ch_{tag}"] = np.asarray(ch, np.int32)
        out[f"on_{tag}"] = np.asarray(on, np.int64)
    ch0, on0, _ = det.detect_onsets_amplitude(x, sr=96000)
    out["on_plain"] = np.asarray(on0, np.int64)
    np.savez_compressed(OUT / "backtrack.npz", **out)
    print("backtrack moved", int((out["on_b128"] != out["on_plain"]).sum()), "of", len(on0))


def gen_stream(det):
    """AmplitudeOnsetDetector block by block (realtime/audio.py:39-52 settings), no warm-up."""
    x, _ = synth.drum_recording(seconds=1.5, seed=7, first_hit=20000)
    od = det.AmplitudeOnsetDetector(3, 128, hipass_freq=0, fast_ar=(0.3, 800), slow_ar=(8000, 8000),
                                    on_threshold=0.45, off_threshold=0.45, cooldown=1323, sr=96000)
    blocks, chans, deltas, rels = [], [], [], []
    for b, i in enumerate(range(0, len(x) - 127, 128)):
        c, d, r = od(x[i:i + 128])
        if b % 8 == 0:
            rels.append(r[::16].copy())
        for cc, dd in zip(c, d):
            blocks.append(b); chans.append(cc); deltas.append(dd)
    np.savez_compressed(OUT / "stream_realtime.npz", blocks=np.asarray(blocks, np.int32),
                        channels=np.asarray(chans, np.int32), deltas=np.asarray(deltas, np.int32),
                        rel_sub=np.asarray(rels, np.float32), x_sha=sha(x), env=env())
    print("stream", len(blocks), "onsets")


def gen_kernels(det):
    """Known answers for the two DLL kernels, find_onset_groups and cross_correlation_lag."""
    rng = np.random.default_rng(11)
    xb = (rng.uniform(-70, 0, (4, 64, 5))).astype(np.float32)
    ar = det.AREnvelopeFollower(np.full((64, 5), -70, np.float32), 3, 383)
    ar_out = np.stack([ar(b).copy() for b in xb])
    mm = det.MinMaxEnvelopeFollower(np.array([[0, 10]] * 5).T, alpha_min=1e-4, alpha_max=1e-5, minmin=2)
    xm = rng.uniform(0, 30, (4, 64, 5)).astype(np.float32)
    xm[1, :10] = 1.0
    mm_out = np.stack([np.stack([v.copy() for v in mm(b)]) for b in xm])
    bw = det.ButterworthFilter(2000, 3, 4, 96000, "high")
    xf = rng.standard_normal((3, 500, 3)).astype(np.float32)
    hp_out = np.stack([bw(b) for b in xf])
    # cross_correlation_lag on random signals / options
    cc_in, cc_out = [], []
    for t in range(400):
        n = int(rng.integers(60, 400))
        a = rng.standard_normal(n).astype(np.float32)
        b = np.roll(a, int(rng.integers(-40, 40))) + 0.3 * rng.standard_normal(n).astype(np.float32)
        tol = int(rng.choice([30, 50, 64, 108]))
        cut = int(rng.choice([10, 20]))
        o0 = int(rng.integers(0, n))
        o1 = int(rng.integers(0, n))
        d = int(rng.integers(0, 2))
        ab = bool(rng.integers(0, 2))
        use_legal = bool(t % 5 == 0)
        l0, l1 = sorted(rng.integers(-20, 80, 2).tolist())
        r = det.cross_correlation_lag(a, b, onsets=None if use_legal else (o0, o1),
                                      legal_lags=(l0, l1) if use_legal else None, d=d,
                                      normalization_cutoff=cut, onset_tolerance=tol, take_abs=ab)
        cc_in.append((t, n, tol, cut, o0, o1, d, int(ab), int(use_legal), l0, l1))
        cc_out.append(-(2**31) if r is None else int(r))
    # find_onset_groups
    g1 = det.find_onset_groups([100, 130, 90, 5000, 5010, 5020, 9000], [0, 1, 2, 0, 1, 2, 0], 1000, 3)
    g2 = det.find_onset_groups([100, 130, 90, 5000, 5010, 5020, 9000, 9010, 9020], [0, 1, 2, 0, 1, 2, 2, 1, 0],
                               1000, 3, close_channel=2)
    np.savez_compressed(OUT / "kernels.npz", ar_in=xb, ar_out=ar_out, mm_in=xm, mm_out=mm_out, hp_in=xf,
                        hp_out=hp_out, hp_b=bw.b, hp_a=bw.a, cc_params=np.asarray(cc_in, np.int64),
                        cc_out=np.asarray(cc_out, np.int64), groups1=g1, groups2=g2, env=env())
    print("kernels ok; cc None count", sum(1 for v in cc_out if v == -(2**31)))


def gen_fix(det):
    for tag, skw, dkw in (("3ch", dict(seconds=3.0, seed=21), dict()),
                          ("16ch", dict(seconds=2.0, seed=22, sensors=synth.SENSORS_16MESH, medium="drumhead"),
                           dict())):
        x, _ = synth.drum_recording(**skw)
        ch, on, _ = det.detect_onsets_amplitude(x, sr=96000, **dkw)
        groups = det.find_onset_groups(on, ch, 1000 if tag == "3ch" else 1500, x.shape[1])
        out = {"groups": groups, "x_sha": sha(x), "env": env()}
        for name, kw in FIX_OPTS.items():
            fixed = np.zeros_like(groups)
            raised = np.zeros(len(groups), np.int32)
            for j in range(len(groups)):  # per group: the reference can raise
                try:
                    fixed[j] = det.fix_onsets(x, groups[j:j + 1], **kw)[0]
                except ValueError:
                    raised[j] = 1
                    fixed[j] = groups[j]
            out[f"fixed_{name}"] = fixed
            out[f"raised_{name}"] = raised
            print(tag, name, "moved", float((fixed != groups + kw.get("shift_onsets", 0)).mean()), "raised",
                  int(raised.sum()))
        np.savez_compressed(OUT / f"fix_{tag}.npz", **out)


def gen_locate(ml):
    rng = np.random.default_rng(31)
    for tag, sensors, medium in (("air3", synth.SENSORS_3MIC, "air"),
                                 ("drumhead3", [(0.9, 0, 0), (0.9, 120, 0), (0.5, 240, 0)], "drumhead")):
        m = ml.Multilaterate3D(sensors, sr=96000, medium=medium)
        locs = synth.sensor_xyz(sensors)
        c = synth.speed_cm_s(medium)
        n = 1500
        onsets = np.zeros((n, 3), np.int64)
        res = np.full((n, 2), np.nan)
        for h in range(n):
            rr = 0.95 * 17.78 * np.sqrt(rng.uniform())
            ang = rng.uniform(0, 2 * np.pi)
            p = np.array([rr * np.cos(ang), rr * np.sin(ang), 0.0])
            dist = np.sqrt(((locs - p) ** 2).sum(1))
            onsets[h] = 100000 * (h + 1) + np.round(dist / c * 96000) + rng.integers(-3, 4, 3)
            m.ongoing = []
            r = None
            with contextlib.redirect_stdout(io.StringIO()):
                for s in np.argsort(onsets[h], kind="stable"):
                    r = m.locate(int(s), int(onsets[h, s]))
            if r is not None:
                res[h] = r
        S = len(sensors)
        maps = np.full((S, S) + m.lag_maps[0][1].shape, np.nan, np.float32)
        mx = np.full((S, S), np.nan, np.float32)
        mn = np.full((S, S), np.nan, np.float32)
        for i in range(S):
            for j in range(S):
                if i != j:
                    maps[i, j] = m.lag_maps[i][j]; mx[i, j] = m.max_lags[i][j]; mn[i, j] = m.min_lags[i][j]
        np.savez_compressed(OUT / f"locate_{tag}.npz", onsets=onsets, xy=res, maps=maps, max_lags=mx, min_lags=mn,
                            max_max=np.asarray(m.max_max_lags, np.float32), c=m.c, radius=m.radius,
                            sensor_locs=np.asarray(m.sensor_locs), sensors=np.asarray(sensors, np.float64),
                            medium=medium, env=env())
        print(tag, "located", int(np.isfinite(res[:, 0]).sum()), "of", n)


def gen_online_cc():
    """c/test.py:5-46 workload (shortened): stream two noisy sines through online_cc."""
    occ = rh.load_online_cc()
    n, bs = 256, 64
    cc = occ.CrossCorrelation(n, bs)
    ns = n * 40
    rng = np.random.default_rng(0)
    t = np.linspace(0, 10 * ns / (n * 10000), ns)
    a = (np.sin(2 * np.pi * t * 300) + 0.01 * rng.random(ns)).astype(np.float32)
    b = (np.sin(2 * np.pi * t * 300 + 0.5) + 0.01 * rng.random(ns)).astype(np.float32)
    frames = []
    for k, i in enumerate(range(0, ns - bs + 1, bs)):
        out = cc.update(a[i:i + bs], b[i:i + bs])
        if k in (3, 10, 50, ns // bs - 1):
            frames.append(out.copy())
    np.savez_compressed(OUT / "online_cc.npz", a=a, b=b, frames=np.asarray(frames), frame_idx=np.asarray(
        [3, 10, 50, ns // bs - 1]), n=n, block=bs, env=env())
    print("online_cc ok")


def tools_inputs():
    """Seeded inputs of the helper-surface fixtures (shared by the generator and the tests)."""
    rng = np.random.default_rng(77)
    d = {}
    d["fd_x"] = rng.standard_normal((200, 3)).astype(np.float32)
    sig, ons = [], []
    for k in range(24):
        n = 700
        x = 0.01 * rng.standard_normal(n)
        o = int(rng.integers(20, 680)) if k % 6 else int(rng.choice([3, 695]))
        t = np.arange(n - o)
        x[o:] += rng.uniform(0.2, 1.0) * np.exp(-t / rng.uniform(30, 200)) * np.sin(2 * np.pi * t / rng.uniform(8, 40))
        sig.append(x.astype(np.float32)); ons.append(o + int(rng.integers(-10, 30)))
    d["or_x"], d["or_on"] = np.stack(sig), np.asarray(ons, np.int64)
    d["rel_x"] = rng.uniform(0, 3, (30, 400)).astype(np.float32)
    d["rel_y"] = rng.uniform(0, 3, (30, 400)).astype(np.float32)
    d["rel_on"] = np.stack([rng.integers(120, 200, 30), rng.integers(200, 280, 30)], 1).astype(np.int64)
    d["rel_lag"] = (d["rel_on"][:, 1] - d["rel_on"][:, 0] + rng.integers(-40, 40, 30)).astype(np.int64)
    d["bw_x"] = rng.standard_normal((3, 300, 4)).astype(np.float32)
    a = rng.standard_normal((40, 256)).astype(np.float32)
    sh = rng.integers(-60, 60, 40)
    d["fl_a"] = a
    d["fl_b"] = np.stack([np.roll(a[k], int(sh[k])) for k in range(40)]) + 0.2 * rng.standard_normal((40, 256)).astype(np.float32)
    d["fl_b"] = d["fl_b"].astype(np.float32)
    # 2-D trilateration problems: sensors on the rim, deltas from a true point plus noise, seed near it
    P = 200
    ang = rng.uniform(0, 2 * np.pi, (P, 3))
    sens = 16.0 * np.stack([np.cos(ang), np.sin(ang)], -1)           # [P, 3 sensors, 2]
    rr, th = 15.0 * np.sqrt(rng.uniform(size=P)), rng.uniform(0, 2 * np.pi, P)
    pt = np.stack([rr * np.cos(th), rr * np.sin(th)], -1)
    dist = np.sqrt(((sens - pt[:, None]) ** 2).sum(-1))
    d["tri_sens"] = sens
    d["tri_da"] = dist[:, 0] - dist[:, 2] + rng.normal(0, 0.05, P)
    d["tri_db"] = dist[:, 1] - dist[:, 2] + rng.normal(0, 0.05, P)
    d["tri_seed"] = pt + rng.normal(0, 2.0, (P, 2))
    d["init_x"], _ = synth.drum_recording(seconds=2.0, seed=9)
    return d


MESH3 = [(0.9, 0.0), (0.9, 120.0), (0.9, 240.0)]
MESH4 = [(0.9, 0.0), (0.9, 90.0), (0.9, 180.0), (0.9, 270.0)]


def mesh2d_onsets(sensors2d, n, seed, sr=96000):
    """Onset triples of n random strikes for planar sensors (drumhead medium)."""
    rng = np.random.default_rng(seed)
    r = 17.78
    locs = np.asarray([(s[0] * r * np.cos(np.radians(s[1])), s[0] * r * np.sin(np.radians(s[1]))) for s in sensors2d])
    rr, th = 0.9 * r * np.sqrt(rng.uniform(size=n)), rng.uniform(0, 2 * np.pi, n)
    pt = np.stack([rr * np.cos(th), rr * np.sin(th)], -1)
    dist = np.sqrt(((locs[None] - pt[:, None]) ** 2).sum(-1))
    return (100000 * (np.arange(n)[:, None] + 1) + np.round(dist / 8200.0 * sr) + rng.integers(-2, 3, dist.shape)).astype(np.int64)


def gen_tools(det, ml):
    """Helper surface of detection.py / multilateration.py."""
    d = tools_inputs()
    out = {"env": env()}
    for direction in ("up", "down"):
        out[f"fd_{direction}"] = det.filter_data(d["fd_x"].copy(), direction)
    out["or_a"] = np.asarray([det.detect_onset_region(x, int(o)) for x, o in zip(d["or_x"], d["or_on"])], np.int64)
    out["or_b"] = np.asarray([det.detect_onset_region(x, int(o), n=128, median_filter_size=7, threshold_factor=0.3)
                              for x, o in zip(d["or_x"], d["or_on"])], np.int64)
    out["rel_out"] = np.asarray([det.adjust_onset_rel(list(o), x, y, int(l)) for o, x, y, l in
                                 zip(d["rel_on"], d["rel_x"], d["rel_y"], d["rel_lag"])], np.int64)
    for tag, args in (("lo2", (3000, 4, 2, 96000, "low")), ("hi3", (500, 4, 3, 96000, "high"))):
        bw = det.ButterworthFilter(*args)
        out[f"bw_{tag}"] = np.stack([bw(b) for b in d["bw_x"]])
    out["fl"] = np.asarray([ml.find_lag(a, b) for a, b in zip(d["fl_a"], d["fl_b"])], np.int64)
    flm = [ml.find_lag_multi(a, b, 3) for a, b in zip(d["fl_a"][:12], d["fl_b"][:12])]
    out["flm_lags"] = np.asarray([np.pad(l, (0, 3 - len(l)), constant_values=-9999) for l, _ in flm], np.int64)
    out["flm_vals"] = np.asarray([np.pad(v, (0, 3 - len(v))) for _, v in flm], np.float64)
    tri = np.full((len(d["tri_da"]), 2), np.nan)
    for k in range(len(tri)):
        r = ml.solve_trilateration(tuple(d["tri_sens"][k, 0]), tuple(d["tri_sens"][k, 1]), tuple(d["tri_sens"][k, 2]),
                                   d["tri_da"][k], d["tri_db"][k], d["tri_seed"][k])
        if r is not None:
            tri[k] = r
    out["tri_xy"] = tri
    print("solve_trilateration converged", int(np.isfinite(tri[:, 0]).sum()), "of", len(tri))
    # Multilaterate (2-D) streaming locate
    m2 = ml.Multilaterate(MESH3, sr=96000, medium="drumhead")
    on = mesh2d_onsets(MESH3, 300, 41)
    res = np.full((len(on), 2), np.nan)
    for h in range(len(on)):
        m2.ongoing = []
        r = None
        for s in np.argsort(on[h], kind="stable"):
            r = m2.locate(int(s), int(on[h, s]))
        if r is not None:
            res[h] = r
    out["m2_onsets"], out["m2_res"] = on, res
    print("Multilaterate located", int(np.isfinite(res[:, 0]).sum()), "of", len(on))
    # MultilateratePaired
    mp = ml.MultilateratePaired(MESH4, scale=10, sr=96000)
    rng = np.random.default_rng(43)
    lags = rng.integers(-150, 150, (60, 2))
    ii = rng.integers(0, 4, 60)
    pres = np.full((60, 2), np.nan)
    for k in range(60):
        try:
            pres[k] = mp.locate([int(lags[k, 0]), int(lags[k, 1])], int(ii[k]))
        except TypeError:  # solve_trilateration returned None and the reference unpacks it (multilateration.py:828)
            pass
    out["mp_lags"], out["mp_i"], out["mp_res"] = lags, ii, pres
    xcc = rng.standard_normal((20, 600, 4)).astype(np.float32)
    for k in range(20):  # delayed copies so that the vote has a clear winner
        for c in range(1, 4):
            xcc[k, :, c] = np.roll(xcc[k, :, 0], int(rng.integers(-80, 80))) + 0.1 * xcc[k, :, c]
    out["mp_cc_x"] = xcc
    out["mp_cc"] = np.asarray([mp.locate_cc(xcc[k], 200, int(k % 4)) for k in range(20)], np.float64)
    lm, sa, sb = ml.lag_intensity_map((10.0, 5.0, 8.0), (-12.0, 3.0, 6.0), reflectivity=0.5, sr=96000)
    out["lim_lag"], out["lim_a"], out["lim_b"] = lm, sa, sb
    # AmplitudeOnsetDetector.init
    od = det.AmplitudeOnsetDetector(3, 128, sr=96000)
    with contextlib.redirect_stdout(io.StringIO()):
        od.init(d["init_x"])
    out["init_mins"], out["init_maxs"] = od.mins, od.maxs
    out["init_on"], out["init_off"], out["init_noise"] = od.on_threshold, od.off_threshold, od.noise_max
    np.savez_compressed(OUT / "tools.npz", **out)
    print("tools ok")


STREAM_CC = dict(seconds=2.5, seed=12)


def gen_stream_cc(det, ml):
    """PlayRec.detect_hits with the ring-buffer refinement (realtime/audio.py:62-74, 102;
    multilateration.py:457-501): detector blocks -> locate(sensor, onset, rec_audio)."""
    x, _ = synth.drum_recording(**STREAM_CC)
    od = det.AmplitudeOnsetDetector(3, 128, hipass_freq=0, fast_ar=(0.3, 800), slow_ar=(8000, 8000),
                                    on_threshold=0.45, off_threshold=0.45, cooldown=1323, sr=96000)
    m = ml.Multilaterate3D(synth.SENSORS_3MIC, sr=96000, medium="air")
    ring = rh.RingStub(np.zeros((4096, 3), np.float32))
    rows, cur = [], 0
    with contextlib.redirect_stdout(io.StringIO()):
        for b, i in enumerate(range(0, len(x) - 127, 128)):
            blk = x[i:i + 128]
            ring.write(blk)
            c, dl, _ = od(blk)
            if len(c) > 0:
                dd = [cur + int(v) for v in dl]
                for k in np.argsort(dd):
                    res = m.locate(int(c[k]), dd[k], ring)
                    rows.append((b, int(c[k]), dd[k], np.nan if res is None else res[0], np.nan if res is None else res[1]))
                    if res is not None:
                        break
            cur += 128
    rows = np.asarray(rows, np.float64)
    np.savez_compressed(OUT / "stream_cc.npz", rows=rows, x_sha=sha(x), env=env())
    print("stream_cc: detections", len(rows), "located", int(np.isfinite(rows[:, 3]).sum()))


def frames_inputs():
    rng = np.random.default_rng(55)
    x, _ = synth.drum_recording(seconds=1.0, seed=14)
    on = np.sort(rng.integers(2000, len(x) - 3000, 40))[:, None] + rng.integers(0, 60, (40, 3))
    return x, on.astype(np.int64)


def gen_frames():
    """data.py window extraction."""
    import torch

    data = rh.load_reference_data()
    x, on = frames_inputs()
    out = {"env": env(), "x_sha": sha(x)}
    out["fe_min"] = data.FrameExtractor(256, 16)(x, on)
    out["fe_each"] = data.FrameExtractor(128, 8, add_pre_samples=True, use_min_onset=False)(x, on)
    out["fe_1d"] = data.FrameExtractor(200, 0)(x[:, 1].copy(), on[:, 1])
    np.random.seed(5)
    out["fe_shift"] = data.FrameExtractor(256, 32, max_shift=10)(x, on)
    out["ffe_sha"] = sha(data.FastFrameExtractor(x, on, 256, 16)().numpy())
    a = torch.tensor(x[1000:1256].T.copy())
    b = torch.tensor(x[1010:1266].T.copy())
    out["bcc"] = data.batch_cc(a, b).numpy()
    pos = np.random.default_rng(1).uniform(-1, 1, (40, 2))
    ds = data.MCPOSD(x, on, pos, 256, 16)
    xx, yy = ds[0]
    out["ds_x_sha"], out["ds_y"] = sha(xx.numpy()), yy.numpy()
    np.savez_compressed(OUT / "frames.npz", **out)
    print("frames ok", out["fe_min"].shape, out["fe_each"].shape, out["bcc"].shape)


def stream_fsm_events(seed=91, n_streams=48, n_blocks=60):
    """Per-block detection lists for the group state machine: mostly plausible hits (three sensors
    within the legal lags), plus out-of-order onsets, repeated sensors, stray single detections and
    lags just outside the legal range -- everything locate()'s branches react to."""
    rng = np.random.default_rng(seed)
    locs = synth.sensor_xyz(synth.SENSORS_3MIC)
    c = synth.speed_cm_s("air")
    ev = np.full((n_streams, n_blocks, 3, 2), -1, np.int64)  # (channel, delta) per slot, -1 = none
    cnt = np.zeros((n_streams, n_blocks), np.int64)
    for s in range(n_streams):
        for b in range(n_blocks):
            kind = rng.uniform()
            dets = []
            if kind < 0.45:      # a hit: all three sensors in this block
                rr, ang = 0.9 * 17.78 * np.sqrt(rng.uniform()), rng.uniform(0, 2 * np.pi)
                p = np.array([rr * np.cos(ang), rr * np.sin(ang), 0.0])
                d = np.round(np.sqrt(((locs - p) ** 2).sum(1)) / c * 96000).astype(int)
                base = int(rng.integers(0, 128 - (d.max() - d.min()) - 1)) if d.max() - d.min() < 120 else 0
                dets = [(k, base + int(d[k] - d.min()) + int(rng.integers(-2, 3))) for k in range(3)]
                dets = [(k, min(max(v, 0), 127)) for k, v in dets]
                if rng.uniform() < 0.2:
                    dets = dets[: int(rng.integers(1, 3))]  # one sensor missed it
            elif kind < 0.6:     # stray detections
                for k in rng.permutation(3)[: int(rng.integers(1, 3))]:
                    dets.append((int(k), int(rng.integers(0, 128))))
            elif kind < 0.65:    # the same sensor twice cannot happen within a block; leave empty
                dets = []
            np.random.default_rng(seed + s * 1000 + b).shuffle(dets)  # K1 reports in channel order, keep both
            dets.sort(key=lambda t: t[0])
            for i, (k, v) in enumerate(dets):
                ev[s, b, i] = (k, v)
            cnt[s, b] = len(dets)
    return ev, cnt


def gen_stream_fsm(ml):
    """Multilaterate3D.locate fed block by block as PlayRec.detect_hits does (no rec_audio)."""
    ev, cnt = stream_fsm_events()
    S, NB = cnt.shape
    res = np.full((S, NB, 2), np.nan)
    glen = np.zeros((S, NB), np.int64)
    for s in range(S):
        m = ml.Multilaterate3D(synth.SENSORS_3MIC, sr=96000, medium="air")
        cur = 0
        with contextlib.redirect_stdout(io.StringIO()):
            for b in range(NB):
                k = int(cnt[s, b])
                c = ev[s, b, :k, 0]
                d = [cur + int(v) for v in ev[s, b, :k, 1]]
                if k:
                    for i in np.argsort(d):
                        r = m.locate(int(c[i]), d[i])
                        if r is not None:
                            res[s, b] = r
                            break
                glen[s, b] = len(m.ongoing)
                cur += 128
    np.savez_compressed(OUT / "stream_fsm.npz", res=res, n_groups=glen, env=env())
    print("stream_fsm: located", int(np.isfinite(res[..., 0]).sum()), "of", int((cnt == 3).sum()), "full blocks")


CONFIG0 = dict(seconds=60.0, seed=101)


def gen_config0(det, ml):
    """BASELINE.json configs[0] at its stated size: one 3-mic 96 kHz recording of 60 s (~500 hits, ~1500 onsets)
    through the whole reference chain: detect_onsets_amplitude -> find_onset_groups -> fix_onsets ->
    Multilaterate3D.locate (streaming, hit by hit)."""
    x, hits = synth.drum_recording(**CONFIG0)
    ch, on, rel = det.detect_onsets_amplitude(x, sr=96000)
    groups = det.find_onset_groups(on, ch, 1000, 3)
    fixed = det.fix_onsets(x, groups)
    m = ml.Multilaterate3D(synth.SENSORS_3MIC, sr=96000, medium="air")
    xy = np.full((len(fixed), 2), np.nan)
    with contextlib.redirect_stdout(io.StringIO()):
        for h, row in enumerate(fixed):
            m.ongoing = []
            r = None
            for s_ in np.argsort(row, kind="stable"):
                r = m.locate(int(s_), int(row[s_]))
            if r is not None:
                xy[h] = r
    np.savez_compressed(OUT / "config0_60s.npz", channels=np.asarray(ch, np.int32), onsets=np.asarray(on, np.int64),
                        rel_sub=rel[::64].astype(np.float32), rel_shape=np.asarray(rel.shape), groups=groups,
                        fixed=fixed, xy=xy, x_sha=sha(x), env=env(), arrival=hits["arrival"])
    print("config0: onsets", len(on), "groups", len(groups), "moved", int((fixed != groups).any(1).sum()), "located",
          int(np.isfinite(xy[:, 0]).sum()))


HITS16_OPTS = dict(filter_size=7, d=1, take_abs=True, normalization_cutoff=20, onset_tolerance=150)


def hits16_sections(n_hits=240, seed=61, L=768):
    """Host twin of bench.py's configs[2] input: one [L, 16] section per hit (a burst reaching the 16 mesh sensors
    between look+4 and look+4+420 samples) and detected onsets = true arrival + jitter."""
    rng = np.random.default_rng(seed)
    look = HITS16_OPTS["onset_tolerance"] + HITS16_OPTS["normalization_cutoff"]
    locs = synth.sensor_xyz(synth.SENSORS_16MESH)
    c = synth.speed_cm_s("drumhead")
    t = np.arange(4096) / 96000
    burst = np.exp(-400.0 * t) * np.sin(2 * np.pi * 900.0 * t)
    xs = (1e-4 * rng.standard_normal((n_hits, L, 16))).astype(np.float32)
    on = np.zeros((n_hits, 16), np.int64)
    for h in range(n_hits):
        rr, ang = 0.85 * 17.78 * np.sqrt(rng.uniform()), rng.uniform(0, 2 * np.pi)
        p = np.array([rr * np.cos(ang), rr * np.sin(ang), 0.0])
        dist = np.sqrt(((locs - p) ** 2).sum(1))
        delay = np.round(dist / c * 96000).astype(int)
        for k in range(16):
            a = look + 4 + delay[k]
            xs[h, a:, k] += (0.5 * 10.0 / dist[k] * burst[: L - a]).astype(np.float32)
            on[h, k] = min(max(a + int(rng.integers(-25, 40)), look + 1), L - look - 1)
    return xs, on


NO_SIMD_SORT = "AVX2 AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR"


def _hits16_reference(det):
    xs, on = hits16_sections()
    fixed = np.zeros_like(on)
    raised = np.zeros(len(on), np.int32)
    for h in range(len(on)):
        try:
            fixed[h] = det.fix_onsets(xs[h], on[h:h + 1], **HITS16_OPTS)[0]
        except ValueError:
            raised[h] = 1
            fixed[h] = on[h]
    return xs, on, fixed, raised


def gen_hits16(det):
    """configs[2] with the exact option set the bench runs (tol 150, cutoff 20, d=1, abs, median 7; 16 channels,
    section 768): the reference's fix_onsets per hit (it can raise).

    H8 (found with this fixture): fix_onsets orders the channels with np.argsort(og) (detection.py:416), and the
    order numpy gives EQUAL onsets depends on the build/CPU: the AVX2 / AVX-512 x86-simd-sort kernels are not
    stable, the scalar introsort is (insertion sort for n <= 16).  Because the reference onset moves between
    pairs (Q6), hits whose later channels hold tied onsets come out differently on the two.  Both are recorded:
    `fixed` = the reference as numpy runs it on this host, `fixed_scalar` = the same reference in a subprocess
    with NPY_DISABLE_CPU_FEATURES (scalar sort = stable order, which is what the oracle and the CUDA path pin)."""
    import os
    import subprocess
    import tempfile

    xs, on, fixed, raised = _hits16_reference(det)
    with tempfile.TemporaryDirectory() as td:
        tmp = os.path.join(td, "scalar.npz")
        envv = dict(os.environ, NPY_DISABLE_CPU_FEATURES=NO_SIMD_SORT)
        subprocess.run([sys.executable, "-m", "oracle.make_golden", "hits16_scalar", tmp], check=True, cwd=str(ROOT), env=envv)
        sc = np.load(tmp)
        fixed_scalar, raised_scalar = sc["fixed"], sc["raised"]
    srt = np.sort(on, axis=1)
    tied = (srt[:, 1:] == srt[:, :-1]).any(1)
    differ = (fixed != fixed_scalar).any(1)
    assert not (differ & ~tied).any(), "the two sort orders may only differ on hits with tied onsets"
    np.savez_compressed(OUT / "hits16_bench_opts.npz", fixed=fixed, raised=raised, fixed_scalar=fixed_scalar,
                        raised_scalar=raised_scalar, tied=tied, x_sha=sha(xs), on_sha=sha(on), env=env())
    print("hits16: moved", float((fixed != on).mean()), "raised", int(raised.sum()), "of", len(on), "| tied hits",
          int(tied.sum()), "of which the SIMD and scalar argsort orders give different results:", int(differ.sum()))
STREAM_RING = dict(n_streams=48, seconds=1.2, seed0=300, ring_rows=4096)


def stream_ring_input(s):
    """Recording of stream s of the ring-refinement fixture: hits every 0.118 s from a stream-dependent start."""
    return synth.drum_recording(seconds=STREAM_RING["seconds"], seed=STREAM_RING["seed0"] + s,
                                first_hit=6000 + 997 * (s % 11))[0]


def gen_stream_ring(det, ml):
    """PlayRec.detect_hits with rec_audio for 48 independent streams (realtime/audio.py:62-74, 102;
    multilateration.py:428-534 incl. the ring-buffer refinement 457-501), block by block: the golden for the
    batched device session (csrc/realtime.cu with ring_rows > 0).  Per stream: the blocks in which a hit was
    located and its position; `raised` = block at which the reference itself raised, -1 = never."""
    S = STREAM_RING["n_streams"]
    rows, raised, ndet = [], np.full(S, -1, np.int64), np.zeros(S, np.int64)
    for s in range(S):
        x = stream_ring_input(s)
        od = det.AmplitudeOnsetDetector(3, 128, hipass_freq=0, fast_ar=(0.3, 800), slow_ar=(8000, 8000),
                                        on_threshold=0.45, off_threshold=0.45, cooldown=1323, sr=96000)
        m = ml.Multilaterate3D(synth.SENSORS_3MIC, sr=96000, medium="air")
        ring = rh.RingStub(np.zeros((STREAM_RING["ring_rows"], 3), np.float32))
        cur = 0
        with contextlib.redirect_stdout(io.StringIO()):
            for b, i in enumerate(range(0, len(x) - 127, 128)):
                blk = x[i:i + 128]
                ring.write(blk)
                c, dl, _ = od(blk)
                if len(c) > 0:
                    dd = [cur + int(v) for v in dl]
                    ndet[s] += len(c)
                    try:
                        for k in np.argsort(dd):
                            res = m.locate(int(c[k]), dd[k], ring)
                            if res is not None:
                                rows.append((s, b, res[0], res[1]))
                                break
                    except ValueError:
                        raised[s] = b
                        break
                cur += 128
    rows = np.asarray(rows, np.float64).reshape(-1, 4)
    np.savez_compressed(OUT / "stream_ring.npz", rows=rows, raised=raised, n_detections=ndet, env=env())
    print("stream_ring: streams", S, "detections", int(ndet.sum()), "located", len(rows), "raised", int((raised >= 0).sum()))


SPECTRAL = dict(seconds=2.0, seed=15)


def gen_spectral(det):
    """Spectral-flux row (a6).  librosa / loopmate are absent, so the pin is (a) the UNMODIFIED reference's
    detect_onsets_spectral run on top of oracle/librosa_standin.py (scipy.signal.ShortTimeFFT, librosa 0.9's
    filter-based peak picker) on the channel mean of a synthetic recording, and (b) the realtime onset strength
    (recording.py:273-296, un-normalised) through scipy.signal.ShortTimeFFT."""
    from oracle import librosa_standin as ls

    x, _ = synth.drum_recording(**SPECTRAL)
    mono = np.ascontiguousarray(x.mean(1), np.float32)
    peaks, oe = det.detect_onsets_spectral(mono, return_oe=True)
    flux = ls.onset_strength_shorttimefft(x, 2048, 128)
    np.savez_compressed(OUT / "spectral.npz", peaks=np.asarray(peaks, np.int64), oe=np.asarray(oe, np.float32),
                        flux2048=flux, x_sha=sha(x), env=env())
    print("spectral: frames", len(oe), "peaks", len(peaks), "flux frames", len(flux))



STFT_CASES = [
    dict(),
    dict(method="prezero"),
    dict(method="pre", hop_edge_padding=True),
    dict(frame_length=128, hop_length=32, n_fft=128, method="prezero"),
    dict(frame_length=200, hop_length=50, n_fft=256, hop_edge_padding=True),
]
STFT_ONSETS = [49100, 60391, 71700]


def stft_inputs():
    x, _ = synth.drum_recording(seconds=1.0, seed=14)
    return np.ascontiguousarray(x.T)  # [C, N], the layout data.stft indexes (audio[..., onset:...])


def gen_stft():
    """data.stft / window_contribution_weights (data.py:560-654) of the unmodified reference; librosa's get_window /
    pad_center come from oracle/librosa_standin.py (scipy.signal.get_window, zero padding)."""
    data = rh.load_reference_data()
    a = stft_inputs()
    out = {"env": env(), "x_sha": sha(a)}
    for i, kw in enumerate(STFT_CASES):
        out[f"S{i}"] = np.stack([data.stft(a, o, **kw) for o in STFT_ONSETS])
        out[f"M{i}"] = data.stft(a[1].copy(), STFT_ONSETS[0], **kw)
    w = np.hanning(256)
    out["wcw"] = data.window_contribution_weights(w, 64)
    out["wcw_edge"] = data.window_contribution_weights(w, 64, True)
    np.savez_compressed(OUT / "stft.npz", **out)
    print("stft ok", [out[f"S{i}"].shape for i in range(len(STFT_CASES))])


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    if "stft" in sys.argv:
        gen_stft()
        return
    if "fsm" in sys.argv:
        gen_stream_fsm(rh.load_reference()[1])
        return
    if "frames" in sys.argv:
        gen_frames()
        return
    det, ml = rh.load_reference()
    if "hits16_scalar" in sys.argv:  # child of gen_hits16, runs with numpy's SIMD sort kernels disabled
        _, _, fixed, raised = _hits16_reference(det)
        np.savez(sys.argv[-1], fixed=fixed, raised=raised)
        return
    if "spectral" in sys.argv:
        gen_spectral(det)
        return
    if "ring" in sys.argv:
        gen_stream_ring(det, ml)
        return
    if "sizes" in sys.argv:  # parity at the sizes BASELINE.json states
        gen_config0(det, ml)
        gen_hits16(det)
        return
    if "tools" in sys.argv:  # only the helper-surface fixtures
        gen_tools(det, ml)
        gen_stream_cc(det, ml)
        return
    gen_detect(det)
    gen_stream(det)
    gen_backtrack(det)
    gen_kernels(det)
    gen_fix(det)
    gen_locate(ml)
    gen_online_cc()
    gen_tools(det, ml)
    gen_stream_cc(det, ml)
    gen_frames()
    gen_stream_fsm(ml)
    gen_config0(det, ml)
    gen_hits16(det)
    gen_stream_ring(det, ml)
    gen_spectral(det)
    gen_stft()


if __name__ == "__main__":
    main()
