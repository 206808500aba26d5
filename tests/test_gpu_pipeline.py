"""Whole hot path (K1 -> K3 -> K4 -> K5) on a batch vs the CPU oracle run recording by recording,
plus size-independent properties at a larger size (time-shift equivariance, determinism)."""
import numpy as np
import pytest

from onset_fingerprinting_b200 import synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def oracle_pipeline(orc, x, sensors, medium):
    ch, on, _ = orc.detect_onsets_amplitude(x, sr=96000, return_rel=False)
    groups = orc.find_onset_groups(on, ch, 1000, x.shape[1])
    if groups is None:
        return np.zeros((0, x.shape[1]), np.int64), np.zeros((0, x.shape[1]), np.int64), np.zeros((0, 2)), np.zeros(0, int)
    fixed, fstat, _ = orc.fix_onsets(x, groups, return_status=True)
    m = orc.Multilaterate3D(sensors, sr=96000, medium=medium)
    xy = np.full((len(fixed), 2), np.nan)
    st = np.zeros(len(fixed), int)
    for h, row in enumerate(fixed):
        got, st[h] = m.locate_hit([0, 1, 2], row[:3])
        if got is not None:
            xy[h] = got
    return groups, fixed, xy, st


def test_pipeline_matches_oracle():
    from onset_fingerprinting_b200 import pipeline
    from oracle import oracle as orc

    xs, truth = synth.drum_batch(6, seconds=2.0, seed=500)
    hp = pipeline.HotPath(len(xs), 3, synth.SENSORS_3MIC, medium="air", sr=96000)
    hb = hp.run(torch.from_numpy(xs).cuda(), return_rel=False)
    rec = hb.rec.cpu().numpy()
    n_loc = 0
    for r in range(len(xs)):
        groups, fixed, xy, st = oracle_pipeline(orc, xs[r], synth.SENSORS_3MIC, "air")
        sel = rec == r
        assert np.array_equal(hb.onsets.cpu().numpy()[sel], groups)
        assert np.array_equal(hb.fixed.cpu().numpy()[sel], fixed)
        assert np.array_equal(hb.loc_status.cpu().numpy()[sel], st)
        got = hb.xy.cpu().numpy()[sel]
        ok = st == 0
        n_loc += int(ok.sum())
        assert np.array_equal(got[ok], xy[ok])  # bit-equal to the oracle's MINPACK replay
        # and the located positions are physically right: within a few cm of the synthetic truth (sanity, not parity)
        if ok.any():
            true_xy = truth[r]["pos"][: len(groups)]
            assert np.median(np.linalg.norm(got[ok] - true_xy[ok], axis=1)) < 6.0
    assert n_loc > 40


def test_host_buffer_entry_equals_device_entry():
    """pipeline.HotPath.run_host (pinned host audio in, time-segmented upload under the detector, host records
    out -- bench.py's e2e leg) returns exactly what run() returns on the resident batch, for segment lengths that
    do and do not divide the recording, with and without the rel envelope."""
    from onset_fingerprinting_b200 import pipeline

    xs, _ = synth.drum_batch(5, seconds=2.0, seed=700)
    xd = torch.from_numpy(xs).cuda()
    hp = pipeline.HotPath(len(xs), 3, synth.SENSORS_3MIC, medium="air", sr=96000)
    want = hp.run(xd, return_rel=True)
    want = {k: getattr(want, k).cpu().numpy().copy() for k in ("rec", "onsets", "fixed", "lags", "fix_status", "xy",
                                                               "loc_status", "onset_counts", "rel")}
    xh = torch.from_numpy(xs).pin_memory()
    for seg, with_rel in ((49152, True), (50000, False), (1 << 20, True)):
        relh = torch.zeros((len(xs), (xs.shape[1] // 128) * 128, 3), dtype=torch.float32).pin_memory() if with_rel else None
        got = hp.run_host(xh, rel_host=relh, segment=seg)
        for k in ("rec", "onsets", "fixed", "lags", "fix_status", "loc_status", "onset_counts"):
            assert np.array_equal(got[k].numpy(), want[k]), (seg, k)
        assert np.array_equal(got["xy"].numpy(), want["xy"], equal_nan=True)
        if with_rel:
            assert np.array_equal(relh.numpy(), want["rel"])
    assert len(want["rec"]) > 40


def test_pipeline_properties_large():
    """Device-generated batch far larger than the oracle can check: results are deterministic, every
    hit has one onset per channel inside the recording, and shifting the audio by whole blocks shifts
    every onset by the same amount (the detector has no absolute-time dependence after warm-up)."""
    from onset_fingerprinting_b200 import pipeline

    R, N = 2000, 96000 * 2
    x = synth.drum_batch_device(R, N, seed=77)
    hp = pipeline.HotPath(R, 3, synth.SENSORS_3MIC, medium="air", sr=96000)
    a = hp.run(x)
    b = hp.run(x)
    for f in ("rec", "onsets", "fixed", "lags", "fix_status", "loc_status"):
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    assert torch.equal(torch.nan_to_num(a.xy), torch.nan_to_num(b.xy))
    assert a.rec.shape[0] > 10 * R
    assert int(a.onsets.min()) >= 0 and int(a.onsets.max()) < N
    assert bool(((a.fixed - a.onsets).abs() <= 2 * 40).all())
    located = a.loc_status == 0
    assert float(located.float().mean()) > 0.8
    # fsolve is free to leave the drumhead (nothing in the reference prevents it); almost all hits stay on it
    assert float((a.xy[located].norm(dim=1) < 17.78 + 3).float().mean()) > 0.97
    # checksum of checksums: per-recording onset sums are identical between the two passes and
    # match a recomputation from the flat onset list
    per_rec = torch.zeros(R, dtype=torch.int64, device="cuda").index_add_(0, a.rec.long(), a.onsets.sum(1).long())
    assert int(per_rec.sum()) == int(a.onsets.sum())


def test_realtime_block_api_matches_reference_semantics():
    """BlockLocator.detect_hits == the reference's PlayRec.detect_hits loop restated with the oracle:
    block detector -> sort -> streaming locate."""
    from onset_fingerprinting_b200.realtime import audio as rt
    from oracle import oracle as orc

    x, truth = synth.drum_recording(seconds=2.0, seed=11, first_hit=30000)
    conf = {"sensor_locations": synth.SENSORS_3MIC, "medium": "air", "c": None}
    bl = rt.BlockLocator(conf)
    od = orc.Detector(3, 128, **{k: v for k, v in rt.REALTIME_DETECTOR.items()})
    mo = orc.Multilaterate3D(synth.SENSORS_3MIC, sr=96000, medium="air")
    got, want, pending = [], [], []
    for i in range(0, len(x) - 127, 128):
        r = bl.detect_hits(x[i:i + 128])
        if r is not None:
            got.append((i // 128, r.x, r.y))
        c, d, _ = od(x[i:i + 128])
        for cc, dd in sorted(zip(c.tolist(), d.tolist()), key=lambda t: t[1]):
            pending.append((int(cc), i + int(dd)))
            if len(pending) == 3:  # hits are far apart: three detections complete a group
                sens = [p[0] for p in pending]; ons = [p[1] for p in pending]
                xy, st = mo.locate_hit(sens, ons)
                if xy is not None:
                    want.append((i // 128, xy[0], xy[1]))
                pending = []
    assert len(got) > 5
    assert [g[0] for g in got] == [w[0] for w in want]
    assert np.allclose([g[1:] for g in got], [w[1:] for w in want], rtol=1e-12)


def test_stream_batch_equals_independent_detectors():
    from onset_fingerprinting_b200.realtime import audio as rt
    from oracle import oracle as orc

    S = 37
    xs, _ = synth.drum_batch(S, seconds=0.3, seed=900, first_hit=4000)
    sb = rt.StreamBatch(S)
    ods = [orc.Detector(3, 128, **rt.REALTIME_DETECTOR) for _ in range(S)]
    for i in range(0, xs.shape[1] - 127, 128):
        ch, dl, cnt, rel = sb.process(xs[:, i:i + 128], return_rel=True)
        ch, dl, cnt, rel = ch.cpu().numpy(), dl.cpu().numpy(), cnt.cpu().numpy(), rel.cpu().numpy()
        for s in range(S):
            c, d, r = ods[s](xs[s, i:i + 128])
            assert ch[s, :cnt[s]].tolist() == c.tolist() and dl[s, :cnt[s]].tolist() == d.tolist()
            assert np.array_equal(rel[s], r)


def test_pipeline_without_any_hit():
    """A batch in which nothing crosses the threshold (noise only) and one in which a single recording has hits:
    K3 / K4 / K5 run on zero hits (empty tensors of the right shapes), and the empty recordings of a mixed batch
    contribute nothing -- through both entry points."""
    from onset_fingerprinting_b200 import pipeline

    rng = np.random.default_rng(3)
    quiet = (1e-4 * rng.standard_normal((4, 96000, 3))).astype(np.float32)
    hp = pipeline.HotPath(4, 3, synth.SENSORS_3MIC, medium="air", sr=96000)
    hb = hp.run(torch.from_numpy(quiet).cuda(), return_rel=True)
    assert hb.rec.numel() == 0 and hb.onsets.shape == (0, 3) and hb.fixed.shape == (0, 3) and hb.xy.shape == (0, 2)
    assert hb.rel is not None and tuple(hb.rel.shape) == (4, 96000 // 128 * 128, 3)
    hh = hp.run_host(torch.from_numpy(quiet).pin_memory())
    assert len(hh["rec"]) == 0 and tuple(hh["xy"].shape) == (0, 2) and int(hh["onset_counts"].sum()) == 0
    mixed = quiet.copy()
    mixed[2], _ = synth.drum_recording(seconds=1.0, seed=77)
    hb = hp.run(torch.from_numpy(mixed).cuda(), return_rel=False)
    rec = hb.rec.cpu().numpy()
    assert len(rec) > 0 and set(rec.tolist()) == {2}
