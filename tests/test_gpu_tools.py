"""Helper surface of detection.py / multilateration.py on the GPU (SURVEY 8a rows a3, a4, a9, a12, a13,
8f rank 1) against what the unmodified reference returned (tests/golden/tools.npz, stream_cc.npz) and
against the numpy restatements in oracle/oracle.py."""
import contextlib
import io

import numpy as np
import pytest

from onset_fingerprinting_b200 import synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir / "tools.npz")


@pytest.fixture(scope="module")
def inp():
    from oracle.make_golden import tools_inputs

    return tools_inputs()


@pytest.fixture(scope="module")
def det():
    from onset_fingerprinting_b200 import detection

    return detection


@pytest.fixture(scope="module")
def ml():
    from onset_fingerprinting_b200 import multilateration

    return multilateration


def test_filter_data_in_place(det, gold, inp):
    for d in ("up", "down"):
        x = inp["fd_x"].copy()
        r = det.filter_data(x, d)
        assert r is x and np.array_equal(x, gold[f"fd_{d}"])
    with pytest.raises(RuntimeError):
        det.filter_data(inp["fd_x"].copy(), "sideways")


def test_detect_onset_region(det, gold, inp):
    a = [det.detect_onset_region(x, int(o)) for x, o in zip(inp["or_x"], inp["or_on"])]
    assert a == gold["or_a"].tolist()
    b = det.detect_onset_region_batch(inp["or_x"], inp["or_on"], n=128, median_filter_size=7, threshold_factor=0.3)
    assert b.cpu().tolist() == gold["or_b"].tolist()


def test_adjust_onset_rel(det, gold, inp):
    got = [det.adjust_onset_rel(list(o), x, y, int(l)) for o, x, y, l in
           zip(inp["rel_on"], inp["rel_x"], inp["rel_y"], inp["rel_lag"])]
    assert np.array_equal(np.asarray(got), gold["rel_out"])


@pytest.mark.parametrize("tag,args", [("lo2", (3000, 4, 2, 96000, "low")), ("hi3", (500, 4, 3, 96000, "high"))])
def test_butterworth_filter_streaming(det, gold, inp, tag, args):
    bw = det.ButterworthFilter(*args)
    for k, blk in enumerate(inp["bw_x"]):
        assert np.array_equal(bw(blk), gold[f"bw_{tag}"][k])  # float32 DF2T, bit for bit


def test_butterworth_matches_detector_highpass(det, golden_dir):
    g = np.load(golden_dir / "kernels.npz")
    bw = det.ButterworthFilter(2000, 3, 4, 96000, "high")
    for k, blk in enumerate(g["hp_in"]):
        assert np.array_equal(bw(blk), g["hp_out"][k])


def test_find_lag_and_multi(ml, gold, inp):
    from oracle import oracle as orc

    got = [ml.find_lag(a, b) for a, b in zip(inp["fl_a"], inp["fl_b"])]
    assert got == gold["fl"].tolist()
    cc = ml.correlate_full(inp["fl_a"][:3], inp["fl_b"][:3])
    for k in range(3):
        assert np.array_equal(cc[k], orc.correlate_full(inp["fl_a"][k], inp["fl_b"][k]))
    for k in range(12):
        lags, vals = ml.find_lag_multi(inp["fl_a"][k], inp["fl_b"][k], 3)
        n = len(lags)
        assert lags.tolist() == gold["flm_lags"][k][:n].tolist()
        assert np.allclose(vals, gold["flm_vals"][k][:n], rtol=1e-5)


def test_solve_trilateration_2d(ml, gold, inp):
    want = gold["tri_xy"]
    xy, ier = ml.solve_trilateration_batch(inp["tri_sens"][:, 0], inp["tri_sens"][:, 1], inp["tri_sens"][:, 2],
                                           inp["tri_da"], inp["tri_db"], inp["tri_seed"])
    xy, ier = xy.cpu().numpy(), ier.cpu().numpy()
    conv = np.isfinite(want[:, 0])
    assert np.array_equal(ier == 1, conv)  # same None set as fsolve
    err = np.abs(xy[conv] - want[conv]) / np.maximum(np.abs(want[conv]), 1e-3)
    assert err.max() <= 1e-9  # north_star: 1e-4
    k = int(np.nonzero(conv)[0][0])
    one = ml.solve_trilateration(tuple(inp["tri_sens"][k, 0]), tuple(inp["tri_sens"][k, 1]),
                                 tuple(inp["tri_sens"][k, 2]), inp["tri_da"][k], inp["tri_db"][k], inp["tri_seed"][k])
    assert np.allclose(one, want[k], rtol=1e-9)
    if (~conv).any():
        k = int(np.nonzero(~conv)[0][0])
        assert ml.solve_trilateration(tuple(inp["tri_sens"][k, 0]), tuple(inp["tri_sens"][k, 1]),
                                      tuple(inp["tri_sens"][k, 2]), inp["tri_da"][k], inp["tri_db"][k],
                                      inp["tri_seed"][k]) is None


def test_multilaterate_2d_streaming(ml, gold):
    from oracle.make_golden import MESH3

    m = ml.Multilaterate(MESH3, sr=96000, medium="drumhead")
    on, want = gold["m2_onsets"], gold["m2_res"]
    for h in range(len(on)):
        m.ongoing = []
        r = None
        for s in np.argsort(on[h], kind="stable"):
            r = m.locate(int(s), int(on[h, s]))
        if np.isfinite(want[h, 0]):
            assert r is not None and np.allclose(r, want[h], rtol=1e-9, atol=1e-9), h
        else:
            assert r is None, h


def test_multilaterate_paired(ml, gold):
    from oracle.make_golden import MESH4

    mp = ml.MultilateratePaired(MESH4, scale=10, sr=96000)
    for k, (lg, i) in enumerate(zip(gold["mp_lags"], gold["mp_i"])):
        want = gold["mp_res"][k]
        if np.isfinite(want[0]):
            assert np.allclose(mp.locate([int(lg[0]), int(lg[1])], int(i)), want, rtol=1e-9, atol=1e-9)
        else:
            with pytest.raises(TypeError):  # the reference unpacks None (multilateration.py:828)
                mp.locate([int(lg[0]), int(lg[1])], int(i))
    got = np.asarray([mp.locate_cc(gold["mp_cc_x"][k], 200, int(k % 4)) for k in range(20)])
    assert np.allclose(got, gold["mp_cc"], rtol=1e-12)


def test_lag_intensity_map(ml, gold):
    lm, sa, sb = ml.lag_intensity_map((10.0, 5.0, 8.0), (-12.0, 3.0, 6.0), reflectivity=0.5, sr=96000)
    assert np.array_equal(lm, gold["lim_lag"])
    assert np.allclose(sa, gold["lim_a"], rtol=1e-6) and np.allclose(sb, gold["lim_b"], rtol=1e-6)


def test_detector_init_calibration(det, gold, inp):
    od = det.AmplitudeOnsetDetector(3, 128, sr=96000)
    od.init(inp["init_x"])
    # dB envelopes: float32 log10 differs from numpy's by a few ulp -> absolute 1e-3 dB
    for name, attr in (("init_mins", "mins"), ("init_maxs", "maxs"), ("init_on", "on_threshold"),
                       ("init_off", "off_threshold"), ("init_noise", "noise_max")):
        assert np.allclose(getattr(od, attr), gold[name], atol=1e-3, rtol=1e-5), name


def test_streaming_locate_with_ring_refinement(golden_dir):
    """PlayRec.detect_hits with rec_audio: every detection's locate() result, block by block."""
    from oracle.make_golden import STREAM_CC
    from onset_fingerprinting_b200 import detection, multilateration
    from onset_fingerprinting_b200.realtime.audio import DeviceRing

    g = np.load(golden_dir / "stream_cc.npz")["rows"]
    x, _ = synth.drum_recording(**STREAM_CC)
    od = detection.AmplitudeOnsetDetector(3, 128, hipass_freq=0, fast_ar=(0.3, 800), slow_ar=(8000, 8000),
                                          on_threshold=0.45, off_threshold=0.45, cooldown=1323, sr=96000)
    m = multilateration.Multilaterate3D(synth.SENSORS_3MIC, sr=96000, medium="air")
    ring = DeviceRing(4096, 3)
    rows, cur = [], 0
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
    assert rows.shape == g.shape
    assert np.array_equal(rows[:, :3], g[:, :3])
    assert np.array_equal(np.isfinite(rows[:, 3]), np.isfinite(g[:, 3]))
    ok = np.isfinite(g[:, 3])
    assert np.allclose(rows[ok, 3:], g[ok, 3:], rtol=1e-9, atol=1e-9)


def test_window_argmax_dataset_builder():
    """notebooks/refresh.org:262-279: og[i] + argmax(audio[og[i] : og[i] + tol, i]) per hit and channel."""
    from onset_fingerprinting_b200 import detection as det

    rng = np.random.default_rng(8)
    R, N, Cn, H = 3, 5000, 4, 40
    x = rng.standard_normal((R, N, Cn)).astype(np.float32)
    x[:, 100:110, 1] = 7.0  # a plateau: the FIRST maximum wins
    rec = rng.integers(0, R, H).astype(np.int32)
    on = rng.integers(0, N - 10, (H, Cn)).astype(np.int32)
    on[0] = [95, 95, 95, 4995]
    on[1, 2] = -1
    got = det.max_onsets_batch(x, torch.from_numpy(rec), torch.from_numpy(on), 300).cpu().numpy()
    for h in range(H):
        for c in range(Cn):
            if on[h, c] < 0:
                assert got[h, c] == -1
            else:
                assert got[h, c] == on[h, c] + int(np.argmax(x[rec[h], on[h, c]:on[h, c] + 300, c])), (h, c)


def test_tempogram_vs_fft_formula():
    """RecAnalysis.tempogram (recording.py:313-327): irfft(|rfft(w * oe[-W:], n=2W-1)|^2)[:W] / (max + 1e-10); the
    kernel evaluates the same autocorrelation directly."""
    from onset_fingerprinting_b200 import spectral
    from oracle import spectral_np

    rng = np.random.default_rng(9)
    W, F = 384, 900
    oe = np.abs(rng.standard_normal((2, F))).astype(np.float32)
    oe[:, ::40] += 3.0
    tg = spectral.tempogram_batch(oe, W, first_frame=10, every=37).cpu().numpy()
    win = spectral.hann(W, periodic=False)
    frames = list(range(10, F, 37))
    assert tg.shape == (2, len(frames), W)
    for r in range(2):
        for k, j in enumerate(frames):
            last = np.zeros(W, np.float32)
            lo = max(0, j - W + 1)
            last[W - (j + 1 - lo):] = oe[r, lo:j + 1]
            want = spectral_np.tempogram_frame(last, win)
            assert np.abs(tg[r, k] - want).max() <= 2e-5, (r, j)
            assert abs(tg[r, k].max() - 1.0) < 1e-6


def test_paired_xcorr_vs_torch():
    """model.paired_xcorr (model.py:12-45) against the reference's own formulation (grouped F.conv1d) in torch."""
    import torch.nn.functional as F

    from onset_fingerprinting_b200 import model

    torch.manual_seed(3)
    B, Cc, K, V = 5, 3, 4, 64
    x = torch.randn(B, Cc * K, V)
    got = model.paired_xcorr(x.cuda(), Cc, K).cpu()
    xv = x.view(B, Cc, K, V)
    a = xv[:, :-1].reshape(B, (Cc - 1) * K, V)
    b = xv[:, 1:].reshape(B, (Cc - 1) * K, V)
    M = B * (Cc - 1) * K
    a_pad = F.pad(a, (V - 1, V - 1)).view(1, M, 3 * V - 2)
    want = F.conv1d(a_pad, b.reshape(M, 1, V), groups=M).view(B, Cc - 1, K, 2 * V - 1).mean(dim=2)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max())


def test_empty_batches_everywhere():
    """Every batch entry point on an EMPTY batch (zero hits / windows / pairs / recordings / frames): results of the
    usual shapes with a zero leading dimension, no error (torch's empty tensors carry null data pointers, which the
    C ABI must not reject when the count is zero).  Empty operands of cross_correlation_lag raise as np.correlate does."""
    import numpy as np
    from onset_fingerprinting_b200 import calibration, data, detection as det, model, multilateration as ml, spectral, synth

    audio = torch.randn(2, 4096, 3, device="cuda")

    def i32(*s):
        return torch.zeros(s, dtype=torch.int32, device="cuda")

    M = ml.Multilaterate3D(synth.SENSORS_3MIC, sr=96000, medium="air")
    xy, st = M.locate_batch(i32(0, 3))
    assert tuple(xy.shape) == (0, 2) and tuple(st.shape) == (0,)
    fixed, lags, fstat = det.fix_onsets_batch(audio, i32(0), i32(0, 3))
    assert tuple(fixed.shape) == (0, 3) and tuple(fstat.shape) == (0,)
    assert tuple(det.max_onsets_batch(audio, i32(0), i32(0, 3), 8).shape) == (0, 3)
    assert tuple(data.extract_frames_batch(audio, i32(0), i32(0, 3), 256).shape) == (0, 3, 256)
    z = torch.zeros(0, 64, device="cuda")
    assert tuple(ml.correlate_full(z, z).shape) == (0, 127)
    xy, ier = ml.solve_trilateration_batch(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0), np.zeros(0),
                                           np.zeros((0, 2)))
    assert tuple(xy.shape) == (0, 2) and tuple(ier.shape) == (0,)
    assert len(det.detect_onset_region_batch(np.zeros((0, 512), np.float32), np.zeros(0, np.int64))) == 0
    hit_rec, hit_on, n_groups = det.find_onset_groups_batch(i32(2, 16), i32(2, 16), i32(2), 3)
    assert hit_rec.numel() == 0 and tuple(hit_on.shape) == (0, 3) and n_groups.tolist() == [0, 0]
    assert tuple(model.CNN(256, 2).cuda()(torch.zeros(0, 3, 256, device="cuda")).shape) == (0, 2)
    assert tuple(model.CCCNN(256, 2).cuda()(torch.zeros(0, 3, 256, device="cuda")).shape) == (0, 2)
    assert tuple(calibration.FCNN(2, 2).cuda()(torch.zeros(0, 2, device="cuda")).shape) == (0, 2)
    assert tuple(spectral.spectral_flux_batch(torch.zeros(0, 8192, device="cuda")).shape) == (0, 64)
    assert tuple(spectral.spectral_flux_batch(torch.zeros(2, 100, device="cuda")).shape) == (2, 0)  # shorter than a hop
    peaks, n_peaks = spectral.peak_pick_batch(torch.zeros(0, 100, device="cuda"), 3, 3, 3, 3, 0.1, 2)
    assert peaks.shape[0] == 0 and tuple(n_peaks.shape) == (0,)
    assert spectral.tempogram_batch(torch.zeros(0, 1000, device="cuda")).shape[0] == 0
    ch, ix, cnt, rel = det.detect_onsets_amplitude_batch(np.zeros((0, 4096, 3), np.float32), sr=96000)
    assert ch.shape[0] == 0 and tuple(cnt.shape) == (0,) and tuple(rel.shape) == (0, 4096, 3)
    assert tuple(np.asarray(det.fix_onsets(np.zeros((4096, 3), np.float32), np.zeros((0, 3), np.int64))).shape) == (0, 3)
    with pytest.raises(ValueError):
        det.cross_correlation_lag(np.zeros(0, np.float32), np.zeros(0, np.float32), onsets=(0, 0))
    torch.cuda.synchronize()
