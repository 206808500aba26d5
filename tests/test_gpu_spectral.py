"""K2 spectral flux on the GPU vs the numpy restatement (oracle/spectral_np.py).  Floating point bar:
the flux is a mean of dB differences; both sides use a float32 FFT (numpy pocketfft vs the Stockham
kernel), whose results agree to ~1e-7 of the largest bin, i.e. ~1e-3 relative in noise-floor bins and
hence a few 1e-3 dB in the flux: tolerance 2e-3 relative + 3e-3 dB absolute.  Peak indices are
identical on identical envelopes."""
import numpy as np
import pytest

from onset_fingerprinting_b200 import synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_onset_strength_realtime_shape():
    from onset_fingerprinting_b200 import spectral
    from oracle import spectral_np as sp

    x, _ = synth.drum_recording(seconds=0.9, seed=5, first_hit=20000)
    for top_db in (0.0, 80.0):
        got = spectral.onset_strength(x, 2048, 128, top_db=top_db)
        want = sp.onset_strength(x, 2048, 128, top_db=top_db)
        assert got.shape == want.shape == (len(x) // 128,)
        assert np.allclose(got, want, rtol=2e-3, atol=3e-3), float(np.abs(got - want).max())
    assert want.max() > 1.0  # the hits are visible in the flux


def test_small_fft_sizes_and_batch():
    from onset_fingerprinting_b200 import spectral
    from oracle import spectral_np as sp

    xs, _ = synth.drum_batch(3, seconds=0.6, seed=50, first_hit=10000)
    for n_fft, hop in ((256, 32), (512, 128), (4096, 256)):
        got = spectral.spectral_flux_batch(xs, n_fft, hop).cpu().numpy()
        for r in range(3):
            want = sp.onset_strength(xs[r], n_fft, hop)
            assert np.allclose(got[r], want, rtol=2e-3, atol=3e-3)


def test_detect_onsets_spectral_restatement():
    from onset_fingerprinting_b200 import spectral
    from oracle import spectral_np as sp

    x, truth = synth.drum_recording(seconds=1.5, seed=6, first_hit=30000, noise=1e-3)
    mono = x[:, 2].copy()
    freq = np.fft.fftfreq(256, 1 / 96000)[:129]
    aw = spectral.a_weighting(freq)
    weight = (aw - aw.min()) / np.abs(aw.min())
    peaks, oe = spectral.detect_onsets_spectral(mono, return_oe=True)
    peaks_o, oe_o = sp.detect_onsets_spectral(mono, weight)
    assert np.allclose(oe, oe_o, rtol=2e-3, atol=1e-3)
    # peak picking itself: identical on identical input
    pk, cnt = spectral.peak_pick_batch(torch.from_numpy(oe_o[None]).cuda(), 360, 30, 360, 31, 0.1, 210)
    assert pk[0, : int(cnt[0])].cpu().tolist() == (peaks_o // 32).tolist()
    # every reported peak sits on a synthetic hit (within 3 ms of an arrival)
    arr = truth["arrival"][:, 2]
    assert len(peaks) >= 5
    assert all(np.abs(arr - p).min() < 300 for p in peaks)
    assert peaks.tolist() == peaks_o.tolist()


def test_spectral_golden_reference_over_scipy_standin(golden_dir):
    """Pin of row a6: the unmodified reference's detect_onsets_spectral run on top of an independent scipy
    implementation of the three librosa calls (oracle/librosa_standin.py: scipy.signal.ShortTimeFFT, librosa 0.9's
    filter-based peak_pick), and the realtime onset strength through ShortTimeFFT (tests/golden/spectral.npz).
    Bars: envelope 2e-4 of its maximum (float32 FFT both sides), identical peak indices."""
    import hashlib

    from onset_fingerprinting_b200 import spectral
    from oracle.make_golden import SPECTRAL

    g = np.load(golden_dir / "spectral.npz")
    x, _ = synth.drum_recording(**SPECTRAL)
    assert hashlib.sha1(np.ascontiguousarray(x).tobytes()).hexdigest() == str(g["x_sha"])
    mono = np.ascontiguousarray(x.mean(1), np.float32)
    peaks, oe = spectral.detect_onsets_spectral(mono, return_oe=True)
    assert oe.shape == g["oe"].shape
    assert np.abs(oe - g["oe"]).max() <= 2e-4 * np.abs(g["oe"]).max()
    assert peaks.tolist() == g["peaks"].tolist() and len(peaks) == 9
    flux = spectral.onset_strength(x, 2048, 128)
    assert flux.shape == g["flux2048"].shape
    assert np.abs(flux - g["flux2048"]).max() <= 3e-3 + 2e-3 * np.abs(g["flux2048"]).max()


def test_peak_pick_kernel_vs_filter_formulation():
    """ofp_peak_pick against librosa 0.9's filter formulation (oracle/librosa_standin.py) on random envelopes."""
    from onset_fingerprinting_b200 import spectral
    from oracle import librosa_standin as ls

    rng = np.random.default_rng(4)
    for trial in range(40):
        n = int(rng.integers(5, 900))
        x = np.abs(rng.standard_normal(n)).astype(np.float32)
        if trial % 3 == 0:
            x[rng.integers(0, n, n // 4)] = 0.0
        if trial % 4 == 0:
            x = (np.round(x * 4) / 4).astype(np.float32)
        a = dict(pre_max=int(rng.integers(1, 40)), post_max=int(rng.integers(1, 40)), pre_avg=int(rng.integers(1, 60)),
                 post_avg=int(rng.integers(1, 60)), delta=float(rng.uniform(0, 0.5)), wait=int(rng.integers(0, 30)))
        pk, cnt = spectral.peak_pick_batch(torch.from_numpy(x[None]).cuda(), a["pre_max"], a["post_max"], a["pre_avg"],
                                           a["post_avg"], a["delta"], a["wait"], cap=n + 1)
        assert pk[0, : int(cnt[0])].cpu().tolist() == ls.peak_pick(x, **a).tolist(), (trial, a)


def test_flux_properties_large_batch():
    """BASELINE-size properties the numpy restatement is too slow to check: per-recording independence of the
    batch, determinism, hop-shift equivariance (delaying a recording by k hops delays its flux by k frames
    once the 2048-sample window no longer sees the inserted silence boundary differently), silence -> 0."""
    from onset_fingerprinting_b200 import spectral

    R, N, hop = 64, 96000, 128
    x = synth.drum_batch_device(R, N, seed=77)
    f = spectral.spectral_flux_batch(x, 2048, hop)
    assert tuple(f.shape) == (R, N // hop) and bool(torch.isfinite(f).all())
    assert torch.equal(f, spectral.spectral_flux_batch(x, 2048, hop))
    assert torch.equal(spectral.spectral_flux_batch(x[5:9].contiguous(), 2048, hop), f[5:9])
    k = 37
    xs = torch.zeros_like(x)
    xs[:, k * hop:] = x[:, : N - k * hop]
    fs = spectral.spectral_flux_batch(xs, 2048, hop)
    # frame j of the delayed signal sees the same 2048 samples as frame j - k of the original wherever both windows
    # lie inside the recording or inside leading zeros; the synthetic recordings start with 0.5 s of noise, so compare
    # bitwise only where the original frame index is past the zero padding of the original (j - k >= 16)
    a, b = fs[:, k + 16:], f[:, 16: N // hop - k]
    assert torch.equal(a, b)
    z = spectral.spectral_flux_batch(torch.zeros(2, 8192, 3, device="cuda"), 2048, hop)
    assert float(z.abs().max()) == 0.0


@pytest.mark.parametrize("channels", [1, 2, 3, 4])
def test_warp_kernel_equals_generic_kernel_2048(channels, monkeypatch):
    """The 2048-point warp-per-frame kernel against the generic Stockham kernel (validated against the numpy
    restatement at other sizes) over the option space: channel counts (the 3-channel vector staging and its
    scalar fallback), centred / reflected framing, both modes, top_db, hops, lengths that are no multiple of
    anything, a strided batch view.  Same float32 FFT noise floor as the oracle comparison."""
    from onset_fingerprinting_b200 import spectral

    rng = np.random.default_rng(channels)
    N = 30011
    x = torch.from_numpy((rng.standard_normal((3, N, channels)) * 0.05).astype(np.float32)).cuda()
    x[1, 9000:9400] += 0.8  # a loud burst: large dynamic range between bins
    w = (0.5 + rng.random(1025)).astype(np.float32)
    cases = [dict(center=False), dict(center=False, top_db=60.0), dict(center=True), dict(center=True, reflect=True),
             dict(center=True, mode="magnitude", weight=w), dict(center=False, mode="magnitude")]
    for hop in (128, 256, 130):
        for kw in cases:
            got = spectral.spectral_flux_batch(x, 2048, hop, **kw)
            monkeypatch.setenv("OFP_K2_GENERIC", "1")
            want = spectral.spectral_flux_batch(x, 2048, hop, **kw)
            monkeypatch.delenv("OFP_K2_GENERIC")
            assert got.shape == want.shape
            tol = 3e-3 if kw.get("mode") != "magnitude" else 1e-4 * float(want.abs().max()) + 1e-6
            assert torch.allclose(got, want, rtol=2e-3, atol=tol), (channels, hop, kw.keys(), float((got - want).abs().max()))
    # a batch that is a strided view (every other recording of a larger buffer)
    big = torch.from_numpy((rng.standard_normal((6, N, channels)) * 0.05).astype(np.float32)).cuda()
    assert torch.equal(spectral.spectral_flux_batch(big[::2], 2048, 128), spectral.spectral_flux_batch(big[::2].contiguous(), 2048, 128))


def test_stft_around_onsets_vs_reference_golden(golden_dir):
    """data.stft / stft_frame / window_contribution_weights (data.py:560-654) on ofp_stft_frames against the unmodified
    reference's output (tests/golden/stft.npz).  Floating-point bar: numpy transforms `window * x` in double and the
    result is stored as complex64; the kernel transforms in double as well (radix-2 Stockham) and rounds once, so
    the two agree to 1e-6 of the largest bin (an ulp or two of complex64)."""
    from onset_fingerprinting_b200 import data
    from oracle.make_golden import STFT_CASES, STFT_ONSETS, stft_inputs

    g = np.load(golden_dir / "stft.npz")
    a = stft_inputs()
    for i, kw in enumerate(STFT_CASES):
        want, mono = g[f"S{i}"], g[f"M{i}"]
        got = data.stft_batch(a, STFT_ONSETS, **kw).cpu().numpy()              # all onsets in one launch
        assert got.shape == want.shape and got.dtype == np.complex64
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max(), (i, float(np.abs(got - want).max()))
        one = data.stft(a, STFT_ONSETS[1], **kw)                                # the reference's call
        assert np.abs(one - want[1]).max() <= 1e-6 * np.abs(want).max()
        got1 = data.stft(a[1].copy(), STFT_ONSETS[0], **kw)                     # mono audio
        assert got1.shape == mono.shape and np.abs(got1 - mono).max() <= 1e-6 * np.abs(mono).max()
    # stft_frame: one frame, shorter than n_fft -> centred
    fr = a[:, 50000:50200]
    win = np.pad(np.hanning(200), (28, 28))
    lw = (256 - 200) // 2
    want = np.fft.rfft(win * np.pad(fr, [(0, 0), (lw, 256 - 200 - lw)])).astype(np.complex64)
    got = data.stft_frame(fr, 256, win)
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    w = np.hanning(256)
    assert np.allclose(data.window_contribution_weights(w, 64), g["wcw"], rtol=1e-14, atol=0)
    assert np.allclose(data.window_contribution_weights(w, 64, True), g["wcw_edge"], rtol=1e-14, atol=0)
    with pytest.raises(IndexError):
        data.stft(a, a.shape[1] - 10)
