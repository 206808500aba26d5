"""Spectral-flux row: oracle/spectral_np.py (per-frame np.fft.rfft loops) against the golden the
unmodified reference produced on top of the independent scipy stand-in for librosa (ShortTimeFFT STFT, librosa
0.9's filter-based peak_pick) -- two implementations that share no code."""
import hashlib

import numpy as np

from onset_fingerprinting_b200 import synth
from oracle import librosa_standin as ls
from oracle import spectral_np
from oracle.make_golden import SPECTRAL


def test_offline_flux_and_peaks(golden_dir):
    g = np.load(golden_dir / "spectral.npz")
    x, _ = synth.drum_recording(**SPECTRAL)
    assert hashlib.sha1(np.ascontiguousarray(x).tobytes()).hexdigest() == str(g["x_sha"])
    mono = np.ascontiguousarray(x.mean(1), np.float32)
    freq = np.fft.fftfreq(256, 1 / 96000)[:129]
    aw = ls.A_weighting(freq)
    weight = (aw - aw.min()) / np.abs(aw.min())
    peaks, oe = spectral_np.detect_onsets_spectral(mono, weight)
    assert oe.shape == g["oe"].shape
    assert np.abs(oe - g["oe"]).max() <= 2e-4 * np.abs(g["oe"]).max()
    assert peaks.tolist() == g["peaks"].tolist() and len(peaks) == 9


def test_realtime_onset_strength(golden_dir):
    g = np.load(golden_dir / "spectral.npz")
    x, _ = synth.drum_recording(**SPECTRAL)
    flux = spectral_np.onset_strength(x, 2048, 128)
    assert flux.shape == g["flux2048"].shape
    assert np.abs(flux - g["flux2048"]).max() <= 2e-3 + 1e-3 * np.abs(g["flux2048"]).max()


def test_peak_pick_loop_equals_filter_formulation():
    """The windowed-loop statement of peak_pick (kernel, oracle) and librosa 0.9's filter formulation agree on
    random envelopes, including plateaus, zeros and short inputs."""
    rng = np.random.default_rng(4)
    for trial in range(60):
        n = int(rng.integers(5, 900))
        x = np.abs(rng.standard_normal(n)).astype(np.float32)
        if trial % 3 == 0:
            x[rng.integers(0, n, n // 4)] = 0.0
        if trial % 4 == 0:
            x = np.round(x * 4) / 4  # ties
        args = dict(pre_max=int(rng.integers(1, 40)), post_max=int(rng.integers(1, 40)), pre_avg=int(rng.integers(1, 60)),
                    post_avg=int(rng.integers(1, 60)), delta=float(rng.uniform(0, 0.5)), wait=int(rng.integers(0, 30)))
        want = ls.peak_pick(x, **args)
        got = spectral_np.peak_pick(x, **args)
        assert got.tolist() == want.tolist(), (trial, args)


def test_stft_around_onset_and_contribution_weights(golden_dir):
    """oracle/spectral_np.stft_around_onset against data.stft of the unmodified reference (tests/golden/stft.npz,
    generated over the stand-in for librosa's get_window / pad_center): both transform in double and round once to
    complex64, so they agree to an ulp of the largest bin."""
    from oracle.make_golden import STFT_CASES, STFT_ONSETS, stft_inputs

    g = np.load(golden_dir / "stft.npz")
    a = stft_inputs()
    assert hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest() == str(g["x_sha"])
    for i, kw in enumerate(STFT_CASES):
        want, mono = g[f"S{i}"], g[f"M{i}"]
        got = np.stack([spectral_np.stft_around_onset(a, o, **kw) for o in STFT_ONSETS])
        assert got.shape == want.shape and got.dtype == want.dtype == np.complex64
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
        got1 = spectral_np.stft_around_onset(a[1].copy(), STFT_ONSETS[0], **kw)
        assert got1.shape == mono.shape and np.abs(got1 - mono).max() <= 1e-6 * np.abs(mono).max()
    w = np.hanning(256)
    assert np.array_equal(spectral_np.window_contribution_weights(w, 64), g["wcw"])
    assert np.array_equal(spectral_np.window_contribution_weights(w, 64, True), g["wcw_edge"])
