"""Stand-in for the three librosa calls of detection.detect_onsets_spectral (detection.py:96-121), built on
scipy primitives.  TEST INFRASTRUCTURE (oracle/make_golden.py registers it as ``librosa`` before importing the
unmodified reference).

librosa is a dependency of the reference that is neither in its tree nor installed here, so the
spectral row cannot be pinned against librosa itself.  This module is the next best thing: an implementation
that shares NO code with the CUDA kernel or with oracle/spectral_np.py -- the STFT is scipy.signal.ShortTimeFFT,
the peak picker is librosa 0.9's own formulation (scipy.ndimage.maximum_filter1d / uniform_filter1d with shifted
origins, explicit edge corrections, greedy `wait`), A-weighting is the IEC 61672 closed form -- and the
reference's own detect_onsets_spectral runs on top of it.  Documented librosa behaviour restated:
  stft(y, n_fft, hop_length): centred frames (frame k covers y[k*hop - n_fft/2 : k*hop + n_fft/2], zero padded
      -- pad_mode="constant", the default since librosa 0.10), periodic Hann window, 1 + len(y)//hop frames,
      complex64 for float32 input;
  A_weighting(f, min_db=-80);
  util.peak_pick(x, pre_max, post_max, pre_avg, post_avg, delta, wait): x[n] is a peak iff
      x[n] == max(x[n-pre_max : n+post_max]), x[n] >= mean(x[n-pre_avg : n+post_avg]) + delta, n - previous > wait
      (half-open windows, clipped at the ends).
"""
from __future__ import annotations

import types

import numpy as np
import scipy.ndimage
import scipy.signal


def stft(y, *, n_fft=2048, hop_length=None, **_):
    hop = hop_length or n_fft // 4
    win = scipy.signal.get_window("hann", n_fft, fftbins=True)
    sft = scipy.signal.ShortTimeFFT(win, hop, fs=1.0, fft_mode="onesided", scale_to=None, phase_shift=None)
    S = sft.stft(np.asarray(y, np.float64), p0=0, p1=1 + len(y) // hop)
    return S.astype(np.complex64)


def A_weighting(frequencies, *, min_db=-80.0):
    f_sq = np.asanyarray(frequencies, dtype=float) ** 2.0
    const = np.array([12194.217, 20.598997, 107.65265, 737.86223]) ** 2.0
    with np.errstate(divide="ignore"):
        weights = 2.0 + 20.0 * (np.log10(const[0]) + 2 * np.log10(f_sq) - np.log10(f_sq + const[0])
                                - np.log10(f_sq + const[1]) - 0.5 * np.log10(f_sq + const[2])
                                - 0.5 * np.log10(f_sq + const[3]))
    return weights if min_db is None else np.maximum(min_db, weights)


def peak_pick(x, *, pre_max, post_max, pre_avg, post_avg, delta, wait):
    x = np.asarray(x)
    pre_max, post_max, pre_avg, post_avg, wait = (int(np.ceil(v)) for v in (pre_max, post_max, pre_avg, post_avg, wait))
    max_length = pre_max + post_max
    max_origin = int(np.ceil(0.5 * (pre_max - post_max)))
    mov_max = scipy.ndimage.maximum_filter1d(x, max_length, mode="constant", origin=max_origin, cval=x.min())
    avg_length = pre_avg + post_avg
    avg_origin = int(np.ceil(0.5 * (pre_avg - post_avg)))
    mov_avg = scipy.ndimage.uniform_filter1d(x, avg_length, mode="nearest", origin=avg_origin)
    n = 0
    while n - pre_avg < 0 and n < x.shape[0]:  # the filter's edge handling is not a clipped mean: redo the ends
        mov_avg[n] = np.mean(x[max(n - pre_avg, 0): n + post_avg])
        n += 1
    n = max(x.shape[0] - post_avg, 0)
    while n < x.shape[0]:
        mov_avg[n] = np.mean(x[max(n - pre_avg, 0): n + post_avg])
        n += 1
    det = x * (x == mov_max)
    det = det * (det >= (mov_avg + delta))
    peaks, last = [], -np.inf
    for i in np.nonzero(det)[0]:
        if i > last + wait:
            peaks.append(i)
            last = i
    return np.array(peaks, dtype=int)


def get_window(window, Nx, *, fftbins=True):
    """librosa.filters.get_window for a window given by name: scipy.signal.get_window (periodic for fftbins)."""
    return scipy.signal.get_window(window, Nx, fftbins=fftbins)


def pad_center(data, *, size, axis=-1, **kw):
    """librosa.util.pad_center: zeros on both sides, the left pad is (size - n) // 2."""
    n = data.shape[axis]
    lpad = int((size - n) // 2)
    widths = [(0, 0)] * data.ndim
    widths[axis] = (lpad, int(size - n - lpad))
    if lpad < 0:
        raise ValueError(f"Target size ({size}) must be at least input size ({n})")
    return np.pad(data, widths, **kw)


def module() -> types.ModuleType:
    m = types.ModuleType("librosa")
    m.stft, m.A_weighting = stft, A_weighting
    m.util = types.ModuleType("librosa.util")
    m.util.peak_pick = peak_pick
    m.util.pad_center = pad_center
    m.filters = types.ModuleType("librosa.filters")
    m.filters.get_window = get_window
    m.__standin__ = True
    return m


def onset_strength_shorttimefft(x, n_fft=2048, hop=128):
    """RecAnalysis.fft + onset_strength (realtime/recording.py:273-296) without the EMA normalisation, through
    scipy.signal.ShortTimeFFT: frame j = the last n_fft samples after hop j+1 (zeros before the start), symmetric
    Hann window, |X|^2 -> 10 log10(max(1e-10, .)) -> mean(max(0, s_j - s_{j-1})), s_{-1} = the all-zero frame."""
    mono = np.asarray(x, np.float32).mean(-1) if np.ndim(x) == 2 else np.asarray(x, np.float32)
    win = scipy.signal.windows.hann(n_fft).astype(np.float32)
    sft = scipy.signal.ShortTimeFFT(win.astype(np.float64), hop, fs=1.0, fft_mode="onesided", scale_to=None, phase_shift=None)
    F = len(mono) // hop
    shift = (n_fft // 2) // hop  # slice p is centred on p*hop; frame j ends at (j+1)*hop
    assert (n_fft // 2) % hop == 0
    S = sft.stft(mono.astype(np.float64), p0=1 - shift, p1=F + 1 - shift)  # [bins, F]
    s = 10.0 * np.log10(np.maximum(1e-10, S.real ** 2 + S.imag ** 2))
    prev = np.concatenate([np.full((s.shape[0], 1), -100.0), s[:, :-1]], axis=1)
    return np.maximum(0.0, s - prev).mean(0).astype(np.float32)
