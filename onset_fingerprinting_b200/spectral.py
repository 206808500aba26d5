"""Spectral-flux onset features on the GPU (K2, csrc/spectral_flux.cu).

Mirrors the two places the reference computes them:
  * realtime: RecAnalysis.fft + onset_strength (realtime/recording.py:273-311) -- one 2048-point
    Hann-windowed rFFT of the channel mean per 128-sample hop, log-power, half-wave rectified frame
    difference averaged over bins (the BASELINE config-5 shape) -> ``onset_strength``;
  * offline: detect_onsets_spectral (detection.py:89-128) -- |STFT(256, hop 32)| with an A-weighting
    ramp, rectified difference, 99.9-percentile normalisation, librosa-style peak picking.
librosa and loopmate are not part of the reference tree; their documented behaviour is restated
(centred zero-padded STFT with a periodic Hann window, ``util.peak_pick``, ``A_weighting``;
``EMA_MinMaxTracker`` is unknown, so the un-normalised flux is the parity point).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr, stream_ptr
from .detection import _to_dev


def hann(n: int, periodic: bool) -> np.ndarray:
    """scipy.signal.windows.hann(n, sym=not periodic) as float32 (recording.py:243 uses the symmetric
    one, librosa.stft the periodic one)."""
    m = n if periodic else n - 1
    return (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / m)).astype(np.float32)


def a_weighting(frequencies: np.ndarray, min_db: float = -80.0) -> np.ndarray:
    """IEC 61672 A-weighting in dB as librosa.A_weighting computes it."""
    f_sq = np.asanyarray(frequencies, dtype=np.float64) ** 2
    c = np.array([12194.217, 20.598997, 107.65265, 737.86223]) ** 2
    with np.errstate(divide="ignore"):
        w = 2.0 + 20.0 * (np.log10(c[0]) + 2 * np.log10(f_sq) - np.log10(f_sq + c[0]) - np.log10(f_sq + c[1])
                          - 0.5 * np.log10(f_sq + c[2]) - 0.5 * np.log10(f_sq + c[3]))
    return w if min_db is None else np.maximum(min_db, w)


def spectral_flux_batch(x, n_fft: int = 2048, hop: int = 128, center: bool = False, reflect: bool = False,
                        mode: str = "logpower", top_db: float = 0.0, window: np.ndarray | None = None,
                        weight: np.ndarray | None = None):
    """x [R, N, C] (or [R, N]) float32 -> flux [R, n_frames] float32 device tensor.
    center=False: frame j is the last n_fft samples after hop j+1 (realtime, recording.py:277),
    n_frames = N // hop.  center=True: frame j is centred on j*hop (librosa.stft), n_frames = 1 + N // hop.
    flux[j] compares frame j with frame j-1 (frame -1 = silence)."""
    torch = _lib.require_cuda()
    x = _to_dev(x, torch)
    if x.dim() == 2:
        x = x[:, :, None]
    R, N, Cn = x.shape
    n_frames = 1 + N // hop if center else N // hop
    if window is None:
        window = hann(n_fft, periodic=center)
    wd = torch.from_numpy(np.ascontiguousarray(window, dtype=np.float32)).cuda()
    wt = None if weight is None else torch.from_numpy(np.ascontiguousarray(weight, dtype=np.float32)).cuda()
    flux = torch.empty((R, n_frames), dtype=torch.float32, device="cuda")
    check(_lib.lib().ofp_spectral_flux(ptr(x), C.c_int64(R), C.c_int64(N), C.c_int64(x.stride(0)), C.c_int32(Cn),
                                       C.c_int32(n_fft), C.c_int32(hop), C.c_int32(int(center)),
                                       C.c_int32(int(reflect)), C.c_int32(0 if mode == "logpower" else 1),
                                       C.c_float(top_db), ptr(wd), ptr(wt), C.c_int32(n_frames), ptr(flux),
                                       stream_ptr()))
    return flux


def onset_strength(x: np.ndarray, n_fft: int = 2048, hop: int = 128, top_db: float = 0.0) -> np.ndarray:
    """Un-normalised onset envelope of RecAnalysis.onset_strength (recording.py:282-296) for a whole
    recording x [N, C]: one value per hop."""
    return spectral_flux_batch(x[None], n_fft, hop, center=False, mode="logpower", top_db=top_db)[0].cpu().numpy()


def peak_pick_batch(oe, pre_max: int, post_max: int, pre_avg: int, post_avg: int, delta: float, wait: int,
                    cap: int | None = None):
    """librosa.util.peak_pick for every row of oe [R, F] -> (peaks [R, cap] int32, counts [R])."""
    torch = _lib.require_cuda()
    R, F = oe.shape
    if cap is None:
        cap = F // max(wait, 1) + 2
    peaks = torch.empty((R, cap), dtype=torch.int32, device="cuda")
    cnt = torch.empty((R,), dtype=torch.int32, device="cuda")
    check(_lib.lib().ofp_peak_pick(ptr(oe.contiguous()), C.c_int32(R), C.c_int32(F), C.c_int32(pre_max),
                                   C.c_int32(post_max), C.c_int32(pre_avg), C.c_int32(post_avg), C.c_float(delta),
                                   C.c_int32(wait), ptr(peaks), ptr(cnt), C.c_int32(cap), stream_ptr()))
    return peaks, cnt


def detect_onsets_spectral(x: np.ndarray, n_fft: int = 256, hop: int = 32, sr: int = 96000, return_oe: bool = False):
    """detection.detect_onsets_spectral (detection.py:89-128) for a mono signal x [N]."""
    torch = _lib.require_cuda()
    freq = np.fft.fftfreq(n_fft, 1 / sr)[: n_fft // 2 + 1]
    aw = a_weighting(freq)
    weight = (aw - aw.min()) / np.abs(aw.min())
    flux = spectral_flux_batch(np.ascontiguousarray(x, np.float32)[None, :, None], n_fft, hop, center=True,
                               mode="magnitude", weight=weight)
    oe = flux[:, 1:] * (n_fft // 2 + 1)  # D[:, 1:] - D[:, :-1] then .mean(0): undo nothing, keep the mean
    oe = oe / (n_fft // 2 + 1)
    oe = oe / torch.quantile(oe[0], 0.999)
    peaks, cnt = peak_pick_batch(oe, int(0.12 * sr // hop), int(0.01 * sr // hop), int(0.12 * sr // hop),
                                 int(0.01 * sr // hop + 1), 0.1, int(sr * 0.07 // hop))
    p = peaks[0, : int(cnt[0].item())].cpu().numpy().astype(np.int64) * hop
    return (p, oe[0].cpu().numpy()) if return_oe else p


def tempogram_batch(oe, win_length: int = 384, first_frame: int = 0, every: int = 1, window: np.ndarray | None = None):
    """RecAnalysis.tempogram (realtime/recording.py:313-327) for frames first_frame, first_frame + every, ... of every
    row of the onset envelope oe [R, F]: the autocorrelation of the Hann-windowed (symmetric window, recording.py:244)
    last win_length envelope values, normalised by (max + 1e-10).  Returns [R, n_selected, win_length] float32."""
    torch = _lib.require_cuda()
    oe = _to_dev(oe, torch)
    if oe.dim() == 1:
        oe = oe[None]
    R, F = oe.shape
    n_sel = 0 if first_frame >= F else (F - 1 - first_frame) // every + 1
    if window is None:
        window = hann(win_length, periodic=False)
    wd = torch.from_numpy(np.ascontiguousarray(window, dtype=np.float32)).cuda()
    tg = torch.empty((R, n_sel, win_length), dtype=torch.float32, device="cuda")
    check(_lib.lib().ofp_tempogram(ptr(oe.contiguous()), C.c_int32(R), C.c_int64(F), ptr(wd), C.c_int32(win_length),
                                   C.c_int64(first_frame), C.c_int64(every), C.c_int64(n_sel), ptr(tg), stream_ptr()))
    return tg
