"""numpy restatement of the spectral-flux path (TEST INFRASTRUCTURE).

recording.py:273-311 (realtime) and detection.py:89-128 (offline).  librosa / loopmate are absent from
the reference tree, so this restates their documented behaviour; "parity unpinned" for this row."""
import numpy as np


def hann(n, periodic):
    m = n if periodic else n - 1
    return (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / m)).astype(np.float32)


def onset_strength(x, n_fft=2048, hop=128, top_db=0.0):
    """x [N, C] -> flux per hop: frame j = last n_fft samples after hop j+1 (recording.py:277)."""
    mono = x.mean(-1).astype(np.float32) if x.ndim == 2 else x.astype(np.float32)
    w = hann(n_fft, periodic=False)
    pad = np.concatenate([np.zeros(n_fft, np.float32), mono])
    prev = np.full(n_fft // 2 + 1, 10.0 * np.log10(1e-10), np.float32)
    out = []
    for j in range(len(mono) // hop):
        e = (j + 1) * hop + n_fft
        f = np.fft.rfft(w * pad[e - n_fft:e])
        mag = (f.real.astype(np.float32) ** 2 + f.imag.astype(np.float32) ** 2)
        s = (10.0 * np.log10(np.maximum(1e-10, mag))).astype(np.float32)
        lo = s.max() - top_db if top_db > 0 else -np.inf
        out.append(np.maximum(0.0, np.maximum(s, lo) - np.maximum(prev, lo)).mean())
        prev = s
    return np.asarray(out, np.float32)


def stft_mag(x, n_fft, hop):
    """|librosa.stft(x, n_fft, hop)|: centred, zero padded, periodic Hann."""
    w = hann(n_fft, periodic=True)
    pad = np.concatenate([np.zeros(n_fft // 2, np.float32), x.astype(np.float32), np.zeros(n_fft // 2, np.float32)])
    frames = 1 + len(x) // hop
    D = np.empty((n_fft // 2 + 1, frames), np.float32)
    for j in range(frames):
        D[:, j] = np.abs(np.fft.rfft(w * pad[j * hop:j * hop + n_fft]))
    return D


def peak_pick(x, pre_max, post_max, pre_avg, post_avg, delta, wait):
    peaks, last = [], -(1 << 30)
    for n in range(len(x)):
        if x[n] != x[max(0, n - pre_max):n + post_max + 1].max():
            continue
        if x[n] < x[max(0, n - pre_avg):n + post_avg + 1].mean() + delta:
            continue
        if n - last <= wait:
            continue
        peaks.append(n)
        last = n
    return np.asarray(peaks, np.int64)


def detect_onsets_spectral(x, weight, n_fft=256, hop=32, sr=96000):
    D = stft_mag(x, n_fft, hop) * weight[:, None].astype(np.float32)
    oe = np.maximum(0.0, D[:, 1:] - D[:, :-1]).mean(0)
    oe = oe / np.percentile(oe, 99.9)
    p = peak_pick(oe, int(0.12 * sr // hop), int(0.01 * sr // hop), int(0.12 * sr // hop), int(0.01 * sr // hop + 1),
                  0.1, int(sr * 0.07 // hop))
    return p * hop, oe
