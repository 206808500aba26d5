"""numpy restatement of the spectral-flux path (TEST INFRASTRUCTURE).

recording.py:273-311 (realtime) and detection.py:89-128 (offline).  librosa / loopmate are absent from the
reference tree, so this restates their documented behaviour.  Pin: tests/golden/spectral.npz = the unmodified
reference's detect_onsets_spectral run over an INDEPENDENT scipy implementation of the three librosa calls
(oracle/librosa_standin.py: ShortTimeFFT, filter-based peak picker) -- tests/test_spectral_cpu.py."""
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
    """librosa.util.peak_pick's documented conditions, half-open windows: x[n] == max(x[n-pre_max : n+post_max]),
    x[n] >= mean(x[n-pre_avg : n+post_avg]) + delta, n - previous > wait; zero-valued samples are never peaks."""
    peaks, last = [], -(1 << 30)
    for n in range(len(x)):
        if n - last <= wait:
            continue
        if x[n] != x[max(0, n - pre_max):max(n + post_max, n + 1)].max() or x[n] == 0:
            continue
        if x[n] < np.float32(x[max(0, n - pre_avg):max(n + post_avg, n + 1)].astype(np.float64).mean()) + np.float32(delta):
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


def tempogram_frame(oe_last, window):
    """recording.py:313-327 for one frame: oe_last = the last W onset-envelope values (zeros before the start)."""
    W = len(window)
    pad = 2 * W - 1
    spec = np.fft.rfft(window.astype(np.float64) * oe_last.astype(np.float64), n=pad)
    tg = np.fft.irfft(spec.real ** 2 + spec.imag ** 2, n=pad)[:W]
    return (tg / (tg.max() + 1e-10)).astype(np.float32)


def stft_around_onset(audio, onset, frame_length=256, hop_length=64, n_fft=512, hop_edge_padding=False,
                      method="zerozero"):
    """data.stft (data.py:593-654) restated: the excerpt after the onset, padded in front with zeros or the
    preceding audio and behind with zeros or nothing, cut into hop-spaced frames; every frame centred in n_fft
    points, multiplied by the (pad-centred) periodic Hann window of frame_length and transformed with numpy's
    double-precision rfft; stored as complex64 [(C,) n_fft/2 + 1, n_frames]."""
    audio = np.asarray(audio)
    pad = frame_length - hop_length if hop_edge_padding else frame_length // 2
    y = audio[..., onset:onset + frame_length]
    z = np.zeros(y.shape[:-1] + (pad,), np.float32)
    pre = audio[..., onset - pad:onset]
    y = np.concatenate({"zerozero": (z, y, z), "prezero": (pre, y, z), "pre": (pre, y)}[method], axis=-1)
    window = hann(frame_length, periodic=True).astype(np.float64)
    lw = (n_fft - frame_length) // 2
    window = np.pad(window, (lw, n_fft - frame_length - lw))
    n_frames = 1 + (y.shape[-1] - frame_length) // hop_length
    out = np.empty(y.shape[:-1] + (n_fft // 2 + 1, n_frames), np.complex64)
    for i in range(n_frames):
        fr = y[..., hop_length * i:hop_length * i + frame_length]
        fr = np.pad(fr, [(0, 0)] * (fr.ndim - 1) + [(lw, n_fft - frame_length - lw)])
        out[..., i] = np.fft.rfft(window * fr)
    return out


def window_contribution_weights(window, hop_length, hop_edge_padding=False):
    """data.py:560-577."""
    start = hop_length if hop_edge_padding else len(window) // 2
    w = [np.trapezoid(window[:i]) for i in range(start, len(window) + hop_length, hop_length)]
    w = w + w[-2::-1]
    return np.array(w) / max(w)
