"""Device mirror of the window-extraction part of the reference's ``data.py``.

``FrameExtractor`` / ``FastFrameExtractor`` (data.py:55-192) cut a fixed-length window per onset
group out of a recording; ``MCPOSD`` (233-327) holds such windows with their strike positions;
``batch_cc`` (226-230) is the row-wise full cross-correlation the CC models feed on.  Here the
windows are gathered by ``ofp_extract_frames`` straight from recordings resident in HBM -- e.g. the
``[R, N, C]`` batch K1/K4 just processed (``extract_frames_batch``) -- so the offline dataset build
(BASELINE config 2, "data.py path") never copies audio back to the host.

``stft`` / ``stft_frame`` / ``window_contribution_weights`` (data.py:560-654), the complex short-time spectra around an
onset that feed the MFCC features, run on ``ofp_stft_frames`` (``stft_batch``: many onsets per launch).

The wav + json session files are read and written by ``posd.py`` (``MCPOSD.from_file`` goes through it);
the augmentation pipeline is host-side and out of scope.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr, stream_ptr
from .detection import _to_dev


def extract_frames_batch(audio, hit_rec, hit_onsets, frame_length: int, pre_samples: int = 0, shifts=None,
                         use_min_onset: bool = True):
    """Windows for H hits of a batch: audio [R, N, C] float32 (device or numpy), hit_rec [H] int32 or
    None (all in recording 0), hit_onsets [H, C] int32 -> frames [H, C, frame_length] float32 CUDA tensor.
    Raises IndexError where numpy's ``view[start]`` would."""
    torch = _lib.require_cuda()
    audio = _to_dev(audio, torch)
    R, N, Cn = audio.shape
    on = torch.as_tensor(hit_onsets).to(device="cuda", dtype=torch.int32).contiguous()
    H = on.shape[0]
    rec = None if hit_rec is None else torch.as_tensor(hit_rec).to(device="cuda", dtype=torch.int32).contiguous()
    sh = None if shifts is None else torch.as_tensor(shifts).to(device="cuda", dtype=torch.int32).contiguous()
    out = torch.empty((H, Cn, frame_length), dtype=torch.float32, device="cuda")
    status = torch.empty((H,), dtype=torch.int32, device="cuda")
    check(_lib.lib().ofp_extract_frames(ptr(audio), C.c_int64(N), C.c_int64(audio.stride(0)), C.c_int32(Cn), ptr(rec),
                                        ptr(on), ptr(sh), C.c_int32(H), C.c_int32(frame_length),
                                        C.c_int32(pre_samples), C.c_int32(int(use_min_onset)), ptr(out), ptr(status),
                                        stream_ptr()))
    if H and bool((status != 0).any()):
        bad = int(torch.nonzero(status)[0].item())
        raise IndexError(f"index out of bounds for the sliding-window view (hit {bad})")
    return out


class FrameExtractor:
    """data.py:55-120: ``fe(audio[N(, C)], onsets[O(, C)]) -> frames`` as a numpy array of shape
    [O, C, F] (2-D audio) or [O, F] (1-D audio)."""

    def __init__(self, frame_length: int, pre_samples: int, max_shift: int = 0, add_pre_samples: bool = False,
                 use_min_onset: bool = True):
        self.frame_length = frame_length + (pre_samples if add_pre_samples else 0)
        self.pre_samples = pre_samples
        self.max_shift = max_shift
        self.use_min_onset = use_min_onset

    def __call__(self, audio: np.ndarray, onsets: np.ndarray) -> np.ndarray:
        audio = np.ascontiguousarray(audio, dtype=np.float32)
        onsets = np.asarray(onsets)
        shifts = None
        if self.max_shift:
            shifts = np.random.randint(-self.max_shift, self.max_shift + 1, len(onsets)).astype(np.int32)
        one_d = audio.ndim == 1
        a3 = audio[None, :, None] if one_d else audio[None]
        on2 = onsets.reshape(len(onsets), -1)
        out = extract_frames_batch(a3, None, on2, self.frame_length, self.pre_samples, shifts,
                                   use_min_onset=self.use_min_onset or one_d)
        out = out.cpu().numpy()
        return out[:, 0, :] if one_d else out


class FastFrameExtractor:
    """data.py:123-192: holds the recording on the device and returns the window tensor; 2-D onsets
    always use the minimum of each group."""

    def __init__(self, audio, onsets, frame_length: int, pre_samples: int, max_shift: int = 0,
                 add_pre_samples: bool = False, device=None):
        torch = _lib.require_cuda()
        self.torch = torch
        self.frame_length = frame_length + (pre_samples if add_pre_samples else 0)
        self.pre_samples, self.max_shift = pre_samples, max_shift
        audio = np.ascontiguousarray(audio, dtype=np.float32) if isinstance(audio, np.ndarray) else audio
        self._one_d = audio.ndim == 1
        a = _to_dev(audio, torch)
        self.audio = (a[None, :, None] if self._one_d else a[None]).contiguous()
        on = torch.as_tensor(np.asarray(onsets)).to("cuda")
        self.onsets = (on.min(1).values if on.dim() == 2 else on).to(torch.int32)[:, None].contiguous()
        if not max_shift:
            self.frames = self._cut(None)

    def _cut(self, shifts):
        on = self.onsets.expand(-1, self.audio.shape[2]).contiguous()
        out = extract_frames_batch(self.audio, None, on, self.frame_length, self.pre_samples, shifts)
        return out[:, 0, :] if self._one_d else out

    def __call__(self):
        if self.max_shift:
            shifts = self.torch.randint(-self.max_shift, self.max_shift + 1, (len(self.onsets),), device="cuda",
                                        dtype=self.torch.int32)
            return self._cut(shifts)
        return self.frames


def batch_cc(a, b):
    """data.py:226-230: full cross-correlation of every row pair, [n, L] x [n, L] -> [n, 2L-1]
    (``F.conv1d`` with padding L-1 == np.correlate(a_i, b_i, "full")), on ofp_correlate_full."""
    from .multilateration import correlate_full

    return correlate_full(a, b)


def window_contribution_weights(window: np.ndarray, hop_length: int, hop_edge_padding: bool = False) -> np.ndarray:
    """data.py:560-577: how much of the signal of interest each STFT frame of ``stft`` saw through the window --
    the trapezoid integral of the window's first i samples for i = start, start + hop, ... (start = half a window,
    or one hop with ``hop_edge_padding``), mirrored for the trailing frames and scaled to a maximum of 1.  Host
    arithmetic on a handful of numbers."""
    window = np.asarray(window)
    first = hop_length if hop_edge_padding else len(window) // 2
    rising = [float(np.trapezoid(window[:i])) for i in range(first, len(window) + hop_length, hop_length)]
    w = np.array(rising + rising[-2::-1])
    return w / w.max()


def _hann_periodic(n: int) -> np.ndarray:
    # librosa.filters.get_window("hann", n, fftbins=True) = scipy's periodic Hann, float64 (data.py:624)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def _pad_center(v: np.ndarray, size: int) -> np.ndarray:
    # librosa.util.pad_center along the last axis: zeros, the odd one on the right (data.py:589, 626)
    lpad = (size - v.shape[-1]) // 2
    width = [(0, 0)] * (v.ndim - 1) + [(lpad, size - v.shape[-1] - lpad)]
    return np.pad(v, width)


def stft_frames_dev(frames, n_fft: int, window: np.ndarray):
    """np.fft.rfft(window * pad_center(frame, n_fft)) for every row of frames [F, frame_length] (device tensor or
    numpy) on ofp_stft_frames -> complex64 device tensor [F, n_fft/2 + 1]."""
    torch = _lib.require_cuda()
    fr = _to_dev(np.ascontiguousarray(frames, dtype=np.float32) if isinstance(frames, np.ndarray) else frames, torch)
    fr = fr.float().contiguous()
    F, flen = fr.shape
    wd = torch.from_numpy(np.ascontiguousarray(window, dtype=np.float64)).cuda()
    if wd.numel() != n_fft:
        raise ValueError(f"window of {wd.numel()} samples for n_fft = {n_fft}")
    out = torch.empty((F, n_fft // 2 + 1, 2), dtype=torch.float32, device="cuda")
    check(_lib.lib().ofp_stft_frames(ptr(fr), C.c_int64(F), C.c_int32(flen), C.c_int32(n_fft), ptr(wd), ptr(out),
                                     stream_ptr()))
    return torch.view_as_complex(out)


def stft_frame(x: np.ndarray, n_fft: int, window: np.ndarray) -> np.ndarray:
    """Drop-in for data.stft_frame (data.py:580-590): one frame (or [C, frame] rows) -> rfft(window * x), with x
    centred in n_fft points when it is shorter.  The reference returns numpy's complex128; the transform here runs in
    double on the device and returns complex64, the precision ``stft`` stores it in."""
    x = np.asarray(x)
    rows = x.reshape(-1, x.shape[-1])
    if rows.shape[-1] > n_fft:
        raise ValueError("operands could not be broadcast together")  # what window * x raises in the reference
    S = stft_frames_dev(rows, n_fft, window).cpu().numpy()
    return S.reshape(x.shape[:-1] + (n_fft // 2 + 1,))


def stft_batch(audio, onsets, frame_length: int = 256, hop_length: int = 64, n_fft: int = 512,
               hop_edge_padding: bool = False, method: str = "zerozero"):
    """``stft`` (data.py:593-654) around MANY onsets of one recording in one launch: audio [N] or [C, N], onsets [H]
    -> complex64 device tensor [H, (C,) n_fft/2 + 1, n_frames].  The padded excerpts are assembled on the device
    (zeros / preceding audio in front, zeros or nothing behind, per ``method``) and cut into hop-spaced frames."""
    torch = _lib.require_cuda()
    a = _to_dev(np.ascontiguousarray(audio, dtype=np.float32) if isinstance(audio, np.ndarray) else audio, torch).float()
    mono = a.dim() == 1
    if mono:
        a = a[None]
    Cn, N = a.shape
    on = torch.as_tensor(np.asarray(onsets, dtype=np.int64)).cuda().reshape(-1)
    H = on.numel()
    pad = frame_length - hop_length if hop_edge_padding else frame_length // 2
    if method not in ("zerozero", "prezero", "pre"):
        raise ValueError(f"method {method!r}")
    if H and (int(on.min()) < (0 if method == "zerozero" else pad) or int(on.max()) + frame_length > N):
        raise IndexError("stft excerpt outside the recording")
    body = a[:, (on[:, None] + torch.arange(frame_length, device="cuda")[None]).reshape(-1)]
    body = body.reshape(Cn, H, frame_length).permute(1, 0, 2)                       # [H, C, frame_length]
    zeros = torch.zeros((H, Cn, pad), dtype=torch.float32, device="cuda")
    if method == "zerozero":
        front = zeros
    else:
        pre = a[:, (on[:, None] - pad + torch.arange(pad, device="cuda")[None]).reshape(-1)]
        front = pre.reshape(Cn, H, pad).permute(1, 0, 2)
    y = torch.cat([front, body] + ([] if method == "pre" else [zeros]), dim=-1)     # [H, C, Ly]
    n_frames = 1 + (y.shape[-1] - frame_length) // hop_length
    frames = y.unfold(-1, frame_length, hop_length)[..., :n_frames, :].contiguous()  # [H, C, n_frames, frame_length]
    window = _hann_periodic(frame_length)
    if n_fft > frame_length:
        window = _pad_center(window, n_fft)
    S = stft_frames_dev(frames.reshape(-1, frame_length), n_fft, window)
    S = S.reshape(H, Cn, n_frames, n_fft // 2 + 1).permute(0, 1, 3, 2)               # [H, C, bins, n_frames]
    return S[:, 0] if mono else S


def stft(audio: np.ndarray, onset: int, frame_length: int = 256, hop_length: int = 64, n_fft: int = 512,
         hop_edge_padding: bool = False, method: str = "zerozero") -> np.ndarray:
    """Drop-in for data.stft (data.py:593-654): complex64 [(C,) n_fft/2 + 1, n_frames] around one onset."""
    S = stft_batch(audio, [int(onset)], frame_length, hop_length, n_fft, hop_edge_padding, method)[0]
    return S.cpu().numpy()


class MCPOSD:
    """data.py:233-327 without the file reader: multi-channel windows ``x [O, C, F]`` and positions
    ``y [O, 2]`` as device tensors; ``ds[0] -> (x, y)`` (the reference's batch_size=None contract)."""

    def __init__(self, data, onsets, sound_positions, frame_length: int = 256, pre_samples: int = 0,
                 max_shift: int = 0, n_extractions: int = 1, device=None, channels=None):
        torch = _lib.require_cuda()
        if channels is not None:
            data = data[:, channels]
        self.data = data
        self.frame_extractor = FastFrameExtractor(data, onsets, frame_length, pre_samples, max_shift)
        pos = np.asarray(sound_positions, dtype=np.float32)
        if n_extractions == 1 and max_shift == 0:
            self.y = torch.from_numpy(pos).cuda()
            self.x = self.frame_extractor()
            self.straight = True
        else:
            self.y = torch.from_numpy(np.concatenate([pos] * n_extractions)).cuda()
            self.straight = False
        self.n_extractions = n_extractions

    def __getitem__(self, index):
        if self.straight:
            return self.x, self.y
        torch = self.frame_extractor.torch
        return torch.cat([self.frame_extractor() for _ in range(self.n_extractions)]), self.y

    def __len__(self):
        return 1

    @classmethod
    def from_file(cls, folder, name: str, frame_length: int = 256, pre_samples: int = 0, max_shift: int = 0,
                  n_extractions: int = 1, channels=None):
        """data.py:285-311: ``<folder>/<name>.wav`` + ``<name>.json`` (POSD session, posd.py)."""
        from . import posd

        data, _, onsets, positions, _ = posd.read_session(folder, name)
        return cls(data, onsets, positions, frame_length, pre_samples, max_shift, n_extractions, channels=channels)

    @classmethod
    def from_xy(cls, x, y):
        ds = cls.__new__(cls)
        ds.x, ds.y, ds.straight = x, y, True
        return ds

    def split(self, r: float = 0.8):
        import torch

        n = len(self.y)
        idx = torch.randperm(n, device=self.y.device)
        k = int(n * r)
        return self.from_xy(self.x[idx[:k]], self.y[idx[:k]]), self.from_xy(self.x[idx[k:]], self.y[idx[k:]])
