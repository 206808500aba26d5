"""Host-side mirror of the reference's ``detection.py`` call surface, running on libofp.so.

Same names, argument meaning and return types as the reference
(/root/reference/onset_fingerprinting/detection.py) so existing callers can switch imports;
the arithmetic runs in the CUDA kernels behind include/ofp.h.  Inputs may be numpy arrays
(copied to the device) or CUDA torch tensors (used in place).  Batched variants
(``*_batch``) take ``[R, N, C]`` and return device tensors.

There is no CPU fallback (BASELINE.json north_star): without the built library or without
a CUDA device these raise ``OfpError``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import DetectorParams, OfpError, check, ptr, stream_ptr


def make_params(n_signals, block_size, floor=-70.0, hipass_freq=2000.0, fast_ar=(3.0, 383.0),
                slow_ar=(2205.0, 2205.0), on_threshold=0.5, off_threshold=0.1, cooldown=1323,
                sr=44100) -> DetectorParams:
    """The host-side conversions of AmplitudeOnsetDetector.__init__ (detection.py:631-712):
    float32 Butterworth coefficients, float32 reciprocals of attack/release, manual flag."""
    p = DetectorParams()
    p.n_channels, p.block_size = int(n_signals), int(block_size)
    p.use_hp = int(hipass_freq != 0)
    if p.use_hp:
        from scipy import signal as sig

        b, a = sig.butter(4, hipass_freq, btype="high", analog=False, output="ba", fs=sr)
        for i in range(5):
            p.b[i], p.a[i] = np.float32(b[i]), np.float32(a[i])
    p.manual = int(on_threshold > 1)
    p.cooldown = int(cooldown)
    p.floor_db = floor
    p.fast_att, p.fast_rel = np.float32(1 / fast_ar[0]), np.float32(1 / fast_ar[1])
    p.slow_att, p.slow_rel = np.float32(1 / slow_ar[0]), np.float32(1 / slow_ar[1])
    p.on_thr, p.off_thr = on_threshold, off_threshold
    p.alpha_min, p.alpha_max, p.minmin = 1e-4, 1e-5, 2.0
    return p


def _to_dev(x, torch, dtype=None):
    if isinstance(x, np.ndarray):
        if x.dtype != np.float32:
            # the reference's ctypes ndpointer rejects anything but float32 (detection.py:520-526)
            raise TypeError("audio must be float32")
        return torch.from_numpy(np.ascontiguousarray(x)).cuda(non_blocking=True)
    if not x.is_cuda:
        x = x.cuda()
    if x.dtype != torch.float32:
        raise TypeError("audio must be float32")
    return x.contiguous()


class BatchedOnsetDetector:
    """n_streams independent AmplitudeOnsetDetector states on the device (one ``ofp_detector``)."""

    def __init__(self, n_streams: int, n_signals: int, block_size: int = 32, **kw):
        self.torch = _lib.require_cuda()
        self.params = make_params(n_signals, block_size, **kw)
        self.n_streams, self.n_signals, self.block_size = int(n_streams), int(n_signals), int(block_size)
        self._h = C.c_void_p()
        check(_lib.lib().ofp_detector_create(C.byref(self._h), C.c_int64(self.n_streams), C.byref(self.params)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().ofp_detector_destroy(h)
            except Exception:
                pass
            self._h = None

    def reset(self):
        check(_lib.lib().ofp_detector_reset(self._h, stream_ptr()))

    STATE_FIELDS = ["z0", "z1", "z2", "z3", "yf", "ys", "min", "max", "prev", "state", "debounce"]

    def state(self) -> dict:
        """Copy of the per-lane state as numpy arrays of shape [n_streams, n_signals]."""
        torch = self.torch
        out = {}
        for i, name in enumerate(self.STATE_FIELDS):
            t = torch.empty((self.n_streams, self.n_signals), dtype=torch.float32 if i < 9 else torch.int32,
                            device="cuda")
            check(_lib.lib().ofp_detector_get_state(self._h, i, ptr(t), stream_ptr()))
            out[name] = t.cpu().numpy()
        return out

    def load_state(self, state: dict):
        torch = self.torch
        for i, name in enumerate(self.STATE_FIELDS):
            t = torch.from_numpy(np.ascontiguousarray(state[name])).cuda()
            check(_lib.lib().ofp_detector_set_state(self._h, i, ptr(t), stream_ptr()))
        torch.cuda.current_stream().synchronize()

    def warmup(self, x):
        """init_minmax_tracker (detection.py:827-840) on x [S, n, C]."""
        x = _to_dev(x, self.torch)
        assert x.dim() == 3 and x.shape[0] == self.n_streams and x.shape[2] == self.n_signals
        check(_lib.lib().ofp_detect_warmup(self._h, ptr(x), C.c_int64(x.shape[1]), C.c_int64(x.stride(0)),
                                           stream_ptr()))

    def process_block(self, x, return_rel=True):
        """One [S, B, C] block for every stream -> (channels [S, C], deltas [S, C], counts [S], rel)."""
        torch = self.torch
        S, B, Cn = self.n_streams, self.block_size, self.n_signals
        if not isinstance(x, np.ndarray) and x.is_cuda and x.dtype == torch.float32 and x.stride(2) == 1 \
                and x.stride(1) == Cn:
            pass  # a [S, B, C] window into longer per-stream buffers: used in place
        else:
            x = _to_dev(x, torch)
        assert tuple(x.shape) == (S, B, Cn), f"expected {(S, B, Cn)}, got {tuple(x.shape)}"
        ch = torch.empty((S, Cn), dtype=torch.int32, device="cuda")
        dl = torch.empty((S, Cn), dtype=torch.int32, device="cuda")
        cnt = torch.empty((S,), dtype=torch.int32, device="cuda")
        rel = torch.empty((S, B, Cn), dtype=torch.float32, device="cuda") if return_rel else None
        check(_lib.lib().ofp_detect_block(self._h, ptr(x), C.c_int64(x.stride(0)), ptr(rel), ptr(ch), ptr(dl),
                                          ptr(cnt), stream_ptr()))
        return ch, dl, cnt, rel

    def default_cap(self, n_samples: int) -> int:
        """Upper bound on onsets per recording: one per channel per cooldown (>= one block)."""
        return int(self.n_signals * (n_samples // max(self.params.cooldown, self.block_size) + 2))

    def detect_offline(self, x, warm_n: int, return_rel=True, cap: Optional[int] = None, out=None):
        """Warm-up on x[:, :warm_n] then the block loop from sample 0 (detection.py:70-82).
        Returns (channels [R, cap] int32, samples [R, cap] int32, counts [R] int32, rel [R, nb*B, C] | None).
        ``out`` may hold preallocated (channels, samples, counts, rel) tensors."""
        torch = self.torch
        x = _to_dev(x, torch)
        R, N, Cn = x.shape
        assert R == self.n_streams and Cn == self.n_signals
        B = self.block_size
        nb = N // B
        if out is not None:
            ch, ix, cnt, rel = out
            cap = ch.shape[1]
        else:
            if cap is None:
                cap = self.default_cap(N)
            ch = torch.empty((R, cap), dtype=torch.int32, device="cuda")
            ix = torch.empty((R, cap), dtype=torch.int32, device="cuda")
            cnt = torch.empty((R,), dtype=torch.int32, device="cuda")
            rel = torch.empty((R, nb * B, Cn), dtype=torch.float32, device="cuda") if return_rel else None
        check(_lib.lib().ofp_detect_offline(self._h, ptr(x), C.c_int64(N), C.c_int64(x.stride(0)), C.c_int64(warm_n),
                                            ptr(rel), C.c_int64(nb * B * Cn), ptr(ch), ptr(ix), ptr(cnt),
                                            C.c_int32(cap), stream_ptr()))
        return ch, ix, cnt, rel


def detect_onsets_amplitude_batch(x, block_size: int = 128, floor: float = -70.0, hipass_freq: float = 2000.0,
                                  fast_ar=(3.0, 383.0), slow_ar=(2205.0, 2205.0), on_threshold=0.5,
                                  off_threshold=0.1, cooldown: int = 1323, sr: int = 96000, return_rel=True,
                                  cap: Optional[int] = None):
    """detect_onsets_amplitude (detection.py:19-86) for a batch x [R, N, C] in one launch.
    Returns device tensors (channels [R, cap], onsets [R, cap], counts [R], rel | None); the first
    counts[r] entries of row r are recording r's onsets in the reference's order."""
    torch = _lib.require_cuda()
    x = _to_dev(x, torch)
    if x.shape[0] == 0:  # an empty batch: empty results of the usual shapes
        nb = x.shape[1] // block_size
        k = cap if cap is not None else 1
        e = torch.empty((0, k), dtype=torch.int32, device="cuda")
        rel = torch.empty((0, nb * block_size, x.shape[2]), dtype=torch.float32, device="cuda") if return_rel else None
        return e, e.clone(), torch.empty((0,), dtype=torch.int32, device="cuda"), rel
    det = BatchedOnsetDetector(x.shape[0], x.shape[2], block_size, floor=floor, hipass_freq=hipass_freq,
                               fast_ar=fast_ar, slow_ar=slow_ar, on_threshold=on_threshold,
                               off_threshold=off_threshold, cooldown=cooldown, sr=sr)
    return det.detect_offline(x, int(0.5 * sr), return_rel=return_rel, cap=cap)


def detect_onsets_amplitude(x: np.ndarray, block_size: int = 128, floor: float = -70.0,
                            hipass_freq: float = 2000.0, fast_ar=(3.0, 383.0), slow_ar=(2205.0, 2205.0),
                            on_threshold=0.5, off_threshold=0.1, cooldown: int = 1323, backtrack: bool = False,
                            backtrack_buffer_size: int = 128, backtrack_smooth_size: int = 5, sr: int = 96000):
    """Drop-in for detection.detect_onsets_amplitude (detection.py:19-86): x [N, C] float32 ->
    (channels_flat, onsets_flat, rel[n_blocks*block_size, C])."""
    ch, ix, cnt, rel = detect_onsets_amplitude_batch(
        x[None], block_size, floor, hipass_freq, fast_ar, slow_ar, on_threshold, off_threshold, cooldown, sr, True)
    if backtrack:
        backtrack_onsets_batch(rel, ch, ix, cnt, block_size, backtrack_buffer_size, backtrack_smooth_size)
    k = int(cnt[0].item())
    if k > ch.shape[1]:
        raise OfpError(f"onset buffer overflow ({k} > {ch.shape[1]})")
    return ch[0, :k].cpu().tolist(), ix[0, :k].cpu().tolist(), rel[0].cpu().numpy()


def _backtrack_consts(buffer_size: int, smooth_size: int):
    """detection.py:722-725."""
    alpha = np.float32(2 / (smooth_size + 1))
    tol = np.float32((1 - alpha) ** buffer_size)
    return alpha, tol


def backtrack_onsets_batch(rel, channels, samples, counts, block_size: int, buffer_size: int = 128,
                           smooth_size: int = 5, streaming: bool = False):
    """AmplitudeOnsetDetector.backtrack_onsets (detection.py:800-825) for every onset of a batch, in place
    on `samples`.  rel [R, n_rows, C] device tensor (whole envelope, or the last rows when streaming)."""
    assert block_size <= buffer_size, "backtrack_buffer_size should be at least block_size!"
    alpha, tol = _backtrack_consts(buffer_size, smooth_size)
    R, n_rows, Cn = rel.shape
    check(_lib.lib().ofp_backtrack_onsets(ptr(rel), C.c_int64(rel.stride(0)), C.c_int64(n_rows), C.c_int32(Cn),
                                          C.c_int32(block_size), C.c_int32(buffer_size), C.c_float(alpha),
                                          C.c_float(tol), C.c_int32(int(streaming)), ptr(channels), ptr(samples),
                                          ptr(counts), C.c_int32(R), C.c_int32(channels.shape[1]), stream_ptr()))
    return samples


def detect_onsets(x: np.ndarray, sr: int = 96000, method="amp"):
    """detection.py:12-16."""
    if method == "amp":
        return detect_onsets_amplitude(x, sr=sr)
    return detect_onsets_spectral(x, sr=sr)


def detect_onsets_spectral(x: np.ndarray, n_fft: int = 256, hop: int = 32, sr: int = 96000, return_oe: bool = False):
    """detection.py:89-128 (K2: csrc/spectral_flux.cu, see spectral.py for the librosa restatement)."""
    from . import spectral

    return spectral.detect_onsets_spectral(x, n_fft=n_fft, hop=hop, sr=sr, return_oe=return_oe)


class AmplitudeOnsetDetector:
    """Drop-in for detection.AmplitudeOnsetDetector (detection.py:595-840), one stream.

    ``od(block[B, C]) -> (channels, deltas, relative_envelope)`` with numpy results like the
    reference.  State lives on the device between calls."""

    def __init__(self, n_signals: int, block_size: int = 32, floor: float = -70.0, hipass_freq: float = 2000.0,
                 fast_ar=(3.0, 383.0), slow_ar=(2205.0, 2205.0), on_threshold: float = 0.5,
                 off_threshold: float = 0.1, cooldown: int = 1323, backtrack: bool = False,
                 backtrack_buffer_size: int = 80, backtrack_smooth_size: int = 5, sr: int = 44100):
        self.backtrack = backtrack
        self._bt = (backtrack_buffer_size, backtrack_smooth_size)
        self._hist = None
        if backtrack:
            assert block_size <= backtrack_buffer_size, "backtrack_buffer_size should be at least block_size!"
        self.n_signals, self.block_size = n_signals, block_size
        self.floor, self.on_threshold, self.off_threshold = floor, on_threshold, off_threshold
        self.manual = bool(on_threshold > 1)
        self.cooldown, self.sr = cooldown, sr
        self._hipass, self._fast_ar, self._slow_ar = hipass_freq, fast_ar, slow_ar
        self._det = BatchedOnsetDetector(1, n_signals, block_size, floor=floor, hipass_freq=hipass_freq,
                                         fast_ar=fast_ar, slow_ar=slow_ar, on_threshold=on_threshold,
                                         off_threshold=off_threshold, cooldown=cooldown, sr=sr)

    def __call__(self, x):
        ch, dl, cnt, rel = self._det.process_block(x[None] if x.ndim == 2 else x)
        if self.backtrack:
            torch = self._det.torch
            N = self._bt[0]
            if self._hist is None:  # the reference's ring starts uninitialised; rows before the start read 0
                self._hist = torch.zeros((1, N, self.n_signals), dtype=torch.float32, device="cuda")
            self._hist = torch.cat([self._hist[:, self.block_size:], rel], 1).contiguous()
            backtrack_onsets_batch(self._hist, ch, dl, cnt, self.block_size, N, self._bt[1], streaming=True)
        k = int(cnt[0].item())
        return ch[0, :k].cpu().numpy().astype(np.int64), dl[0, :k].cpu().numpy().astype(np.int64), rel[0].cpu().numpy()

    def init_minmax_tracker(self, x):
        self._det.warmup(x[None] if x.ndim == 2 else x)

    def backtrack_onsets(self, channels, deltas):
        """detection.py:800-825 on the detector's envelope history (the last backtrack_buffer_size rows
        ending with the block passed to the latest call)."""
        if self._hist is None:
            raise AttributeError("backtrack_onsets needs backtrack=True and at least one processed block")
        torch = self._det.torch
        k = len(channels)
        ch = torch.zeros((1, self.n_signals), dtype=torch.int32, device="cuda")
        dl = torch.zeros((1, self.n_signals), dtype=torch.int32, device="cuda")
        ch[0, :k] = torch.as_tensor(np.asarray(channels, np.int32))
        dl[0, :k] = torch.as_tensor(np.asarray(deltas, np.int32))
        cnt = torch.tensor([k], dtype=torch.int32, device="cuda")
        backtrack_onsets_batch(self._hist, ch, dl, cnt, self.block_size, self._bt[0], self._bt[1], streaming=True)
        return dl[0, :k].cpu().numpy().astype(np.int64)

    def init(self, x):
        """detection.py:842-888: calibrate per-channel thresholds and noise level on a take x [N, C] that
        holds silence and full-level playing.  The envelope followers run through ofp_ar_envelope and
        the high-pass through ofp_lfilter; the statistics (median / max / running maximum) are torch
        reductions on the device.  Like the reference it does not touch the min/max trackers __call__
        uses; it sets mins, maxs, on_threshold, off_threshold, noise_max and leaves the followers
        settled at the start of x."""
        torch = self._det.torch
        p = self._det.params
        xd = _to_dev(np.ascontiguousarray(x, dtype=np.float32), torch)
        st = self._det.state()
        hp = None
        if p.use_hp:  # the detector's own filter, state carried over (detection.py:852-853)
            hp = ButterworthFilter(self._hipass, self.n_signals, 4, self.sr, "high")
            hp.zi = torch.from_numpy(np.concatenate([st[f"z{i}"] for i in range(4)], 0)).cuda()
            xd = hp(xd)
        xd = 20 * torch.log10(torch.abs(xd + 1e-10))
        B, sr = self.block_size, self.sr
        fast = AREnvelopeFollower(torch.from_numpy(st["yf"]).expand(B, -1), self._fast_ar[0], self._fast_ar[1])
        slow = AREnvelopeFollower(torch.from_numpy(st["ys"]).expand(B, -1), self._slow_ar[0], self._slow_ar[1])

        def run(block):
            # whole blocks only: with a shorter trailing block the reference's C follower reads past the
            # end of its input (num_samples is fixed at block_size, detection.py:533-538)
            if block.shape[0] != B:
                return torch.zeros_like(block)
            block = block.contiguous()
            return fast(block).clone() - slow(block)

        for i in range(int(0.1 * sr), int(0.5 * sr), B):
            run(xd[i:i + B])
        rel = torch.zeros_like(xd)
        for i in range(0, len(xd), B):
            rel[i:i + B] = run(xd[i:i + B])
        self.mins = torch.quantile(rel[:sr], 0.5, dim=0).cpu().numpy()
        self.maxs = rel.max(0).values.cpu().numpy()
        self.on_threshold = self.maxs * self.on_threshold + self.mins
        self.off_threshold = self.maxs * self.off_threshold + self.mins
        w = int(sr * 0.01)
        pad_l, pad_r = w // 2, w - 1 - w // 2
        mf = torch.nn.functional.max_pool1d(
            torch.nn.functional.pad(rel.T[None], (pad_l, pad_r), mode="reflect"), w, stride=1)[0].T
        self.noise_max = torch.quantile(mf, 0.5, dim=0).cpu().numpy()
        xr = torch.flip(xd[:sr], dims=(0,)).contiguous()
        for i in range(0, sr, B):
            run(xr[i:i + B])
        # hand the settled filter / follower state back to the detector __call__ uses
        if hp is not None:
            zi = hp.zi.cpu().numpy()
            for i in range(4):
                st[f"z{i}"] = zi[i:i + 1].copy()
        st["yf"], st["ys"] = fast.y[-1:].cpu().numpy(), slow.y[-1:].cpu().numpy()
        self._det.load_state(st)


class ButterworthFilter:
    """detection.py:487-501: Butterworth filter applied to n signals in parallel, state carried between
    calls.  The coefficients come from scipy.signal.butter on the host exactly as in the reference
    (cast to float32); the filtering is ofp_lfilter (scipy's float32 direct form II transposed)."""

    def __init__(self, cutoff, n, order=2, sr=44100, btype="high"):
        from scipy import signal as sig

        torch = _lib.require_cuda()
        b, a = sig.butter(order, cutoff, btype=btype, analog=False, output="ba", fs=sr)
        self.b, self.a = np.float32(b), np.float32(a)
        self.order = len(self.b) - 1  # band filters double the order
        self.zi = torch.zeros((self.order, n), dtype=torch.float32, device="cuda")
        self.n = n

    def __call__(self, x):
        torch = _lib.require_cuda()
        xd = _to_dev(x, torch)
        assert xd.dim() == 2 and xd.shape[1] == self.n
        y = torch.empty_like(xd)
        b = (C.c_float * (self.order + 1))(*self.b.tolist())
        a = (C.c_float * (self.order + 1))(*self.a.tolist())
        check(_lib.lib().ofp_lfilter(b, a, C.c_int32(self.order), ptr(xd), ptr(y), ptr(self.zi),
                                     C.c_int32(xd.shape[0]), C.c_int32(self.n), stream_ptr()))
        return y.cpu().numpy() if isinstance(x, np.ndarray) else y


class AREnvelopeFollower:
    """detection.py:504-538 over ofp_ar_envelope (twin of envelope_follower.c:6-25)."""

    def __init__(self, x0, attack=3, release=383):
        torch = _lib.require_cuda()
        self.attack = np.float32(1 / attack)
        self.release = np.float32(1 / release)
        if isinstance(x0, np.ndarray):
            self.y = torch.from_numpy(np.ascontiguousarray(x0, dtype=np.float32)).cuda()
        else:
            self.y = x0.to(device="cuda", dtype=torch.float32).contiguous().clone()
        self.n, self.size = x0.shape

    def __call__(self, x):
        torch = _lib.require_cuda()
        xd = _to_dev(x, torch)
        check(_lib.lib().ofp_ar_envelope(ptr(xd), ptr(self.y), C.c_float(self.attack), C.c_float(self.release),
                                         C.c_int(self.size), C.c_int(self.n), stream_ptr()))
        return self.y.cpu().numpy() if isinstance(x, np.ndarray) else self.y


class MinMaxEnvelopeFollower:
    """detection.py:541-592 over ofp_minmax_envelope (twin of envelope_follower.c:27-57)."""

    def __init__(self, x0: np.ndarray, alpha_min=1e-5, alpha_max=1e-5, minmin=0.0):
        torch = _lib.require_cuda()
        self.alpha_min, self.alpha_max, self.minmin = np.float32(alpha_min), np.float32(alpha_max), np.float32(minmin)
        self.min_val = torch.from_numpy(np.float32(np.min(x0, axis=0))).cuda()
        self.max_val = torch.from_numpy(np.float32(np.max(x0, axis=0))).cuda()
        self.n_channels = x0.shape[1]

    def __call__(self, x):
        torch = _lib.require_cuda()
        xd = _to_dev(x, torch)
        check(_lib.lib().ofp_minmax_envelope(ptr(xd), ptr(self.min_val), ptr(self.max_val), C.c_float(self.alpha_min),
                                             C.c_float(self.alpha_max), C.c_float(self.minmin), C.c_int(len(x)),
                                             C.c_int(self.n_channels), stream_ptr()))
        if isinstance(x, np.ndarray):
            return self.min_val.cpu().numpy(), self.max_val.cpu().numpy()
        return self.min_val, self.max_val


def find_onset_groups(onsets, channels, max_distance: int = 1000, min_channels: int = 3,
                      close_channel: Optional[int] = None):
    """detection.py:131-189.  Sequential scan over (onset, channel) in detection order; the work is
    a few thousand integers per recording, so it stays on the host."""
    if len(onsets) == 0:
        raise ValueError("max() arg is an empty sequence")  # what the reference raises (line 158)
    width = max(channels) + 1
    rows, members = [], []

    def close_group():
        if len({ch for _, ch in members}) >= min_channels:
            row = np.full((width,), -1, dtype=int)
            for s, ch in members:
                row[ch] = s
            rows.append(row)

    for s, ch in zip(onsets, channels):
        if members and abs(s - members[0][0]) > max_distance:
            close_group()
            members = []
        members.append((s, ch))
    close_group()
    if close_channel is not None:
        rows = [r for r in rows if all(r[close_channel] <= r)]
    return np.array(rows, dtype=int) if rows else None


# ------------------------------------------------------------------------------------------------
# K3 / K4: grouping on the device, lag refinement
# ------------------------------------------------------------------------------------------------
LAG_NONE = -(2 ** 31)
_DIRECTION = {None: 0, "up": 1, "down": 2}


def find_onset_groups_batch(channels, onsets, counts, n_channels: int, max_distance: int = 1000,
                            min_channels: int = 3, close_channel: Optional[int] = None,
                            max_groups: Optional[int] = None, with_span: bool = False):
    """find_onset_groups for every recording of a batch on the device.
    channels/onsets [R, cap] int32 and counts [R] int32 as returned by detect_onsets_amplitude_batch.
    Returns (hit_rec [H] int32, hit_onsets [H, C] int32, n_groups [R] int32): the groups of all
    recordings concatenated in recording order (rows may contain -1 for a missing channel).
    with_span: also return the largest onset spread over the complete groups (what sizes the sections of
    fix_onsets_batch) -- it rides on the same host round trip as the hit count, so a pass has ONE."""
    torch = _lib.require_cuda()
    R, cap = channels.shape
    if max_groups is None:
        max_groups = max(1, cap // max(min_channels, 1))
    groups = torch.empty((R, max_groups, n_channels), dtype=torch.int32, device="cuda")
    ng = torch.empty((R,), dtype=torch.int32, device="cuda")
    check(_lib.lib().ofp_group_onsets(ptr(channels), ptr(onsets), ptr(counts), C.c_int32(R), C.c_int32(cap),
                                      C.c_int32(n_channels), C.c_int32(max_distance), C.c_int32(min_channels),
                                      C.c_int32(-1 if close_channel is None else close_channel),
                                      C.c_int32(max_groups), ptr(groups), ptr(ng), stream_ptr()))
    kept = torch.clamp(ng, max=max_groups).to(torch.int64)
    offsets = torch.cumsum(kept, 0) - kept
    span = 0
    if with_span:
        # element-wise min / max over the channel columns (a torch reduction over a trailing dimension of 3 costs
        # milliseconds on 10^7 elements; C - 1 strided maximum / minimum kernels cost microseconds)
        mn = mx = groups[:, :, 0]
        for c in range(1, n_channels):
            mn, mx = torch.minimum(mn, groups[:, :, c]), torch.maximum(mx, groups[:, :, c])
        valid = torch.arange(max_groups, device="cuda")[None, :] < kept[:, None]
        spread = torch.where(valid & (mn >= 0), mx - mn, torch.zeros_like(mx)).max().to(torch.int64)
        H, span = (int(v) for v in torch.stack([kept.sum(), spread]).tolist())
    else:
        H = int(kept.sum().item())
    hit_rec = torch.empty((H,), dtype=torch.int32, device="cuda")
    hit_on = torch.empty((H, n_channels), dtype=torch.int32, device="cuda")
    if H:
        check(_lib.lib().ofp_compact_groups(ptr(groups), ptr(ng), ptr(offsets), C.c_int32(R), C.c_int32(max_groups),
                                            C.c_int32(n_channels), ptr(hit_rec), ptr(hit_on), stream_ptr()))
    if with_span:
        return hit_rec, hit_on, ng, span
    return hit_rec, hit_on, ng


def section_budget(max_section: int, n_channels: int) -> int:
    """Section length a K4 CTA can hold: longer sections are flagged OFP_FIX_TOO_LONG.  When the [L, C] section does
    not fit, the kernel keeps two channel columns instead (column mode): 16 + 8 + 8 + 8 B per sample."""
    budget = max((200 * 1024 - 128) // (16 + 8 * n_channels), (200 * 1024 - 16 * 1024) // 40)
    return min(int(max_section), budget)


def fix_onsets_batch(audio, hit_rec, hit_onsets, filter_size: int = 5, d: int = 0, onset_direction=None,
                     take_abs: bool = False, zero_left: bool = False, normalization_cutoff: int = 10,
                     onset_tolerance: int = 30, shift_onsets: int = 0, max_section: Optional[int] = None,
                     to_end: bool = False):
    """fix_onsets (detection.py:373-451) for H onset groups in one launch.
    audio [R, N, C] float32 (device or numpy); hit_rec [H] int32 or None (hit h in recording h);
    hit_onsets [H, C] int32.  Returns (onsets [H, C], lags [H, C], status [H]) device tensors.
    to_end: sections run to the end of the recording (the ring-buffer sections of
    Multilaterate3D.locate, multilateration.py:457-466)."""
    torch = _lib.require_cuda()
    audio = _to_dev(audio, torch)
    R, N, Cn = audio.shape
    hit_onsets = hit_onsets.to(device="cuda", dtype=torch.int32).contiguous()
    H = hit_onsets.shape[0]
    if hit_rec is not None:
        hit_rec = hit_rec.to(device="cuda", dtype=torch.int32).contiguous()
    look = normalization_cutoff + onset_tolerance
    if max_section is None:
        span = 0
        if H:  # largest onset spread over the complete groups: one small reduction, one host round trip
            mn, mx = hit_onsets.min(1).values, hit_onsets.max(1).values
            span = int(torch.where(mn >= 0, mx - mn, torch.zeros_like(mx)).max().item())
        max_section = section_budget(N if to_end else span + 2 * look + 1, Cn)
    out = torch.empty_like(hit_onsets)
    lags = torch.empty_like(hit_onsets)
    status = torch.empty((H,), dtype=torch.int32, device="cuda")
    check(_lib.lib().ofp_fix_onsets_ex(ptr(audio), C.c_int64(N), C.c_int64(audio.stride(0)), C.c_int32(Cn),
                                       ptr(hit_rec), ptr(hit_onsets), C.c_int32(H), C.c_int32(filter_size),
                                       C.c_int32(d), C.c_int32(_DIRECTION[onset_direction]),
                                       C.c_int32(bool(take_abs)), C.c_int32(bool(zero_left)),
                                       C.c_int32(normalization_cutoff), C.c_int32(onset_tolerance),
                                       C.c_int32(shift_onsets), C.c_int32(max_section), C.c_int32(int(to_end)),
                                       ptr(out), ptr(lags), ptr(status), stream_ptr()))
    return out, lags, status


def max_onsets_batch(audio, hit_rec, hit_onsets, tolerance: int):
    """The peak refinement of the notebooks' dataset builder (notebooks/refresh.org:262-279):
    ``og[i] + np.argmax(audio[og[i] : og[i] + tolerance, i])`` for every hit and channel in one launch.
    audio [R, N, C] (device or numpy), hit_rec [H] or None (hit h in recording h), hit_onsets [H, C] int32."""
    torch = _lib.require_cuda()
    audio = _to_dev(audio, torch)
    R, N, Cn = audio.shape
    hit_onsets = hit_onsets.to(device="cuda", dtype=torch.int32).contiguous()
    if hit_rec is not None:
        hit_rec = hit_rec.to(device="cuda", dtype=torch.int32).contiguous()
    out = torch.empty_like(hit_onsets)
    check(_lib.lib().ofp_window_argmax(ptr(audio), C.c_int64(N), C.c_int64(audio.stride(0)), C.c_int32(Cn), ptr(hit_rec),
                                       ptr(hit_onsets), C.c_int32(hit_onsets.shape[0]), C.c_int32(int(tolerance)),
                                       ptr(out), stream_ptr()))
    return out


def fix_onsets(audio: np.ndarray, onsets: np.ndarray, filter_size: int = 5, d: int = 0, onset_direction=None,
               take_abs: bool = False, zero_left: bool = False, normalization_cutoff: int = 10,
               onset_tolerance: int = 30, shift_onsets: int = 0, return_status: bool = False):
    """Drop-in for detection.fix_onsets (detection.py:373-451): audio [N, C], onsets [G, C] -> [G, C].
    Where the reference raises (ValueError in adjust_onset; negative section start, Q6) this
    raises ValueError too, unless return_status=True, in which case (onsets, status, lags) is returned
    with the affected groups flagged."""
    torch = _lib.require_cuda()
    on = torch.from_numpy(np.ascontiguousarray(onsets).astype(np.int32))
    rec = torch.zeros(len(onsets), dtype=torch.int32)
    out, lags, status = fix_onsets_batch(audio[None], rec, on, filter_size, d, onset_direction, take_abs, zero_left,
                                         normalization_cutoff, onset_tolerance, shift_onsets)
    out, lags, status = out.cpu().numpy().astype(np.int64), lags.cpu().numpy(), status.cpu().numpy()
    if return_status:
        return out, status, lags
    if (status != 0).any():
        bad = np.nonzero(status)[0]
        raise ValueError(f"fix_onsets: the reference raises on groups {bad.tolist()} (status {status[bad].tolist()})")
    return out


def cross_correlation_lag(x: np.ndarray, y: np.ndarray, onsets=None, legal_lags=None, d: int = 0,
                          normalization_cutoff: int = 10, onset_tolerance: int = 50, take_abs: bool = False):
    """Drop-in for detection.cross_correlation_lag (detection.py:195-268) -> int or None."""
    torch = _lib.require_cuda()
    if len(x) - d < 1 or len(y) - d < 1:  # np.correlate refuses empty operands (detection.py:232)
        raise ValueError("first array argument cannot be empty")
    xd = _to_dev(np.ascontiguousarray(x, dtype=np.float32)[None], torch)
    yd = _to_dev(np.ascontiguousarray(y, dtype=np.float32)[None], torch)
    use_legal = legal_lags is not None
    pair = legal_lags if use_legal else (onsets if onsets is not None else None)
    if pair is None:
        raise UnboundLocalError("max_adjust")  # what the reference does without onsets / legal_lags
    o = torch.tensor([[int(pair[0]), int(pair[1])]], dtype=torch.int32, device="cuda")
    out = torch.empty((1,), dtype=torch.int32, device="cuda")
    check(_lib.lib().ofp_cross_correlation_lag(ptr(xd), ptr(yd), C.c_int32(1), C.c_int32(xd.shape[1]), C.c_int32(d),
                                               C.c_int32(bool(take_abs)), C.c_int32(use_legal),
                                               C.c_int32(normalization_cutoff), C.c_int32(onset_tolerance), ptr(o),
                                               ptr(out), stream_ptr()))
    r = int(out.item())
    return None if r == LAG_NONE else r


def adjust_onset(onsets, x: np.ndarray, y: np.ndarray, new_lag: int):
    """Drop-in for detection.adjust_onset (detection.py:299-352) -> (change_x, change_y)."""
    torch = _lib.require_cuda()
    xd = _to_dev(np.ascontiguousarray(x, dtype=np.float32)[None], torch)
    yd = _to_dev(np.ascontiguousarray(y, dtype=np.float32)[None], torch)
    o = torch.tensor([[int(onsets[0]), int(onsets[1])]], dtype=torch.int32, device="cuda")
    nl = torch.tensor([int(new_lag)], dtype=torch.int32, device="cuda")
    out = torch.empty((1, 2), dtype=torch.int32, device="cuda")
    check(_lib.lib().ofp_adjust_onset(ptr(xd), ptr(yd), C.c_int32(1), C.c_int32(xd.shape[1]), ptr(o), ptr(nl),
                                      ptr(out), stream_ptr()))
    a, b = out[0].cpu().tolist()
    if a == LAG_NONE:
        raise ValueError("operands could not be broadcast together (adjust_onset, detection.py:335)")
    return a, b


def adjust_onset_rel(onsets, relx: np.ndarray, rely: np.ndarray, new_lag: int):
    """detection.py:271-296: move the onset whose relative envelope rises more towards the target lag.
    Four envelope reads and one comparison per pair -- host arithmetic on whatever array type the
    envelopes are (numpy, or device tensors straight from detect_onsets_amplitude_batch)."""
    oa, ob = onsets[0], onsets[1]
    lag_diff = (ob - oa) - new_lag
    da = relx[oa + lag_diff] - relx[oa]
    db = rely[ob - lag_diff] - rely[ob]
    if da > db:
        oa += lag_diff
    else:
        ob -= lag_diff
    return oa, ob


def filter_data(x, direction: str):
    """detection.py:355-370: zero the samples whose first difference (along axis 0) has the wrong sign,
    in place, and return x (numpy array or device tensor, [N] or [N, C])."""
    if direction not in ("up", "down"):
        raise RuntimeError(f"Unknown onset direction {direction=}!")
    torch = _lib.require_cuda()
    is_np = isinstance(x, np.ndarray)
    xd = _to_dev(x, torch)
    n = xd.shape[0]
    cn = int(xd.numel() // max(n, 1))
    out = torch.empty_like(xd)
    check(_lib.lib().ofp_filter_data(ptr(xd), ptr(out), C.c_int64(n), C.c_int32(cn),
                                     C.c_int32(_DIRECTION[direction]), stream_ptr()))
    if is_np:
        x[...] = out.cpu().numpy()
    else:
        x.copy_(out)
    return x


def detect_onset_region_batch(audio, detected_onsets, n: int = 256, median_filter_size: int = 5,
                              threshold_factor: float = 0.5):
    """detect_onset_region for P signals at once: audio [P, len], detected_onsets [P] -> int32 [P] tensor."""
    torch = _lib.require_cuda()
    ad = _to_dev(audio, torch)
    on = torch.as_tensor(np.asarray(detected_onsets, np.int32)).to(device="cuda", dtype=torch.int32).contiguous()
    out = torch.empty_like(on)
    check(_lib.lib().ofp_detect_onset_region(ptr(ad), C.c_int32(ad.shape[0]), C.c_int64(ad.shape[1]), ptr(on),
                                             C.c_int32(n), C.c_int32(median_filter_size),
                                             C.c_float(threshold_factor), ptr(out), stream_ptr()))
    return out


def detect_onset_region(audio, detected_onset, n=256, median_filter_size=5, threshold_factor=0.5):
    """detection.py:454-484: start of the loud part around a detected onset in a 1-D signal."""
    a = np.ascontiguousarray(audio, dtype=np.float32)[None] if isinstance(audio, np.ndarray) else audio[None]
    return int(detect_onset_region_batch(a, [int(detected_onset)], n, median_filter_size, threshold_factor)[0].item())
