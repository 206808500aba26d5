"""The whole hot path for a batch of recordings in device memory:

    detect_onsets_amplitude (K1) -> find_onset_groups (K3) -> fix_onsets (K4) -> Multilaterate3D (K5)

which is what the reference's notebooks run per recording in Python (notebooks/refresh.org:149-172,
1507-1509) and what PlayRec.detect_hits runs per block (realtime/audio.py:62-74).  Everything stays
on the device; the only host round trip is ONE pair of integers (hit count and largest onset spread) that
sizes the K4/K5 launches.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

from . import _lib, detection, multilateration


@dataclass
class HitBatch:
    """Per-hit records of a batch (device tensors)."""

    rec: "torch.Tensor"         # [H] int32 recording index
    onsets: "torch.Tensor"      # [H, C] int32 detected onsets (K1/K3)
    fixed: "torch.Tensor"       # [H, C] int32 onsets after lag refinement (K4)
    lags: "torch.Tensor"        # [H, C] int32 lag returned by the CC per later channel
    fix_status: "torch.Tensor"  # [H] int32 OFP_FIX_*
    xy: "torch.Tensor"          # [H, 2] float64 cm, NaN when not located (K5)
    loc_status: "torch.Tensor"  # [H] int32
    onset_counts: "torch.Tensor"  # [R] int32 onsets per recording
    rel: Optional["torch.Tensor"]  # [R, nb*B, C] float32 or None


class HotPath:
    """Reusable state for repeated passes over batches of the same shape."""

    def __init__(self, n_rec: int, n_channels: int, sensors, medium: str = "air", sr: int = 96000,
                 block_size: int = 128, detector_kw: Optional[dict] = None, group_kw: Optional[dict] = None,
                 fix_kw: Optional[dict] = None):
        self.sr, self.n_rec, self.n_channels = sr, n_rec, n_channels
        self.det = detection.BatchedOnsetDetector(n_rec, n_channels, block_size, sr=sr, **(detector_kw or {}))
        self.ml = multilateration.Multilaterate3D(sensors, sr=sr, medium=medium)
        self.group_kw = dict(max_distance=1000, min_channels=n_channels)
        self.group_kw.update(group_kw or {})
        self.fix_kw = dict(fix_kw or {})
        self._out = None
        self._out_key = None
        self._host = None

    def _ensure_out(self, R, N, C, return_rel):
        """Onset / envelope buffers of a pass, kept between passes of the same shape (R, N and rel mode)."""
        torch = self.det.torch
        key = (R, N, C, bool(return_rel))
        if self._out is None or self._out_key != key:
            cap = self.det.default_cap(N)
            nb = N // self.det.block_size
            self._out = (torch.empty((R, cap), dtype=torch.int32, device="cuda"),
                         torch.empty((R, cap), dtype=torch.int32, device="cuda"),
                         torch.empty((R,), dtype=torch.int32, device="cuda"),
                         torch.empty((R, nb * self.det.block_size, C), dtype=torch.float32, device="cuda")
                         if return_rel else None)
            self._out_key = key
        return self._out

    def run(self, x, return_rel: bool = False, warm_n: Optional[int] = None, k1_events=None) -> HitBatch:
        """x [R, N, C] device tensor.  k1_events: optional (start, end) CUDA events recorded around the detector
        launch on the current stream (bench.py's roofline leg)."""
        R, N, C = x.shape
        if warm_n is None:
            warm_n = int(0.5 * self.sr)
        self._ensure_out(R, N, C, return_rel)
        self.det.reset()
        if k1_events is not None:
            k1_events[0].record()
        ch, ix, cnt, rel = self.det.detect_offline(x, warm_n, out=self._out)
        if k1_events is not None:
            k1_events[1].record()
        hit_rec, hit_on, _, span = detection.find_onset_groups_batch(ch, ix, cnt, C, with_span=True, **self.group_kw)
        # sections are sized by the largest onset spread actually present (the shared memory of a K4 CTA -- and
        # with it how many hits an SM works on -- follows it); the spread comes back with the hit count
        fixed, lags, fstat = self._fix(x, hit_rec, hit_on, span)
        xy, lstat = self.ml.locate_batch(fixed)
        return HitBatch(hit_rec, hit_on, fixed, lags, fstat, xy, lstat, cnt, rel)

    def _fix(self, x, hit_rec, hit_on, span):
        kw = dict(self.fix_kw)
        if "max_section" not in kw and not kw.get("to_end", False):
            look = kw.get("normalization_cutoff", 10) + kw.get("onset_tolerance", 30)
            kw["max_section"] = detection.section_budget(span + 2 * look + 1, x.shape[2])
        return detection.fix_onsets_batch(x, hit_rec, hit_on, **kw)

    # -- host buffers in, host results out -------------------------------------------------------------
    def run_host(self, x_host, rel_host=None, x_dev=None, warm_n: Optional[int] = None, segment: int = 49152):
        """The same pass for a batch that lives in (pinned) HOST memory: x_host [R, N, C] float32 torch tensor.

        The batch is uploaded in TIME segments of all recordings into a resident device copy (lag refinement reads
        its sections from it later); the detector resumes from its state after every segment
        (ofp_detect_continue), so the kernel of segment s runs under the copy of segment s+1 and, in drop-in mode
        (rel_host given), the envelope of segment s returns to the host under both.  Grouping, lag refinement
        and multilateration follow on the device; the per-hit records return to pinned host memory.
        Returns a dict of host tensors (rec, onsets, fixed, lags, fix_status, xy, loc_status, onset_counts);
        rel_host is filled in place.  x_dev: optional preallocated [R, N, C] device buffer to upload into."""
        torch = self.det.torch
        R, N, Cn = x_host.shape
        B = self.det.block_size
        if warm_n is None:
            warm_n = int(0.5 * self.sr)
        nb = N // B
        ch, ix, cnt, rel = self._ensure_out(R, N, Cn, rel_host is not None)
        if x_dev is None:
            if self._host is None or tuple(self._host.shape) != (R, N, Cn):
                self._host = torch.empty((R, N, Cn), dtype=torch.float32, device="cuda")
            x_dev = self._host
        if not hasattr(self, "_streams"):
            self._streams = (torch.cuda.Stream(), torch.cuda.Stream())
        s_copy, s_back = self._streams
        cur = torch.cuda.current_stream()
        s_copy.wait_stream(cur)
        s_back.wait_stream(cur)
        seg = max(int(segment), min(warm_n, N))
        seg = (seg + B - 1) // B * B
        if seg % 4:
            seg *= 4  # 16-byte rows for the TMA path
        self.det.reset()
        L = _lib.lib()
        t0 = 0
        first = True
        while t0 < N:
            ln = min(seg, N - t0)
            if t0 + ln < N and N - (t0 + ln) < B:
                ln = N - t0  # the trailing partial block rides with the last segment (only the warm-up reads it)
            with torch.cuda.stream(s_copy):  # one strided copy: this time segment of every recording
                _lib.check(L.ofp_copy2d_async(C.c_void_p(x_dev.data_ptr() + 4 * t0 * Cn), C.c_size_t(4 * x_dev.stride(0)),
                                              C.c_void_p(x_host.data_ptr() + 4 * t0 * Cn), C.c_size_t(4 * x_host.stride(0)),
                                              C.c_size_t(4 * ln * Cn), C.c_size_t(R), C.c_int(0), _lib.stream_ptr()))
                ev = torch.cuda.Event()
                ev.record()
            cur.wait_event(ev)
            blocks = ln // B
            view = x_dev[:, t0:t0 + ln]
            rview = rel[:, t0:t0 + blocks * B] if rel is not None else None
            if first:
                _lib.check(L.ofp_detect_offline(self.det._h, _lib.ptr(view), C.c_int64(ln), C.c_int64(x_dev.stride(0)),
                                                C.c_int64(warm_n), _lib.ptr(rview), C.c_int64(rel.stride(0) if rel is not None else 0),
                                                _lib.ptr(ch), _lib.ptr(ix), _lib.ptr(cnt), C.c_int32(ch.shape[1]),
                                                _lib.stream_ptr()))
                first = False
            elif blocks:
                _lib.check(L.ofp_detect_continue(self.det._h, _lib.ptr(view), C.c_int64(blocks * B),
                                                 C.c_int64(x_dev.stride(0)), C.c_int64(t0 // B), _lib.ptr(rview),
                                                 C.c_int64(rel.stride(0) if rel is not None else 0), _lib.ptr(ch),
                                                 _lib.ptr(ix), _lib.ptr(cnt), C.c_int32(ch.shape[1]), _lib.stream_ptr()))
            if rel_host is not None and blocks:
                done = torch.cuda.Event()
                done.record()
                with torch.cuda.stream(s_back):
                    s_back.wait_event(done)
                    _lib.check(L.ofp_copy2d_async(C.c_void_p(rel_host.data_ptr() + 4 * t0 * Cn), C.c_size_t(4 * rel_host.stride(0)),
                                                  C.c_void_p(rel.data_ptr() + 4 * t0 * Cn), C.c_size_t(4 * rel.stride(0)),
                                                  C.c_size_t(4 * blocks * B * Cn), C.c_size_t(R), C.c_int(1), _lib.stream_ptr()))
            t0 += ln
        hit_rec, hit_on, _, span = detection.find_onset_groups_batch(ch, ix, cnt, Cn, with_span=True, **self.group_kw)
        fixed, lags, fstat = self._fix(x_dev, hit_rec, hit_on, span)
        xy, lstat = self.ml.locate_batch(fixed)
        out = {}
        for name, t in (("rec", hit_rec), ("onsets", hit_on), ("fixed", fixed), ("lags", lags), ("fix_status", fstat),
                        ("xy", xy), ("loc_status", lstat), ("onset_counts", cnt)):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            out[name] = h
        cur.wait_stream(s_back)
        cur.synchronize()
        return out
