"""The whole hot path for a batch of recordings in device memory:

    detect_onsets_amplitude (K1) -> find_onset_groups (K3) -> fix_onsets (K4) -> Multilaterate3D (K5)

which is what the reference's notebooks run per recording in Python (notebooks/refresh.org:149-172,
1507-1509) and what PlayRec.detect_hits runs per block (realtime/audio.py:62-74).  Everything stays
on the device; the only host round trip is the hit count (one integer) that sizes the K4/K5 launches.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from . import detection, multilateration


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

    def run(self, x, return_rel: bool = False, warm_n: Optional[int] = None) -> HitBatch:
        torch = self.det.torch
        R, N, C = x.shape
        if warm_n is None:
            warm_n = int(0.5 * self.sr)
        if self._out is None or self._out[0].shape[0] != R or (self._out[3] is None) == return_rel:
            cap = self.det.default_cap(N)
            nb = N // self.det.block_size
            self._out = (torch.empty((R, cap), dtype=torch.int32, device="cuda"),
                         torch.empty((R, cap), dtype=torch.int32, device="cuda"),
                         torch.empty((R,), dtype=torch.int32, device="cuda"),
                         torch.empty((R, nb * self.det.block_size, C), dtype=torch.float32, device="cuda")
                         if return_rel else None)
        self.det.reset()
        ch, ix, cnt, rel = self.det.detect_offline(x, warm_n, out=self._out)
        hit_rec, hit_on, _ = detection.find_onset_groups_batch(ch, ix, cnt, C, **self.group_kw)
        # sections are sized by the largest onset spread actually present (one tiny reduction + host round
        # trip): the shared memory of a K4 CTA -- and with it how many hits an SM works on -- follows it
        fixed, lags, fstat = detection.fix_onsets_batch(x, hit_rec, hit_on, **self.fix_kw)
        xy, lstat = self.ml.locate_batch(fixed)
        return HitBatch(hit_rec, hit_on, fixed, lags, fstat, xy, lstat, cnt, rel)
