"""Margin audit of the oracle's decisions (SURVEY.md H2 iii).  TEST INFRASTRUCTURE.

Bit-exact onset indices and lags hinge on comparisons (``rel > thr_on``, ``rel < thr_off``,
``argmax(cc)``, ``da > db``).  The CUDA path reproduces the oracle's arithmetic operation by
operation, but the oracle itself pins two things the reference leaves to the platform (float32
log10 / 10**x results, np.correlate's summation order).  This module measures how far every
decision of a run is from flipping, so that the platform-dependent part (a few float32 ulps,
~1e-7 relative; CC sums ~1e-7 relative) can be shown to be irrelevant on the data the numbers
are quoted on -- or the ambiguous decisions get listed, not hidden.

    detector_margins(x, **detector_kw)  -> dict   (detection.py:759-792 decisions)
    fix_margins(audio, groups, **fix_kw) -> dict  (detection.py:195-268, 299-352 decisions)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle as orc


def detector_margins(x: np.ndarray, block_size: int = 128, sr: int = 96000, **kw) -> dict:
    """Replays the oracle's detector block by block (same C code as detect_onsets_amplitude) and
    records, relative to the threshold in force:
      onset_at / onset_before : |rel - thr_on| / thr_on at every crossing sample and the one before it
      armed_closest           : the closest any sample of an ARMED channel (state off, debounce over)
                                that did not produce an onset came to thr_on (a near miss would be as
                                ambiguous as a near hit)
      off_closest             : per block and active channel, how close the decisive sample of the
                                off test (the block minimum from the first onset row on) came to thr_off
    """
    x = np.ascontiguousarray(x, np.float32)
    n, ch = x.shape
    det = orc.Detector(ch, block_size, sr=sr, **kw)
    det.init_minmax_tracker(x[: int(0.5 * sr)])
    p = det.p
    on_at, on_before, armed, off = [], [], [], []
    onsets, channels = [], []
    prev = np.array([det.st[c].prev for c in range(ch)], np.float32)
    for b, i in enumerate(range(0, n - block_size + 1, block_size)):
        state0 = np.array([det.st[c].state for c in range(ch)])
        deb0 = np.array([det.st[c].deb for c in range(ch)])
        cs, ds, rel = det(x[i:i + block_size])
        mn = np.array([det.st[c].mn for c in range(ch)], np.float32)
        mx = np.array([det.st[c].mx for c in range(ch)], np.float32)
        if p.manual:
            thr_on = np.full(ch, p.on_thr, np.float32)
            thr_off = np.full(ch, p.off_thr, np.float32)
        else:
            thr_on = mx * np.float32(p.on_thr) + mn
            thr_off = mx * np.float32(p.off_thr) + mn
        hit = dict(zip(cs.tolist(), ds.tolist()))
        M = max(hit.values()) if hit else 0
        for c in range(ch):
            col = rel[:, c].astype(np.float64)
            t_on, t_off = float(thr_on[c]), float(thr_off[c])
            if c in hit:
                k = hit[c]
                before = float(prev[c]) if k == 0 else col[k - 1]
                on_at.append(abs(col[k] - t_on) / t_on)
                on_before.append(abs(before - t_on) / t_on)
                onsets.append(i + k)
                channels.append(c)
                # samples before the crossing also had to fail the test
                seq = np.concatenate([[float(prev[c])], col[:k]])
            elif state0[c] == 0 and deb0[c] < 1:
                seq = np.concatenate([[float(prev[c])], col])
            else:
                seq = None
            if seq is not None and len(seq) > 1 and t_on > 0:
                # a crossing at k needs seq[k+1] > thr and seq[k] < thr: its distance from happening is the
                # larger of the two shortfalls; here simply the closest approach of any sample to thr_on
                armed.append(float(np.min(np.abs(seq - t_on)) / t_on))
            # off test: rows M.. of the block against thr_off, relevant while the channel is on
            if (state0[c] == 1 or c in hit) and t_off > 0:
                off.append(float(abs(col[M:].min() - t_off) / t_off))
        prev = rel[-1].copy()
    a = lambda v: np.asarray(v, np.float64)
    return {"onsets": np.asarray(onsets), "channels": np.asarray(channels), "onset_at": a(on_at),
            "onset_before": a(on_before), "armed_closest": a(armed), "off_closest": a(off)}


_DIR = {None: 0, "up": 1, "down": 2}


def fix_margins(audio: np.ndarray, groups: np.ndarray, filter_size=5, d=0, onset_direction=None, take_abs=False,
                zero_left=False, normalization_cutoff=10, onset_tolerance=30, shift_onsets=0) -> dict:
    """Per evaluated channel pair of fix_onsets: the relative gap between the winning cross-correlation
    value and the best value at any other lag, and the relative gap between adjust_onset's da and db."""
    audio = np.ascontiguousarray(audio, np.float32)
    out = np.array(groups, dtype=np.int64) + shift_onsets
    L = orc.lib()
    cc_gap, ab_gap, status = [], [], []
    for j in range(len(out)):
        row = np.ascontiguousarray(out[j])
        lags = np.empty(out.shape[1], np.int32)
        audit = np.empty((out.shape[1], 4), np.float64)
        st = L.orc_fix_group_audit(audio.ctypes.data_as(C.c_void_p), C.c_int64(audio.shape[0]), C.c_int(audio.shape[1]),
                                   row.ctypes.data_as(C.c_void_p), filter_size, d, _DIR[onset_direction], int(take_abs),
                                   int(zero_left), normalization_cutoff, onset_tolerance,
                                   lags.ctypes.data_as(C.c_void_p), audit.ctypes.data_as(C.c_void_p))
        status.append(st)
        for c in range(out.shape[1]):
            top, second, da, db = audit[c]
            if np.isfinite(top) and np.isfinite(second):
                cc_gap.append((top - second) / max(abs(top), 1e-300))
            if np.isfinite(da) and np.isfinite(db) and (da != 0 or db != 0):
                ab_gap.append(abs(da - db) / max(abs(da), abs(db)))
    return {"cc_gap": np.asarray(cc_gap), "ab_gap": np.asarray(ab_gap), "status": np.asarray(status)}


def histogram(v: np.ndarray, edges=(0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, np.inf)) -> str:
    v = np.asarray(v)
    h, _ = np.histogram(v, bins=np.asarray(edges))
    cells = [f"<{e:g}: {c}" for e, c in zip(edges[1:], h)]
    return f"n={len(v)} min={v.min() if len(v) else float('nan'):.3g} | " + ", ".join(cells)
