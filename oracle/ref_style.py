"""Reference-shaped CPU implementation of the onset detector, for TIMING the reference's CPU path.

TEST / BENCH INFRASTRUCTURE (bench.py --impl reference and the cpu_baseline leg only).

The reference's Python files cannot travel to the GPU box, its C kernels can: this module
drives the reference's own compiled DLL (oracle/_ref/envelope_follower.so, built from
/root/reference/onset_fingerprinting/envelope_follower.c with the reference's flags) from a
numpy/scipy block loop of the same granularity as detection.py:19-86 / 727-798 -- one
``lfilter`` call, ~20 numpy ufunc dispatches and three ctypes calls per 128-sample block -- so
its throughput is what the reference achieves on the same host.  When _ref is absent it
falls back to the oracle's own C twins of those two kernels (kind = "port").
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np
from scipy import signal as sig

HERE = Path(__file__).resolve().parent
_REF = HERE / "_ref" / "envelope_follower.so"

_f2 = np.ctypeslib.ndpointer(dtype=np.float32, ndim=2, flags="C_CONTIGUOUS")
_f1 = np.ctypeslib.ndpointer(dtype=np.float32, ndim=1, flags="C_CONTIGUOUS")


def kind() -> str:
    return "reference" if _REF.exists() else "port"


def _dll():
    if _REF.exists():
        d = C.CDLL(str(_REF))
        ar, mm = d.ar_envelope, d.minmax_envelope
    else:
        from . import oracle

        d = oracle.lib()
        ar, mm = d.orc_ar_envelope, d.orc_minmax_envelope
    ar.argtypes = [_f2, _f2, C.c_float, C.c_float, C.c_int, C.c_int]
    mm.argtypes = [_f2, _f1, _f1, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
    return ar, mm


class BlockDetector:
    """Block-wise detector with the reference's structure (detection.py:595-840)."""

    def __init__(self, n_signals, block_size=128, floor=-70.0, hipass_freq=2000.0, fast_ar=(3.0, 383.0),
                 slow_ar=(2205.0, 2205.0), on_threshold=0.5, off_threshold=0.1, cooldown=1323, sr=96000):
        self.ar, self.mm = _dll()
        self.B, self.C, self.floor = block_size, n_signals, floor
        self.on, self.off, self.manual, self.cooldown = on_threshold, off_threshold, on_threshold > 1, cooldown
        self.hp = None
        if hipass_freq != 0:
            b, a = sig.butter(4, hipass_freq, btype="high", output="ba", fs=sr)
            self.hp = (np.float32(b), np.float32(a))
            self.zi = np.zeros((4, n_signals), np.float32)
        self.fa, self.fr = np.float32(1 / fast_ar[0]), np.float32(1 / fast_ar[1])
        self.sa, self.sr_ = np.float32(1 / slow_ar[0]), np.float32(1 / slow_ar[1])
        self.yf = np.full((block_size, n_signals), floor, np.float32)
        self.ys = np.full((block_size, n_signals), floor, np.float32)
        self.mn = np.zeros(n_signals, np.float32)
        self.mx = np.full(n_signals, 10, np.float32)
        self.state = np.zeros(n_signals, bool)
        self.prev = np.zeros(n_signals)
        self.deb = np.zeros(n_signals, int)

    def _filter(self, x):
        if self.hp is None:
            return x
        y, self.zi = sig.lfilter(self.hp[0], self.hp[1], x, axis=0, zi=self.zi)
        return y

    def _envelope(self, db):
        self.ar(db, self.yf, self.fa, self.fr, self.C, self.B)
        self.ar(db, self.ys, self.sa, self.sr_, self.C, self.B)
        rel = 10 ** ((self.yf - self.ys) / 20) - 1e-10
        return rel.clip(0, -self.floor)

    def warmup(self, x):
        x = self._filter(x)
        db = (20 * np.log10(np.abs(x + 1e-10))).clip(self.floor)
        for i in range(0, len(db) - self.B + 1, self.B):
            rel = self._envelope(np.ascontiguousarray(db[i:i + self.B]))
            self.mm(rel, self.mn, self.mx, 1e-4, 1e-5, 2.0, self.B, self.C)

    def __call__(self, x):
        x = self._filter(x)
        db = (20 * np.log10(np.abs(x + 1e-10))).clip(self.floor)
        rel = self._envelope(np.ascontiguousarray(db))
        if self.manual:
            thr_on, thr_off = self.on, self.off
        else:
            self.mm(rel, self.mn, self.mx, 1e-4, 1e-5, 2.0, self.B, self.C)
            thr_on, thr_off = self.mx * self.on + self.mn, self.mx * self.off + self.mn
        cross = (rel > thr_on) & ~self.state & (self.deb < 1)
        cross[0] &= self.prev < thr_on
        cross[1:] &= rel[:-1] < thr_on
        first = np.argmax(cross, axis=0)
        hit = (first > 0) | cross[0]
        self.state[hit] = True
        self.deb[hit] = self.cooldown
        self.deb[self.deb > 0] -= self.B
        below = rel < thr_off
        below[:first.max()] = False
        self.state[below.any(axis=0)] = False
        self.prev[:] = rel[-1]
        return np.where(hit)[0], first[hit], rel


def detect_onsets_amplitude(x, block_size=128, sr=96000, **kw):
    """Same contract as detection.detect_onsets_amplitude (detection.py:19-86)."""
    od = BlockDetector(x.shape[1], block_size, sr=sr, **kw)
    od.warmup(x[: int(0.5 * sr)])
    chans, onsets, rels = [], [], []
    for i in range(0, len(x) - block_size + 1, block_size):
        c, d, r = od(x[i:i + block_size])
        rels.append(r)
        chans += c.tolist()
        onsets += (i + d).tolist()
    return chans, onsets, np.concatenate(rels) if rels else np.zeros((0, x.shape[1]), np.float32)
