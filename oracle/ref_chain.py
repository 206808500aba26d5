"""The stages AFTER the detector at the reference's own granularity, for TIMING the reference's CPU path
(bench.py --impl reference and the cpu_baseline leg) and for cross-checking the C oracle.

TEST / BENCH INFRASTRUCTURE.  The reference's .py files cannot travel to the GPU box (they live under
/root/reference, which is absent there), so the stages after the detector -- find_onset_groups
(detection.py:131-189), fix_onsets / cross_correlation_lag / adjust_onset (detection.py:195-451) and
Multilaterate3D.locate / trilaterate / solve_trilateration_3d (multilateration.py:230-316, 397-566) -- are
restated here with the SAME library calls at the SAME granularity the reference uses: one
scipy.ndimage.median_filter + np.diff per hit, one np.correlate(..., "full") + np.argmax per channel pair,
np.exp(np.linspace) windows, one scipy.optimize.fsolve (MINPACK hybrj through Python callbacks) per hit.
Its per-hit cost is therefore the reference's (0.29 ms fix + 0.2 ms locate per hit in SURVEY section 6).
kind = "port": it is a restatement, not the reference's files; tests/test_ref_chain_cpu.py checks it against
goldens recorded from the unmodified reference.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import median_filter
from scipy.optimize import fsolve

from . import oracle as orc

find_onset_groups = orc.find_onset_groups  # already plain Python at the reference's granularity


def cc_lag(x, y, o_x, o_y, cutoff, tol):
    """detection.py:239-268 with onsets given (d = 0, inputs already differenced / rectified)."""
    n = len(x)
    cc = np.correlate(x, y, "full")
    norm = np.arange(n) + 1
    norm[:cutoff] = cutoff
    cc[:n] /= norm
    cc[n:] /= norm[n - 2::-1]
    cur = o_y - o_x
    mid = n - cur
    win = cc[mid - tol: mid + tol]
    if len(win) == 0:
        return None
    return int(cur + tol - np.argmax(win))


def adjust(o_x, o_y, x, y, new_lag):
    """detection.py:299-352: which onset of the pair moves by (o_y - o_x) - new_lag."""
    shift = (o_y - o_x) - new_lag
    w = np.exp(np.linspace(0, -np.e, abs(shift)))
    n = len(x)
    if shift < 0:
        xs, xe, ys, ye = max(o_x + shift, 0), min(o_x, n), min(o_y, n), min(o_y - shift, n)
    else:
        xs, xe, ys, ye = o_x, min(o_x + shift, n), max(o_y - shift, 0), min(o_y, n)
    da = np.sum(x[xs:xe] * w[-(xe - xs):]) / x.max()
    db = 0 if ye == ys else np.sum(y[ys:ye] * w[-(ye - ys):][::-1]) / y.max()
    if da > db and o_x + shift >= 0:
        return shift, 0
    return 0, -shift


def fix_onsets(audio, onsets, filter_size=5, d=0, onset_direction=None, take_abs=False, zero_left=False,
               normalization_cutoff=10, onset_tolerance=30, shift_onsets=0):
    """detection.py:373-451 (the sort is pinned to the stable order, H8)."""
    look = normalization_cutoff + onset_tolerance
    out = onsets.copy() + shift_onsets
    for og in out:
        order = np.argsort(og, kind="stable")
        first = order[0]
        start = og[first] - look
        sec = np.diff(median_filter(audio[start: og[order[-1]] + look], filter_size, axes=0), d, axis=0)
        if onset_direction == "up":
            sec[sec < 0] = 0
        elif onset_direction == "down":
            sec[sec > 0] = 0
        if take_abs:
            sec = np.abs(sec)
        pos = og - start
        for i in order[1:]:
            x, y = sec[:, first], sec[:, i]
            if zero_left:
                x[: pos[first]] = 0.0
                y[: pos[i]] = 0.0
            lag = cc_lag(x, y, pos[first], pos[i], normalization_cutoff, onset_tolerance)
            if lag is not None:
                ca, cb = adjust(pos[first], pos[i], x, y, lag)
                og[first] += ca; og[i] += cb
                pos[first] += ca; pos[i] += cb
    return out


def solve_tri3d(sa, sb, so, da, db, guess):
    """multilateration.py:230-316: two range-difference equations, analytic Jacobian, MINPACK hybrj."""
    xa, ya, za = sa; xb, yb, zb = sb; xo, yo, zo = so

    def dist(p, x0, y0, z0):
        return np.sqrt((p[0] - x0) ** 2 + (p[1] - y0) ** 2 + z0 ** 2)

    def f(p):
        d0 = dist(p, xo, yo, zo)
        return np.array([dist(p, xa, ya, za) - d0 - da, dist(p, xb, yb, zb) - d0 - db])

    def jac(p):
        ra, rb, r0 = dist(p, xa, ya, za), dist(p, xb, yb, zb), dist(p, xo, yo, zo)
        return np.array([[(p[0] - xa) / ra - (p[0] - xo) / r0, (p[1] - ya) / ra - (p[1] - yo) / r0],
                         [(p[0] - xb) / rb - (p[0] - xo) / r0, (p[1] - yb) / rb - (p[1] - yo) / r0]])

    root, _, ier, _ = fsolve(f, guess, full_output=True, xtol=0.01, maxfev=20, fprime=jac)
    return tuple(root) if ier == 1 else None


class Locator:
    """Multilaterate3D.locate for complete hits (multilateration.py:428-566 fed the three detections of a hit
    in time order; no ring buffer), on the lag maps of the oracle's numpy port."""

    def __init__(self, sensor_locations, sr=96000, medium="air"):
        self.m = orc.Multilaterate3D(sensor_locations, sr=sr, medium=medium)

    def locate_hit(self, onsets, sensors=None):
        """onsets of three detections; sensors = their sensor indices (default 0, 1, 2)."""
        m = self.m
        ids = list(range(len(onsets))) if sensors is None else [int(v) for v in sensors]
        order = np.argsort(onsets, kind="stable")
        s0, t0 = ids[order[0]], int(onsets[order[0]])
        sens, ons = [s0], [t0]
        for j in order[1:]:
            k = ids[j]
            lag = int(onsets[j]) - t0
            if lag > m.max_max_lags[s0]:
                return None
            if not (m.min_lags[s0][k] < lag < m.max_lags[s0][k]):
                return None
            sens.append(k); ons.append(int(onsets[j]))
        tol = m.samples_per_cm
        l1, l2 = ons[1] - ons[0], ons[2] - ons[0]
        m1, m2 = m.maps[sens[0], sens[1]], m.maps[sens[0], sens[2]]
        legal = (m1 < l1 + tol) & (m1 > l1 - tol) & (m2 < l2 + tol) & (m2 > l2 - tol)
        cell = np.unravel_index(np.argmax(legal > 0), legal.shape, "F")
        if cell == (0, 0):
            return None
        guess = np.array(cell) - m.radius
        if sens[1] == 1:  # multilateration.py:542-544
            sens[1:] = [0, 1]
            ons[1:] = ons[2:0:-1]
        loc = m.sensor_locs
        return solve_tri3d(loc[sens[1]], loc[sens[2]], loc[sens[0]], (ons[1] - ons[0]) / m.sr * m.c,
                           (ons[2] - ons[0]) / m.sr * m.c, guess)


def chain(x, locator: Locator, block_size=128, sr=96000, detect=None):
    """detect -> group -> fix -> locate for one recording x [N, C]; returns (n_onsets, n_hits, n_located)."""
    from . import ref_style

    detect = detect or ref_style.detect_onsets_amplitude
    ch, on, _ = detect(x, block_size=block_size, sr=sr)
    groups = find_onset_groups(on, ch, 1000, x.shape[1])
    if groups is None:
        return len(on), 0, 0
    fixed = fix_onsets(x, groups)
    located = sum(locator.locate_hit(row) is not None for row in fixed)
    return len(on), len(fixed), located
