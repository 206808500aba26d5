"""Host-side mirror of the reference's ``multilateration.py`` call surface on libofp.so.

Geometry set-up (lag maps, sensor positions; reference multilateration.py:23-157, 902-1001) is
small one-off numpy work and stays on the host exactly as in the reference; the per-hit work --
legality checks, lag-map seed search and the MINPACK hybrj solve that ``fsolve`` performs
(multilateration.py:230-316, 397-426, 536-566) -- runs on the GPU (K5, csrc/multilaterate.cu).

``Multilaterate3D.locate`` keeps the reference's streaming (sensor, onset) contract for the
realtime block API; ``Multilaterate3D.locate_batch`` is the batched form (one launch for H hits).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import check, ptr, stream_ptr

TEMPERATURE = 20.0
HUMIDITY = 0.5
DIAMETER = 14 * 2.54
STRIKE_FORCE = 1.0
C_drumhead = 82
MEDIUM = "air"
ONSET_TOL = 50
NORM_CUTOFF = 10
lookaround = ONSET_TOL + NORM_CUTOFF

STATUS = {0: "located", 1: "lag beyond max_max_lags", 2: "is_legal failed", 3: "no seed cell",
          4: "solver did not converge", 5: "invalid sensors"}


def speed_of_sound(scale: int = 1, temperature: float = TEMPERATURE, humidity: float = HUMIDITY, medium=MEDIUM):
    """multilateration.py:23-39 (m/s times scale)."""
    if medium == "air":
        return scale * (331.3 + 0.606 * temperature) * (1 + 0.0124 * humidity)
    return scale * C_drumhead


def cartesian_to_polar(x, y, r=None):
    """multilateration.py:42-59."""
    rad = np.sqrt(x ** 2 + y ** 2)
    if r is not None:
        rad = rad / r
    phi = np.arctan2(y, x) % (2 * np.pi)
    return rad, np.degrees(phi)


def polar_to_cartesian(r, phi):
    """multilateration.py:62-72."""
    p = np.radians(phi)
    return r * np.cos(p), r * np.sin(p)


def spherical_to_cartesian(r, phi, theta):
    """multilateration.py:75-102."""
    p = np.radians(phi)
    theta = -theta if theta < 0 else 90 - theta
    t = np.radians(theta)
    return r * np.cos(p) * np.sin(t), r * np.sin(p) * np.sin(t), r * np.cos(t)


def cartesian_to_spherical(x, y, z):
    """multilateration.py:105-123."""
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    phi = np.arctan2(y, x) % (2 * np.pi)
    theta = np.degrees(np.arccos(z / r))
    theta = -theta if theta < 0 else 90 - theta
    return r, np.degrees(phi), theta


def cartesian_to_cylindrical(x, y, z, r=None):
    """multilateration.py:126-144."""
    rad, phi = cartesian_to_polar(x, y, r)
    return rad, phi, z


def cylindrical_to_cartesian(r, phi, z):
    """multilateration.py:147-157."""
    x, y = polar_to_cartesian(r, phi)
    return x, y, z


def remove_seed(groups, group):
    """multilateration.py:160-167."""
    return [g for g in groups if not (g[0][0] == group[0][0] and g[1][0] == group[1][0])]


def _lag_map(mic_a, mic_b, r, tol, scale, c, sr):
    g = np.arange(-r, r + 1)
    i, j = np.meshgrid(g, g)
    outside = i ** 2 + j ** 2 > (r + tol * scale) ** 2
    za = mic_a[2] if len(mic_a) > 2 else 0.0
    zb = mic_b[2] if len(mic_b) > 2 else 0.0
    la = np.sqrt((i - mic_a[0]) ** 2 + (j - mic_a[1]) ** 2 + (0 - za) ** 2) / c
    lb = np.sqrt((i - mic_b[0]) ** 2 + (j - mic_b[1]) ** 2 + (0 - zb) ** 2) / c
    lm = np.round((la - lb) * sr).astype(np.float32)
    lm[outside] = np.nan
    return lm


def lag_map_2d(mic_a, mic_b, d=DIAMETER, sr=96000, scale=1, medium=MEDIUM, tol=1, c=None):
    """multilateration.py:902-942."""
    if c is None:
        c = speed_of_sound(100 * scale, medium=medium)
    return _lag_map(tuple(mic_a) + (0.0,), tuple(mic_b) + (0.0,), int(np.round(d * scale / 2)), tol, scale, c, sr)


def lag_map_3d(mic_a, mic_b, d=DIAMETER, sr=96000, scale=1, medium=MEDIUM, tol=1, c=None):
    """multilateration.py:945-1001."""
    if c is None:
        c = speed_of_sound(100 * scale, medium=medium)
    return _lag_map(mic_a, mic_b, int(np.round(d, 1) * scale) // 2, tol, scale, c, sr)


def correlate_full(a, b):
    """np.correlate(a, b, "full") on the device (ofp_correlate_full): a, b [n] or [P, n] float32 ->
    float32 [2n-1] / [P, 2n-1] (numpy in, numpy out; tensors in, tensor out)."""
    from .detection import _to_dev

    torch = _lib.require_cuda()
    is_np = isinstance(a, np.ndarray)
    ad = _to_dev(np.ascontiguousarray(a, dtype=np.float32) if is_np else a, torch)
    bd = _to_dev(np.ascontiguousarray(b, dtype=np.float32) if isinstance(b, np.ndarray) else b, torch)
    one = ad.dim() == 1
    if one:
        ad, bd = ad[None], bd[None]
    assert ad.shape == bd.shape, "np.correlate twin: equal lengths only"
    P, n = ad.shape
    out = torch.empty((P, 2 * n - 1), dtype=torch.float32, device="cuda")
    check(_lib.lib().ofp_correlate_full(ptr(ad), ptr(bd), C.c_int32(P), C.c_int32(n), ptr(out), stream_ptr()))
    out = out[0] if one else out
    return out.cpu().numpy() if is_np else out


def find_lag(a: np.ndarray, b: np.ndarray):
    """multilateration.py:878-886: argmax of the full cross-correlation (first maximum) minus len(a) - 1."""
    cc = correlate_full(np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32))
    return int(np.argmax(cc) - (len(a) - 1))


def find_lag_multi(a, b, top_n=3):
    """multilateration.py:889-899: the top_n peaks of the full cross-correlation (lags, squared heights).
    The correlation runs on the device; peak picking over the 2n-1 values is scipy's find_peaks as in
    the reference."""
    from scipy.signal import find_peaks

    cc = correlate_full(np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32))
    peaks, _ = find_peaks(cc)
    peaks = peaks[np.argsort(-cc[peaks])][:top_n]
    return peaks - len(a) + 1, cc[peaks] ** 2


def solve_trilateration_batch(sensor_a, sensor_b, sensor_origin, delta_d_a, delta_d_b, initial_guess,
                              xtol: float = 0.01, maxfev: int = 20):
    """P trilateration problems in one launch (ofp_solve_trilateration = MINPACK hybrj as fsolve runs it).
    sensor_* [P, 2|3] (or one sensor for all), delta_d_* [P], initial_guess [P, 2] ->
    (xy [P, 2] float64, ier [P] int32) device tensors; ier == 1 where the reference returns a root."""
    torch = _lib.require_cuda()
    g = np.atleast_2d(np.asarray(initial_guess, np.float64))
    P = g.shape[0]
    prob = np.zeros((P, 11), np.float64)
    for k, sensor in enumerate((sensor_a, sensor_b, sensor_origin)):
        v = np.atleast_2d(np.asarray(sensor, np.float64))
        prob[:, 3 * k:3 * k + v.shape[1]] = v
    prob[:, 9] = np.asarray(delta_d_a, np.float64)
    prob[:, 10] = np.asarray(delta_d_b, np.float64)
    pd = torch.from_numpy(prob).cuda()
    gd = torch.from_numpy(np.ascontiguousarray(g)).cuda()
    xy = torch.empty((P, 2), dtype=torch.float64, device="cuda")
    ier = torch.empty((P,), dtype=torch.int32, device="cuda")
    check(_lib.lib().ofp_solve_trilateration(ptr(pd), ptr(gd), C.c_int32(P), C.c_double(xtol), C.c_int32(maxfev),
                                             ptr(xy), ptr(ier), None, stream_ptr()))
    return xy, ier


def solve_trilateration(sensor_a, sensor_b, sensor_origin, delta_d_a, delta_d_b, initial_guess):
    """multilateration.py:170-227 (2-D): tuple (x, y) or None when fsolve does not report convergence."""
    xy, ier = solve_trilateration_batch(sensor_a, sensor_b, sensor_origin, delta_d_a, delta_d_b, initial_guess)
    if int(ier[0].item()) != 1:
        return None
    return tuple(xy[0].cpu().tolist())


def solve_trilateration_3d(sensor_a, sensor_b, sensor_origin, delta_d_a, delta_d_b, initial_guess):
    """multilateration.py:230-316: sensors in 3-D, source on the z = 0 plane."""
    return solve_trilateration(sensor_a, sensor_b, sensor_origin, delta_d_a, delta_d_b, initial_guess)


def sound_intensity_at_source(strike_location, strike_force=STRIKE_FORCE, diameter=DIAMETER) -> float:
    """multilateration.py:1004-1008 (placeholder in the reference too)."""
    return strike_force


def vec_sub(a, b):
    """multilateration.py:1011-1015."""
    x = a[0] - b[0].reshape(-1)
    y = a[1] - b[1].reshape(-1)
    z = np.full_like(x, a[2] - b[2], dtype=float)
    return np.vstack((x, y, z)).T


def attenuate_intensity(source_loc, mic_loc, reflectivity, intensity_at_source):
    """multilateration.py:1018-1041: 1/r law with an angle term towards the drumhead normal."""
    direction = vec_sub(mic_loc, source_loc)
    distance = np.linalg.norm(direction, axis=-1)
    direction /= np.linalg.norm(direction, axis=-1, keepdims=True)
    thetas = np.arccos(np.dot(direction, np.array([0.0, 0.0, 1.0])))
    A = intensity_at_source * (1 + reflectivity * (1 - np.abs(np.cos(thetas)))) / distance
    return A, np.degrees(thetas)


def lag_intensity_map(mic_a, mic_b, reflectivity: float = 0.5, d: int = DIAMETER, sr: int = 96000,
                      scale: float = 1, medium: str = MEDIUM):
    """multilateration.py:1044-1101: lag map plus the two intensity maps (dB) of a microphone pair;
    one-off geometry set-up on the host like lag_map_2d/3d."""
    r = int(np.round(d, 1) * scale) // 2
    i, j = np.meshgrid(range(-r, r + 1), range(-r, r + 1))
    c = speed_of_sound(100 * scale, medium=medium)

    def at_mic(mic):
        A, _ = attenuate_intensity((i, j, 0), np.array(mic), reflectivity, 1)
        return A.reshape(i.shape)

    la = np.sqrt((i - mic_a[0]) ** 2 + (j - mic_a[1]) ** 2 + (0 - mic_a[2]) ** 2) / c
    lb = np.sqrt((i - mic_b[0]) ** 2 + (j - mic_b[1]) ** 2 + (0 - mic_b[2]) ** 2) / c
    return (np.round((la - lb) * sr).astype(np.float32), (10 * np.log10(at_mic(mic_a))).astype(np.float32),
            (10 * np.log10(at_mic(mic_b))).astype(np.float32))


class Multilaterate3D:
    """multilateration.py:319-575."""

    def __init__(self, sensor_locations, drum_diameter: float = DIAMETER, medium: str = "drumhead", sr: int = 44100,
                 c: Optional[float] = None, model=None):
        self.torch = _lib.require_cuda()
        torch = self.torch
        self.c = speed_of_sound(100, medium=medium) if c is None else c * 100
        # multilateration.py:350, 553-557: an optional network that maps the two lags straight to (x, y) in m
        self.model = model
        self.radius = drum_diameter / 2
        self.sensor_locs = [spherical_to_cartesian(x[0] * self.radius, x[1], x[2]) for x in sensor_locations]
        self.medium, self.sr = medium, sr
        self.samples_per_cm = sr / self.c
        S = len(self.sensor_locs)
        self.lag_maps = [{} for _ in range(S)]
        self.max_lags = [{} for _ in range(S)]
        self.min_lags = [{} for _ in range(S)]
        for i in range(S):
            for j in range(S):
                if i == j:
                    continue
                lm = lag_map_3d(self.sensor_locs[j], self.sensor_locs[i], d=drum_diameter, sr=sr, scale=1,
                                medium=medium, tol=2, c=self.c)
                lm[lm < -self.samples_per_cm * 1] = np.nan  # multilateration.py:387
                self.lag_maps[i][j] = lm
                self.max_lags[i][j] = np.nanmax(lm)
                self.min_lags[i][j] = np.nanmin(lm)
        self.max_max_lags = [np.nanmax(list(d.values())) for d in self.max_lags]
        # device copies for K5
        M = self.lag_maps[0][1].shape[0]
        maps = np.full((S, S, M, M), np.nan, np.float32)
        mx = np.full((S, S), np.nan, np.float32)
        mn = np.full((S, S), np.nan, np.float32)
        for i in range(S):
            for j in range(S):
                if i != j:
                    maps[i, j], mx[i, j], mn[i, j] = self.lag_maps[i][j], self.max_lags[i][j], self.min_lags[i][j]
        self._S, self._M = S, M
        self._maps = torch.from_numpy(maps).cuda()
        self._mx = torch.from_numpy(mx).cuda()
        self._mn = torch.from_numpy(mn).cuda()
        self._mm = torch.from_numpy(np.asarray(self.max_max_lags, np.float32)).cuda()
        self._locs = torch.from_numpy(np.asarray(self.sensor_locs, np.float64)).cuda()

    # -- batched form ---------------------------------------------------
    def locate_batch(self, onsets, sensors=None):
        """onsets [H, >=3] int32 (device or numpy): the first three columns are the onsets of the
        three sensors given by `sensors` [H, 3] (default 0, 1, 2).  Returns (xy [H, 2] float64 with
        NaN where the reference returns None, status [H] int32) as device tensors."""
        torch = self.torch
        if isinstance(onsets, np.ndarray):
            onsets = torch.from_numpy(np.ascontiguousarray(onsets))
        onsets = onsets.to(device="cuda", dtype=torch.int32).contiguous()
        if sensors is not None:
            if isinstance(sensors, np.ndarray):
                sensors = torch.from_numpy(np.ascontiguousarray(sensors))
            sensors = sensors.to(device="cuda", dtype=torch.int32).contiguous()
        H = onsets.shape[0]
        xy = torch.empty((H, 2), dtype=torch.float64, device="cuda")
        st = torch.empty((H,), dtype=torch.int32, device="cuda")
        if self.model is not None:
            # same legality / seed checks, then res = model((d_a1, d_b1)) * 100 (multilateration.py:553-557)
            lags = torch.zeros((H, 2), dtype=torch.float32, device="cuda")
            check(_lib.lib().ofp_locate_hits_lags(ptr(self._locs), C.c_int32(self._S), ptr(self._maps),
                                                  C.c_int32(self._M), ptr(self._mx), ptr(self._mn), ptr(self._mm),
                                                  C.c_double(self.radius), C.c_double(self.samples_per_cm),
                                                  C.c_double(self.sr), C.c_double(self.c), ptr(sensors), ptr(onsets),
                                                  C.c_int32(onsets.stride(0)), C.c_int32(H), ptr(lags), ptr(xy),
                                                  ptr(st), stream_ptr()))
            self.model.forward_device(lags, status=st, out_scale=100.0, out_f64=xy)
            return xy, st
        check(_lib.lib().ofp_locate_hits(ptr(self._locs), C.c_int32(self._S), ptr(self._maps), C.c_int32(self._M),
                                         ptr(self._mx), ptr(self._mn), ptr(self._mm), C.c_double(self.radius),
                                         C.c_double(self.samples_per_cm), C.c_double(self.sr), C.c_double(self.c),
                                         ptr(sensors), ptr(onsets), C.c_int32(onsets.stride(0)), C.c_int32(H),
                                         ptr(xy), ptr(st), stream_ptr()))
        return xy, st

    # -- reference surface ---------------------------------------------------------------------
    def is_legal(self, first_sensor: int, later_sensor: int, lag: int) -> bool:
        """multilateration.py:397-411."""
        return bool(self.min_lags[first_sensor][later_sensor] < lag < self.max_lags[first_sensor][later_sensor])

    def is_legal_3d(self, group, tolerance=1):
        """multilateration.py:413-426 (host lookup; the batched path does this inside K5)."""
        tolerance *= self.samples_per_cm
        sensors, onsets = group[0], group[1]
        lag1, lag2 = onsets[1] - onsets[0], onsets[2] - onsets[0]
        lm1, lm2 = self.lag_maps[sensors[0]][sensors[1]], self.lag_maps[sensors[0]][sensors[2]]
        legal = (lm1 < lag1 + tolerance) & (lm1 > lag1 - tolerance) & (lm2 < lag2 + tolerance) & (lm2 > lag2 - tolerance)
        return np.unravel_index(np.argmax(legal > 0), legal.shape, "F")

    def trilaterate(self, group, initial_guess=None):
        """multilateration.py:536-575.  With a seed: the reference's sensor rewrite (Q8, in place on the
        caller's lists like the reference) and one solve.  Without: legality + seed + solve of a
        complete group in one K5 launch, the seed derived from the lag maps as locate() does (511-516)."""
        sensors, onsets = group[0], group[1]
        if initial_guess is None:
            xy, st = self.locate_batch(np.asarray([list(onsets)[:3]], np.int32), np.asarray([list(sensors)[:3]], np.int32))
            if int(st[0].item()) != 0:
                return None
            return tuple(xy[0].cpu().tolist())
        if sensors[1] == 1:
            sensors[1:] = [0, 1]
            onsets[1:] = onsets[2:0:-1]
        d_a1, d_b1 = onsets[1] - onsets[0], onsets[2] - onsets[0]
        if self.model is not None:
            return self.model.call_np((d_a1, d_b1)) * 100
        return solve_trilateration_3d(self.sensor_locs[sensors[1]], self.sensor_locs[sensors[2]],
                                      self.sensor_locs[sensors[0]], d_a1 / self.sr * self.c, d_b1 / self.sr * self.c,
                                      initial_guess)

    # -- streaming form: the group state machine lives on the device ------------------------------------
    _SL_GMAX, _SL_LEN = 16, 4  # csrc/multilaterate.cu: groups kept per stream, members kept per group

    def _stream(self):
        """Device state of the streaming contract for ONE stream (the realtime sessions hold S of them):
        the `ongoing` group lists (csrc/multilaterate.cu:k5_stream_locate), one detection slot, the results."""
        st = getattr(self, "_sl", None)
        if st is None:
            torch = self.torch
            sizes = [C.c_int64() for _ in range(4)]
            check(_lib.lib().ofp_stream_locate_state_bytes(C.c_int32(1), *[C.byref(v) for v in sizes]))
            st = self._sl = {
                "state": [torch.zeros((v.value,), dtype=torch.uint8, device="cuda") for v in sizes],
                "det": torch.zeros((3, 32), dtype=torch.int32, device="cuda"),  # channel / delta / count rows
                "det_h": torch.zeros((3, 32), dtype=torch.int32).pin_memory(),
                "cur": torch.zeros((1,), dtype=torch.int64, device="cuda"),
                "xy": torch.empty((1, 2), dtype=torch.float64, device="cuda"),
                "found": torch.empty((1,), dtype=torch.int32, device="cuda"),
                "ring": None, "ring_count": 0, "ring_src": None,
            }
        return st

    @property
    def ongoing(self):
        """The reference's list of (sensors, onsets) groups, decoded from the device state."""
        st = getattr(self, "_sl", None)
        if st is None:
            return []
        cnt, ln, sen, ons = (t.cpu().numpy() for t in st["state"])
        ng = int(cnt.view(np.int32)[0]) & 0xff
        ln, sen, ons = ln.view(np.int32), sen.view(np.int32), ons.view(np.int64)
        out = []
        for g in range(ng):
            k = int(ln[g]) & 0xff
            out.append(([int(v) for v in sen[g * self._SL_LEN:g * self._SL_LEN + k]],
                        [int(v) for v in ons[g * self._SL_LEN:g * self._SL_LEN + k]]))
        return out

    @ongoing.setter
    def ongoing(self, groups):
        """``m.ongoing = []`` (the reference's reset) or an explicit list of (sensors, onsets) groups; equal
        neighbouring groups are taken to be the same tuple, as the reference's own lists hold them."""
        groups = list(groups)
        if not groups and getattr(self, "_sl", None) is None:
            return
        if len(groups) > self._SL_GMAX or any(len(g[0]) > self._SL_LEN for g in groups):
            raise ValueError(f"at most {self._SL_GMAX} groups of {self._SL_LEN} members")
        st = self._stream()
        ln = np.zeros(self._SL_GMAX, np.int32)
        sen = np.full(self._SL_GMAX * self._SL_LEN, -1, np.int32)
        ons = np.zeros(self._SL_GMAX * self._SL_LEN, np.int64)
        uid = 0
        for g, (ss, oo) in enumerate(groups):
            if g and (list(ss), list(oo)) != (list(groups[g - 1][0]), list(groups[g - 1][1])):
                uid += 1
            ln[g] = len(ss) | (uid << 8)
            sen[g * self._SL_LEN:g * self._SL_LEN + len(ss)] = ss
            ons[g * self._SL_LEN:g * self._SL_LEN + len(oo)] = oo
        cnt = np.asarray([len(groups) | ((uid + 1) << 8)], np.int32)
        for t, v in zip(st["state"], (cnt, ln, sen, ons)):
            t.copy_(self.torch.from_numpy(v.view(np.uint8)))

    def _mirror_ring(self, st, rec_audio):
        """Keep a device copy of rec_audio in the layout the kernel reads (row = sample index % rows): only the
        rows written since the last call are copied."""
        torch = self.torch
        NR = int(rec_audio.N)
        probe = rec_audio[-1:]
        Cn = int(probe.shape[1])
        if st["ring"] is None or st["ring_src"] is not rec_audio or tuple(st["ring"].shape) != (1, NR, Cn) \
                or rec_audio.counter < st["ring_count"]:
            st["ring"] = torch.zeros((1, NR, Cn), dtype=torch.float32, device="cuda")
            st["ring_count"], st["ring_src"] = 0, rec_audio
        k = min(int(rec_audio.counter) - st["ring_count"], NR)
        if k > 0:
            new = rec_audio[-k:]
            new = torch.from_numpy(np.ascontiguousarray(new, np.float32)).cuda() if isinstance(new, np.ndarray) \
                else new.to(device="cuda", dtype=torch.float32)
            rows = (torch.arange(int(rec_audio.counter) - k, int(rec_audio.counter), device="cuda") % NR)
            st["ring"][0, rows] = new
        st["ring_count"] = int(rec_audio.counter)
        return NR, Cn

    def locate(self, sensor_index: int, onset_index: int, rec_audio=None):
        """multilateration.py:428-534, streaming contract: feed detections one at a time, get (x, y)
        in cm when a third legal sensor completes a group, else None.  rec_audio: a ring of the most
        recent audio rows (``counter`` = rows written so far, ``N`` rows, ``ring[-k:]`` = last k rows in time
        order; realtime.audio.DeviceRing or loopmate's CircularArray) enables the cross-correlation
        refinement of each new pair (multilateration.py:457-501).

        One launch of the device state machine that the realtime sessions run for thousands of streams
        (ofp_stream_locate / ofp_stream_locate_ring_dev with one stream and one detection); the group lists
        stay on the device (`ongoing` decodes them)."""
        if self.model is not None:
            raise NotImplementedError("streaming locate with a model: use locate_batch / trilaterate (the model "
                                      "bypass of multilateration.py:553-557 is built for complete groups)")
        st = self._stream()
        L = _lib.lib()
        h = st["det_h"]
        h[0, 0], h[1, 0], h[2, 0] = int(sensor_index), 0, 1
        st["det"].copy_(h, non_blocking=True)
        geo = (ptr(self._locs), C.c_int32(self._S), ptr(self._maps), C.c_int32(self._M), ptr(self._mx), ptr(self._mn),
               ptr(self._mm), C.c_double(self.radius), C.c_double(self.samples_per_cm), C.c_double(self.sr),
               C.c_double(self.c))
        state = [ptr(t) for t in st["state"]]
        if rec_audio is None:
            check(L.ofp_stream_locate(*geo, C.c_int32(1), C.c_int32(32), ptr(st["det"][0]), ptr(st["det"][1]),
                                      ptr(st["det"][2]), C.c_int64(int(onset_index)), *state, ptr(st["xy"]),
                                      ptr(st["found"]), stream_ptr()))
        else:
            NR, Cn = self._mirror_ring(st, rec_audio)
            written = int(rec_audio.counter) - int(onset_index)  # rows of the ring after the detection's own
            if not 1 <= written <= NR:
                raise ValueError("the detection lies outside the ring buffer")
            st["cur"].fill_(int(onset_index))
            check(L.ofp_stream_locate_ring_dev(*geo, C.c_int32(1), C.c_int32(Cn), ptr(st["det"][0]), ptr(st["det"][1]),
                                               ptr(st["det"][2]), ptr(st["cur"]), C.c_int32(0), ptr(st["ring"]),
                                               C.c_int32(NR), C.c_int32(written), C.c_int32(ONSET_TOL),
                                               C.c_int32(NORM_CUTOFF), *state, ptr(st["xy"]), ptr(st["found"]),
                                               stream_ptr()))
        found = int(st["found"].item())
        if found == -2:
            raise ValueError("operands could not be broadcast together (adjust_onset, detection.py:335)")
        if found == -1:
            raise RuntimeError("streaming locate: more than 16 ongoing groups / 4 members, or a refinement section "
                               "longer than 1024 samples")
        if found != 1:
            return None
        return tuple(st["xy"][0].cpu().tolist())


class Multilaterate:
    """multilateration.py:578-733: the 2-D (sensors on the drumhead plane) streaming locator.  Set-up
    as in the reference (cm lag maps on the host); the solve is ofp_solve_trilateration."""

    def __init__(self, sensor_locations, drum_diameter: float = DIAMETER, medium: str = "drumhead", sr: int = 44100):
        _lib.require_cuda()
        self.radius = drum_diameter / 2
        self.sensor_locs = [polar_to_cartesian(x[0] * self.radius, x[1]) for x in sensor_locations]
        self.medium, self.sr = medium, sr
        self.samples_per_cm = sr / speed_of_sound(100, medium=medium)
        S = len(self.sensor_locs)
        self.lag_maps = [{} for _ in range(S)]
        self.max_lags = [{} for _ in range(S)]
        self.min_lags = [{} for _ in range(S)]
        for i in range(S):
            for j in range(S):
                if i == j:
                    continue
                lm = lag_map_2d(self.sensor_locs[j], self.sensor_locs[i], d=drum_diameter, sr=sr, scale=1,
                                medium=medium, tol=2)
                lm[lm < -self.samples_per_cm * 1] = np.nan
                self.lag_maps[i][j] = lm
                self.max_lags[i][j] = np.nanmax(lm)
                self.min_lags[i][j] = np.nanmin(lm)
        self.max_max_lags = [np.nanmax(list(d.values())) for d in self.max_lags]
        self.ongoing = []

    is_legal = Multilaterate3D.is_legal
    is_legal_3d = Multilaterate3D.is_legal_3d

    def locate(self, sensor_index: int, onset_index: int):
        """multilateration.py:679-713: the 2-D streaming contract.  Every ongoing group that the detection legally
        extends is tried; the first group it completes to three sensors with a legal lag-map cell is solved and
        ends the call (the groups visited so far stay, the rest are dropped -- as the reference does).  Quirk kept:
        an extended group is listed twice (the reference's loop variable is rebound to the extended tuple before the
        keep-alive test), the un-extended one is not kept."""
        kept = []

        def solved(group):
            """(done, result): a complete group with a legal seed cell goes to the solver."""
            if len(group[0]) != 3:
                return False, None
            cell = self.is_legal_3d(group)
            if cell == (0, 0):
                return False, None
            return True, self.trilaterate(group, np.array(cell) - self.radius)

        for sensors, onsets in self.ongoing:
            lag = onset_index - onsets[0]
            current = (sensors, onsets)
            if sensor_index not in sensors and self.is_legal(sensors[0], sensor_index, lag):
                current = (sensors + [sensor_index], onsets + [onset_index])
                done, res = solved(current)
                if done:
                    self.ongoing = kept
                    return res
                kept.append(current)
            if lag <= self.max_max_lags[current[0][0]]:
                kept.append(current)
        kept.append(([sensor_index], [onset_index]))
        self.ongoing = kept
        return None

    def trilaterate(self, group, initial_guess):
        """multilateration.py:715-733 -> (relative radius, angle in degrees) or None."""
        sensors, onsets = group[0], group[1]
        c = speed_of_sound(100, medium=self.medium)
        res = solve_trilateration(self.sensor_locs[sensors[1]], self.sensor_locs[sensors[2]],
                                  self.sensor_locs[sensors[0]], (onsets[1] - onsets[0]) * c / self.sr,
                                  (onsets[2] - onsets[0]) * c / self.sr, initial_guess)
        return None if res is None else cartesian_to_polar(*res, self.radius)


class MultilateratePaired:
    """multilateration.py:736-875: neighbour-pair lag maps at `scale` resolution; locate() solves from
    two lags, locate_cc() votes the lag maps with cross-correlation lags found on the device."""

    def __init__(self, sensor_locations, drum_diameter: float = DIAMETER, scale: float = 10,
                 medium: str = "drumhead", sr: int = 44100):
        _lib.require_cuda()
        self.radius = int(np.round(drum_diameter * scale / 2, 1))
        self.sensor_locs = [polar_to_cartesian(x[0] * self.radius, x[1]) for x in sensor_locations]
        self.scale, self.medium, self.sr = scale, medium, sr
        S = len(self.sensor_locs)
        self.lag_maps = [{} for _ in range(S)]
        for i in range(S):
            for k in (-1, 1):
                j = (i + k) % S
                self.lag_maps[i][j] = lag_map_2d(self.sensor_locs[i], self.sensor_locs[j], d=drum_diameter, sr=sr,
                                                 scale=scale, medium="drumhead")
        self.res = np.zeros_like(self.lag_maps[0][1])

    def locate(self, lags, i: int):
        """multilateration.py:801-832."""
        S = len(self.sensor_locs)
        sensor_a, sensor_b, sensor_origin = self.sensor_locs[(i - 1) % S], self.sensor_locs[(i + 1) % S], self.sensor_locs[i]
        c = speed_of_sound(100 * self.scale, medium=self.medium)
        d_a1, d_b1 = lags[0] * c / self.sr, lags[1] * c / self.sr
        wa, wb, wo = abs(d_a1) / self.radius, abs(d_b1) / self.radius, abs(d_a1 + d_b1) / (2 * self.radius)
        guess = np.array([sensor_a[0] * wa + sensor_b[0] * wb + sensor_origin[0] * wo,
                          sensor_a[1] * wa + sensor_b[1] * wb + sensor_origin[1] * wo])
        x, y = solve_trilateration(sensor_a, sensor_b, sensor_origin, d_a1, d_b1, guess)
        return cartesian_to_polar(x, y, self.radius)

    def locate_cc(self, x: np.ndarray, onset_idx: int, i: int, tol: int = 2, left: int = 0, right: int = 256):
        """multilateration.py:834-875: both neighbour lags in one correlation launch, then the map vote."""
        js = list(self.lag_maps[i])
        seg = np.ascontiguousarray(x[onset_idx - left:onset_idx + right], np.float32)
        a = np.stack([seg[:, i]] * len(js))
        b = np.stack([seg[:, j] for j in js])
        cc = correlate_full(a, b)
        self.res[:] = 0
        for row, j in zip(cc, js):
            lag = int(np.argmax(row)) - (seg.shape[0] - 1)
            self.res += (self.lag_maps[i][j] < lag + tol) & (self.lag_maps[i][j] > lag - tol)
        coord = np.unravel_index(np.argmax(self.res), self.res.shape)
        px = coord[1] - (self.res.shape[1] - 1) / 2
        py = (self.res.shape[0] - 1) / 2 - coord[0]
        return cartesian_to_polar(px, py, self.radius)
