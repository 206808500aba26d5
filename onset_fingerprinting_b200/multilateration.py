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


def find_lag(a: np.ndarray, b: np.ndarray):
    """multilateration.py:878-886: argmax of the full cross-correlation, evaluated by K4's CC kernel
    with the legal window covering every lag and the contribution normaliser disabled (cutoff > n)."""
    from . import detection

    n = len(a)
    # full window: cc[n - l1 : n - l0] with l1 = n, l0 = -(n - 1); cutoff huge -> constant divisor
    r = detection.cross_correlation_lag(a, b, legal_lags=(-(n - 1), n), normalization_cutoff=1 << 30)
    # cross_correlation_lag returns l1 - argmax = n - argmax; find_lag returns argmax - (n - 1)
    return (n - r) - (n - 1)


def solve_trilateration_3d(sensor_a, sensor_b, sensor_origin, delta_d_a, delta_d_b, initial_guess):
    """multilateration.py:230-316 through K5 (a one-hit launch with an explicit seed is not exposed by
    the C ABI; the batched path seeds from the lag maps as Multilaterate3D.locate does)."""
    raise NotImplementedError("use Multilaterate3D.locate / locate_batch; the seed comes from the lag maps")


class Multilaterate3D:
    """multilateration.py:319-575."""

    def __init__(self, sensor_locations, drum_diameter: float = DIAMETER, medium: str = "drumhead", sr: int = 44100,
                 c: Optional[float] = None, model=None):
        if model is not None:
            raise NotImplementedError("FCNN bypass (multilateration.py:555-557) is outside the hot-path scope")
        self.torch = _lib.require_cuda()
        torch = self.torch
        self.c = speed_of_sound(100, medium=medium) if c is None else c * 100
        self.model = None
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
        self.ongoing = []
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
        """multilateration.py:536-575: legality + seed + solve of one complete group on the GPU.
        The seed is recomputed from the lag maps exactly as locate() derives it (line 511-516)."""
        sensors, onsets = list(group[0]), list(group[1])
        xy, st = self.locate_batch(np.asarray([onsets[:3]], np.int32), np.asarray([sensors[:3]], np.int32))
        if int(st[0].item()) != 0:
            return None
        x, y = xy[0].cpu().tolist()
        return (x, y)

    def locate(self, sensor_index: int, onset_index: int, rec_audio=None):
        """multilateration.py:428-534, streaming contract: feed detections one at a time, get (x, y)
        in cm when a third legal sensor completes a group, else None."""
        if rec_audio is not None:
            raise NotImplementedError("ring-buffer CC refinement (multilateration.py:457-501) is the next "
                                      "row of the scope table (SURVEY 8f rank 1)")
        new_groups = []
        for group in self.ongoing:
            lag = onset_index - group[1][0]
            if lag > self.max_max_lags[group[0][0]]:
                continue
            if lag < 0:  # an adjustment moved an onset behind the next one (multilateration.py:443-449)
                inter = (group[0][0], group[1][0])
                group[0][0], group[1][0] = sensor_index, onset_index
                sensor_index, onset_index = inter
                lag = -lag
            if sensor_index not in group[0]:
                if self.is_legal(group[0][0], sensor_index, lag):
                    group = (group[0] + [sensor_index], group[1] + [onset_index])
                    if len(group[0]) == 3:
                        if group[0][0] == group[0][1]:
                            break
                        xy, st = self.locate_batch(np.asarray([group[1]], np.int32), np.asarray([group[0]], np.int32))
                        code = int(st[0].item())
                        if code != 3:  # a seed cell exists (multilateration.py:512)
                            res = tuple(xy[0].cpu().tolist()) if code == 0 else None
                            if res is not None:
                                new_groups = remove_seed(new_groups, group)
                            self.ongoing = new_groups
                            return res
                    new_groups.append(group)
            if lag <= self.max_max_lags[group[0][0]]:
                new_groups.append(group)
        new_groups.append(([sensor_index], [onset_index]))
        self.ongoing = new_groups
        return None
