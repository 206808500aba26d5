"""Seeded synthetic multi-mic drum audio (host/numpy generator).

The reference ships no data; every parity vector and benchmark
input is synthetic.  Geometry and hit process follow SURVEY.md section 8(d):
14" drum, 96 kHz, one decaying 900 Hz burst every 0.118 s at a uniformly random
position, per-channel arrival delay round(dist / c * sr), 1/dist amplitude, white
noise sigma 1e-4, float32, layout [N, C] (time-major interleaved, as soundfile
would load it) or [R, N, C] for a batch of recordings.

This numpy generator feeds the tests and golden fixtures; the bench generates the
same signal model on device (csrc/synth.cu) and copies a sample back for the CPU
baseline.
"""
from __future__ import annotations

import math

import numpy as np

DIAMETER = 14 * 2.54  # cm, multilateration.py:12 of the reference
SR = 96000

# close mic is index 2 (Multilaterate3D.trilaterate assumes it)
SENSORS_3MIC = [(0.9, 140.0, 75.0), (0.9, 10.0, 55.0), (0.5, 100.0, 15.0)]
SENSORS_16MESH = [(0.9, 22.5 * k, 0.0) for k in range(16)]


def sensor_xyz(sensor_locations, diameter: float = DIAMETER) -> np.ndarray:
    """Spherical (r_rel, phi_deg, theta_deg) -> cartesian cm, same convention as the
    reference's spherical_to_cartesian (multilateration.py:75-102)."""
    out = []
    radius = diameter / 2
    for r, phi, theta in sensor_locations:
        r = r * radius
        th = -theta if theta < 0 else 90 - theta
        p, t = math.radians(phi), math.radians(th)
        out.append((r * math.cos(p) * math.sin(t), r * math.sin(p) * math.sin(t), r * math.cos(t)))
    return np.asarray(out, dtype=np.float64)


def speed_cm_s(medium: str) -> float:
    """cm/s; multilateration.py:23-39 with scale=100."""
    if medium == "air":
        return 100 * (331.3 + 0.606 * 20.0) * (1 + 0.0124 * 0.5)
    return 100 * 82


def drum_recording(
    seconds: float = 2.0,
    sensors=SENSORS_3MIC,
    medium: str = "air",
    seed: int = 0,
    sr: int = SR,
    hit_period: float = 0.118,
    noise: float = 1e-4,
    first_hit: int | None = None,
    burst_len: int = 4096,
    freq: float = 900.0,
    decay: float = 400.0,
):
    """Returns (x[N, C] float32, hits) where hits is a dict with the ground truth:
    'pos' [H, 2] cm, 'arrival' [H, C] int sample index of the burst start per channel."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    locs = sensor_xyz(sensors)
    n_ch = len(locs)
    c = speed_cm_s(medium)
    x = (noise * rng.standard_normal((n, n_ch))).astype(np.float32)
    t = np.arange(burst_len) / sr
    burst = np.exp(-decay * t) * np.sin(2 * np.pi * freq * t)
    start = int(0.5 * sr) + 1000 if first_hit is None else first_hit
    step = int(round(hit_period * sr))
    pos, arrival = [], []
    radius = DIAMETER / 2
    for s in range(start, n - burst_len - 2048, step):
        rr = 0.85 * radius * math.sqrt(rng.uniform())
        ang = rng.uniform(0, 2 * math.pi)
        p = np.array([rr * math.cos(ang), rr * math.sin(ang), 0.0])
        dist = np.sqrt(((locs - p) ** 2).sum(1))
        delay = np.round(dist / c * sr).astype(int)
        for ch in range(n_ch):
            a = s + delay[ch]
            x[a : a + burst_len, ch] += (0.5 * 10.0 / dist[ch] * burst).astype(np.float32)
        pos.append(p[:2])
        arrival.append(s + delay)
    hits = {
        "pos": np.asarray(pos, dtype=np.float64).reshape(-1, 2),
        "arrival": np.asarray(arrival, dtype=np.int64).reshape(-1, n_ch),
    }
    return x, hits


def drum_batch(n_rec: int, seconds: float = 2.0, seed: int = 0, **kw):
    """[R, N, C] batch; recording r uses seed + r."""
    xs, hs = [], []
    for r in range(n_rec):
        x, h = drum_recording(seconds, seed=seed + r, **kw)
        xs.append(x)
        hs.append(h)
    return np.stack(xs), hs


def drum_batch_device(n_rec: int, n_samples: int, sensors=SENSORS_3MIC, medium: str = "air", seed: int = 0,
                      sr: int = SR, hit_period: float = 0.118, noise: float = 1e-4, rec_offset: int = 0, out=None,
                      first_hit: int | None = None, tail_guard: int = 4096 + 2048):
    """Same signal model generated on the GPU (csrc/synth.cu): returns a CUDA tensor [R, N, C]."""
    import ctypes as C

    from . import _lib

    torch = _lib.require_cuda()
    locs = sensor_xyz(sensors).astype(np.float32)
    n_ch = len(locs)
    x = out if out is not None else torch.empty((n_rec, n_samples, n_ch), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().ofp_synth_drum(
        _lib.ptr(x), C.c_int64(n_rec), C.c_int64(n_samples), C.c_int32(n_ch),
        locs.ctypes.data_as(C.c_void_p), C.c_float(speed_cm_s(medium)), C.c_float(sr), C.c_float(noise),
        C.c_float(DIAMETER / 2), C.c_int64(int(0.5 * sr) + 1000 if first_hit is None else first_hit),
        C.c_int64(int(round(hit_period * sr))), C.c_int32(tail_guard), C.c_uint64(seed), C.c_int64(rec_offset), _lib.stream_ptr()))
    return x
