"""POSD session files: the on-disk format on either side of the hot path.

The reference stores a recording session as ``<name>.wav`` (all channels interleaved, read with
``soundfile.read`` in data.py:297) next to ``<name>.json`` = ``{"meta": {...}, "hits": [{"i",
"onset_start": [one index per channel], "zone", "location", ...}]}`` (notebooks/dataset_spec_draft.org:
240-268, 333-397; ``MCPOSD.from_file``, data.py:285-311; ``read_json`` / ``parse_hits``, data.py:31-52).
This module writes and reads exactly that pair without soundfile (absent here): a RIFF/WAVE file with
IEEE float32 samples (format tag 3) or 16/24/32-bit PCM on the read side, so the results of the GPU path
(onset groups after lag refinement + located positions) can be saved as a POSD session and a saved session
can be loaded back into ``data.MCPOSD`` for window extraction on the device.  Host-side I/O only; nothing
here touches the GPU.
"""
from __future__ import annotations

import json
import struct
from pathlib import Path

import numpy as np


def write_wav(path, audio: np.ndarray, sr: int) -> None:
    """audio [N] or [N, C] float32 -> IEEE-float WAVE file (what ``soundfile.write(..., subtype='FLOAT')`` emits)."""
    a = np.ascontiguousarray(audio, dtype="<f4")
    if a.ndim == 1:
        a = a[:, None]
    n, c = a.shape
    data = a.tobytes()
    fmt = struct.pack("<HHIIHH", 3, c, sr, sr * c * 4, c * 4, 32)
    fact = struct.pack("<I", n)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"fact" + struct.pack("<I", 4) + fact
    chunks += b"data" + struct.pack("<I", len(data)) + data + (b"\x00" if len(data) % 2 else b"")
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks)


def read_wav(path) -> tuple[np.ndarray, int]:
    """-> (audio [N, C] float32 (or [N] for mono, like soundfile.read), sample rate)."""
    raw = Path(path).read_bytes()
    if raw[:4] != b"RIFF" or raw[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(raw):
        cid, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        body = raw[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:  # WAVE_FORMAT_EXTENSIBLE: the sub-format's first word
                fmt = (struct.unpack("<H", body[24:26])[0],) + fmt[1:]
        elif cid == b"data":
            data = body
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, c, sr, _, _, bits = fmt
    if tag == 3 and bits == 32:
        a = np.frombuffer(data, dtype="<f4").astype(np.float32)
    elif tag == 3 and bits == 64:
        a = np.frombuffer(data, dtype="<f8").astype(np.float32)
    elif tag == 1 and bits == 16:
        a = np.frombuffer(data, dtype="<i2").astype(np.float32) / 32768.0
    elif tag == 1 and bits == 32:
        a = np.frombuffer(data, dtype="<i4").astype(np.float32) / 2147483648.0
    elif tag == 1 and bits == 24:
        b = np.frombuffer(data, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        a = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag} with {bits} bits")
    a = a[: len(a) // c * c].reshape(-1, c)
    return (a[:, 0] if c == 1 else a), sr


def session_dict(onsets, locations=None, sensors=None, meta: dict | None = None, zones=None, extra=None) -> dict:
    """Build the POSD session dictionary.
    onsets [H, C] ints (one onset per channel and hit, -1 = missing, dataset_spec_draft.org:247-250);
    locations [H, 2] or None (NaN rows = not located: the key is omitted for that hit);
    sensors: list of (r, phi[, z]) per channel -> meta["channels"] with polar locations (spec example);
    zones / extra: optional per-hit zone names / dicts of further per-hit fields (velocity, conditions ...)."""
    onsets = np.asarray(onsets)
    m = dict(meta or {})
    if sensors is not None and "channels" not in m:
        m["channels"] = {f"ch{i}": {"location": [float(v) for v in s], "coordinate_system": "polar" if len(s) == 2 else "spherical"}
                         for i, s in enumerate(sensors)}
    hits = []
    for i, row in enumerate(onsets):
        h = {"i": i, "onset_start": [int(v) for v in np.atleast_1d(row)]}
        if zones is not None:
            h["zone"] = zones[i]
        if locations is not None and not np.isnan(np.asarray(locations[i], dtype=float)).any():
            h["location"] = [float(v) for v in locations[i]]
        if extra is not None:
            h.update(extra[i])
        hits.append(h)
    return {"meta": m, "hits": hits}


def write_session(folder, name: str, audio: np.ndarray, sr: int, onsets, locations=None, sensors=None,
                  meta: dict | None = None, zones=None, extra=None) -> dict:
    """<folder>/<name>.wav + <folder>/<name>.json, the pair MCPOSD.from_file reads (data.py:296-301)."""
    folder = Path(folder)
    folder.mkdir(parents=True, exist_ok=True)
    write_wav(folder / f"{name}.wav", audio, sr)
    d = session_dict(onsets, locations, sensors, meta, zones, extra)
    with open(folder / f"{name}.json", "w") as f:
        json.dump(d, f, indent=1)
    return d


def onsets_to_hits(onsets, zone: str = "center") -> dict:
    """notebooks/refresh.org:243-249: the `combined.json` table -- {"hits": [{"i", "zone", "onset_start"}]}."""
    return {"hits": [{"i": i, "zone": zone, "onset_start": [int(v) for v in np.atleast_1d(row)]}
                     for i, row in enumerate(np.asarray(onsets))]}


def write_combined(folder, audio: np.ndarray, sr: int, onsets, name: str = "combined", zone: str = "center") -> dict:
    """notebooks/refresh.org:281-287: <folder>/combined.wav + combined.json (sf.write + json.dump(onsets_to_hits))."""
    folder = Path(folder)
    folder.mkdir(parents=True, exist_ok=True)
    write_wav(folder / f"{name}.wav", audio, sr)
    d = onsets_to_hits(onsets, zone)
    with open(folder / f"{name}.json", "w") as f:
        json.dump(d, f)
    return d


def read_json(file) -> dict:
    """data.read_json (data.py:31-38)."""
    with open(file, "r") as f:
        return json.load(f)


def parse_hits(d):
    """data.parse_hits (data.py:41-52): the hits as a DataFrame, `conditions` unwrapped into columns.
    The argument is the mapping of COLUMNS the reference passes (e.g. a combined.json table); a session's
    list of hit dicts is accepted as well."""
    import pandas as pd

    if isinstance(d, list):
        rows = []
        for h in d:
            h = dict(h)
            h.update(h.pop("conditions", {}))
            rows.append(h)
        return pd.DataFrame(rows)
    d = dict(d)
    if "conditions" in d:
        for cond in d["conditions"]:
            d[cond] = d["conditions"][cond]
        del d["conditions"]
    return pd.DataFrame(d)


def read_session(folder, name: str):
    """-> (audio [N, C] float32, sr, onsets int64 [H, C], locations float64 [H, 2] with NaN where absent, meta)."""
    folder = Path(folder)
    audio, sr = read_wav(folder / f"{name}.wav")
    d = read_json(folder / f"{name}.json")
    onsets = np.array([h["onset_start"] for h in d["hits"]], dtype=np.int64)
    loc = np.array([h.get("location", [np.nan, np.nan]) for h in d["hits"]], dtype=np.float64)
    return audio, sr, onsets, loc, d["meta"]
