"""POSD session files (posd.py): wav + json round trip, the reference's reader contract, foreign PCM files."""
import json
import struct

import numpy as np
import pytest

from onset_fingerprinting_b200 import posd


def test_session_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    audio = rng.standard_normal((4801, 3)).astype(np.float32) * 0.1
    onsets = np.array([[100, 130, 90], [2000, 2010, -1], [4000, 4002, 4004]])
    loc = np.array([[1.5, -2.25], [np.nan, np.nan], [0.0, 7.0]])
    sensors = [(0.9, 140, 75), (0.9, 10, 55), (0.5, 100, 15)]
    d = posd.write_session(tmp_path, "s1", audio, 96000, onsets, loc, sensors, meta={"instrument": "snare"},
                           zones=["center", "edge", "rim"], extra=[{"velocity": 0.5, "conditions": {"wires": "on"}}] * 3)
    # exactly what MCPOSD.from_file reads (data.py:296-301)
    meta = json.load(open(tmp_path / "s1.json"))
    assert [h["onset_start"] for h in meta["hits"]] == onsets.tolist()
    assert meta["hits"][0]["location"] == [1.5, -2.25] and "location" not in meta["hits"][1]
    assert meta["meta"]["channels"]["ch2"]["location"] == [0.5, 100.0, 15.0] and meta["meta"]["instrument"] == "snare"
    a2, sr, on2, loc2, m2 = posd.read_session(tmp_path, "s1")
    assert sr == 96000 and np.array_equal(a2, audio) and np.array_equal(on2, onsets)
    assert np.array_equal(np.isnan(loc2), np.isnan(loc)) and np.array_equal(np.nan_to_num(loc2), np.nan_to_num(loc))
    assert m2 == d["meta"]
    df = posd.parse_hits(meta["hits"])
    assert list(df["wires"]) == ["on"] * 3 and list(df["zone"]) == ["center", "edge", "rim"]
    # the reference's column-mapping form (data.py:41-52)
    df2 = posd.parse_hits({"i": [0, 1], "zone": ["a", "b"], "conditions": {"wires": ["on", "off"]}})
    assert list(df2.columns) == ["i", "zone", "wires"] and list(df2["wires"]) == ["on", "off"]


def test_wav_header_and_mono(tmp_path):
    x = np.linspace(-1, 1, 101, dtype=np.float32)
    posd.write_wav(tmp_path / "m.wav", x, 44100)
    raw = (tmp_path / "m.wav").read_bytes()
    assert raw[:4] == b"RIFF" and struct.unpack("<I", raw[4:8])[0] == len(raw) - 8
    assert struct.unpack("<HHI", raw[20:28]) == (3, 1, 44100)  # IEEE float, mono
    y, sr = posd.read_wav(tmp_path / "m.wav")
    assert y.shape == (101,) and np.array_equal(y, x) and sr == 44100


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_reads_pcm_files(tmp_path, bits):
    n, c, sr = 50, 2, 48000
    v = (np.arange(n * c) - 40) * (1 << (bits - 8))
    if bits == 16:
        body = v.astype("<i2").tobytes()
    elif bits == 32:
        body = v.astype("<i4").tobytes()
    else:
        body = b"".join(struct.pack("<i", int(k))[:3] for k in v)
    fmt = struct.pack("<HHIIHH", 1, c, sr, sr * c * bits // 8, c * bits // 8, bits)
    chunks = b"fmt " + struct.pack("<I", 16) + fmt + b"LIST" + struct.pack("<I", 4) + b"abcd" + b"data" + struct.pack("<I", len(body)) + body
    (tmp_path / "p.wav").write_bytes(b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks)
    a, got_sr = posd.read_wav(tmp_path / "p.wav")
    assert got_sr == sr and a.shape == (n, c)
    assert np.allclose(a.reshape(-1), v / float(1 << (bits - 1)), atol=1e-7)


def test_spec_example_session(tmp_path):
    """The example session of the POSD draft (notebooks/dataset_spec_draft.org:333-397): its hits parse into the
    table the reference's parse_hits builds (conditions unwrapped), and a file written from it reads back."""
    example = {
        "meta": {"channels": {"SP": {"location": [0.95, 0], "coordinate_system": "polar"},
                              "OP": {"location": [0.95, 30], "coordinate_system": "polar"}},
                 "instrument": "snare", "manufacturer": "sonor", "model": "BG SDW 2.0", "size": "13x5.75",
                 "head_top": "ambassador", "head_bottom": "ambassador_ss", "rim": "triple-flange", "tuning": "low",
                 "player": "rodrigo", "context": "full kit"},
        "hits": [
            {"i": 0, "zone": "center", "onset_start": [0, 2, 3], "velocity": 0.0, "isolated": True, "pitch": 220,
             "conditions": {"wires": "on"}},
            {"i": 1, "zone": "center", "onset_start": [48000, 4805, 47900], "velocity": 1.0, "isolated": True,
             "pitch": 220, "conditions": {"wires": "on"}},
            {"i": 2, "zone": "edge", "onset_start": [96000, 96000, 96000], "velocity": 0.0, "isolated": True,
             "pitch": 219, "conditions": {"wires": "off"}},
        ],
    }
    df = posd.parse_hits(example["hits"])
    assert list(df.columns) == ["i", "zone", "onset_start", "velocity", "isolated", "pitch", "wires"]
    assert list(df["wires"]) == ["on", "on", "off"] and df["onset_start"][1] == [48000, 4805, 47900]
    (tmp_path / "session1.json").write_text(json.dumps(example))
    posd.write_wav(tmp_path / "session1.wav", np.zeros((100, 3), np.float32), 48000)
    audio, sr, onsets, loc, meta = posd.read_session(tmp_path, "session1")
    assert sr == 48000 and audio.shape == (100, 3) and onsets.tolist() == [h["onset_start"] for h in example["hits"]]
    assert np.isnan(loc).all() and meta["channels"]["OP"]["location"] == [0.95, 30]


def test_combined_json_writer(tmp_path):
    """notebooks/refresh.org:243-287: combined.wav + combined.json = {"hits": [{"i", "zone": "center", "onset_start"}]}."""
    import json

    from onset_fingerprinting_b200 import posd

    rng = np.random.default_rng(0)
    audio = rng.standard_normal((4000, 3)).astype(np.float32) * 0.1
    onsets = np.array([[100, 120, 90], [2000, 2010, 2020]])
    d = posd.write_combined(tmp_path / "Setup 1", audio, 96000, onsets)
    want = {"hits": [{"i": i, "zone": "center", "onset_start": o.tolist()} for i, o in enumerate(onsets)]}
    assert d == want
    assert json.load(open(tmp_path / "Setup 1" / "combined.json")) == want
    back, sr = posd.read_wav(tmp_path / "Setup 1" / "combined.wav")
    assert sr == 96000 and np.array_equal(back, audio)
    assert posd.parse_hits(d["hits"])["onset_start"].tolist() == onsets.tolist()
