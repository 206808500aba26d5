"""K1 parity on the GPU: CUDA path (through the C ABI) vs the CPU oracle and the golden vectors
the unmodified reference produced.  Bars: onset channel/sample lists bit-exact; envelope <= 1e-5
relative (north_star); additionally the fraction of bit-identical envelope samples vs the oracle
is reported and required to be > 99.99 % (both round log10/10**x once from double)."""
import hashlib

import numpy as np
import pytest

from onset_fingerprinting_b200 import synth

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def rel_err(a, b):
    return float((np.abs(a - b) / np.maximum(np.abs(b), 1e-6)).max())


@pytest.fixture(scope="module")
def det():
    from onset_fingerprinting_b200 import detection

    return detection


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle

    return oracle


@pytest.mark.parametrize("name", ["default3", "realtime3", "manual_b100", "b256_partial", "mesh16"])
def test_offline_vs_golden_and_oracle(name, det, orc, golden_dir):
    from oracle.make_golden import DETECT_CASES

    g = np.load(golden_dir / f"detect_{name}.npz")
    skw, dkw = DETECT_CASES[name]
    x, _ = synth.drum_recording(**skw)
    assert sha(x) == str(g["x_sha"])
    ch, on, rel = det.detect_onsets_amplitude(x, sr=96000, **dkw)
    # reference (golden)
    assert ch == g["channels"].tolist()
    assert on == g["onsets"].tolist()
    assert rel.shape == tuple(g["rel_shape"])
    assert rel_err(rel[::16], g["rel_sub"]) <= 1e-5
    # oracle, full resolution
    ch_o, on_o, rel_o = orc.detect_onsets_amplitude(x, sr=96000, **dkw)
    assert ch == ch_o and on == on_o
    assert rel_err(rel, rel_o) <= 1e-5
    assert float((rel == rel_o).mean()) > 0.9999


def test_streaming_blocks_vs_golden(det, golden_dir):
    g = np.load(golden_dir / "stream_realtime.npz")
    x, _ = synth.drum_recording(seconds=1.5, seed=7, first_hit=20000)
    od = det.AmplitudeOnsetDetector(3, 128, hipass_freq=0, fast_ar=(0.3, 800), slow_ar=(8000, 8000),
                                    on_threshold=0.45, off_threshold=0.45, cooldown=1323, sr=96000)
    got, rels = [], []
    for b, i in enumerate(range(0, len(x) - 127, 128)):
        c, d, r = od(x[i:i + 128])
        if b % 8 == 0:
            rels.append(r[::16])
        got += [(b, int(cc), int(dd)) for cc, dd in zip(c, d)]
    assert got == list(zip(g["blocks"].tolist(), g["channels"].tolist(), g["deltas"].tolist()))
    assert rel_err(np.asarray(rels), g["rel_sub"]) <= 1e-5


def test_batch_matches_per_recording_oracle(det, orc):
    """[R, N, C] batch in one launch == R independent oracle runs (ragged tail warp: R=13 -> 2 warps)."""
    xs, _ = synth.drum_batch(13, seconds=1.2, seed=100)
    ch, ix, cnt, rel = det.detect_onsets_amplitude_batch(xs, sr=96000)
    ch, ix, cnt, rel = ch.cpu().numpy(), ix.cpu().numpy(), cnt.cpu().numpy(), rel.cpu().numpy()
    for r in range(len(xs)):
        c_o, o_o, rel_o = orc.detect_onsets_amplitude(xs[r], sr=96000)
        k = cnt[r]
        assert ch[r, :k].tolist() == c_o and ix[r, :k].tolist() == o_o
        assert rel_err(rel[r], rel_o) <= 1e-5


def test_onsets_only_mode_and_state_carry(det, orc):
    """rel_dev = NULL gives the same onsets; two half-length calls == one call (state in the handle)."""
    x, _ = synth.drum_recording(seconds=2.0, seed=9)
    n = (len(x) // 256) * 256
    x = x[:n]
    d1 = det.BatchedOnsetDetector(1, 3, 128, sr=96000)
    ch, ix, cnt, _ = d1.detect_offline(x[None], warm_n=48000, return_rel=False)
    c_o, o_o, _ = orc.detect_onsets_amplitude(x, sr=96000)
    k = int(cnt[0])
    assert ch[0, :k].tolist() == c_o and ix[0, :k].tolist() == o_o
    d2 = det.BatchedOnsetDetector(1, 3, 128, sr=96000)
    h = n // 2
    a = d2.detect_offline(x[None, :h], warm_n=48000, return_rel=False)
    b = d2.detect_offline(x[None, h:], warm_n=0, return_rel=False)
    ka, kb = int(a[2][0]), int(b[2][0])
    joined = a[1][0, :ka].tolist() + [v + h for v in b[1][0, :kb].tolist()]
    assert joined == o_o


def test_unaligned_input_uses_generic_loader(det, orc):
    """A view whose base is not 16-byte aligned cannot use TMA; results must not change."""
    x, _ = synth.drum_recording(seconds=1.0, seed=12)
    buf = torch.empty(x.size + 1, dtype=torch.float32, device="cuda")
    view = buf[1:].view(1, *x.shape)
    view.copy_(torch.from_numpy(x))
    assert view.data_ptr() % 16 != 0
    d = det.BatchedOnsetDetector(1, 3, 128, sr=96000)
    ch, ix, cnt, rel = d.detect_offline(view, warm_n=48000)
    c_o, o_o, rel_o = orc.detect_onsets_amplitude(x, sr=96000)
    k = int(cnt[0])
    assert ch[0, :k].tolist() == c_o and ix[0, :k].tolist() == o_o
    assert rel_err(rel[0].cpu().numpy(), rel_o) <= 1e-5


def test_dll_twins(det, golden_dir):
    g = np.load(golden_dir / "kernels.npz")
    ar = det.AREnvelopeFollower(np.full((64, 5), -70, np.float32), 3, 383)
    for xb, ref in zip(g["ar_in"], g["ar_out"]):
        assert np.array_equal(ar(np.ascontiguousarray(xb)), ref)
    mm = det.MinMaxEnvelopeFollower(np.array([[0, 10]] * 5).T, alpha_min=1e-4, alpha_max=1e-5, minmin=2)
    for xb, ref in zip(g["mm_in"], g["mm_out"]):
        mn, mx = mm(np.ascontiguousarray(xb))
        assert np.array_equal(mn, ref[0]) and np.array_equal(mx, ref[1])


def test_host_buffer_entry_point(det, orc):
    import ctypes as C

    from onset_fingerprinting_b200 import _lib

    xs, _ = synth.drum_batch(3, seconds=1.0, seed=40)
    p = det.make_params(3, 128, sr=96000)
    R, N, Cn = xs.shape
    cap = 64
    rel = np.empty((R, (N // 128) * 128, Cn), np.float32)
    ch = np.empty((R, cap), np.int32); ix = np.empty((R, cap), np.int32); cnt = np.empty(R, np.int32)
    _lib.check(_lib.lib().ofp_detect_offline_host(
        C.byref(p), xs.ctypes.data_as(C.c_void_p), C.c_int64(R), C.c_int64(N), C.c_int64(48000),
        rel.ctypes.data_as(C.c_void_p), ch.ctypes.data_as(C.c_void_p), ix.ctypes.data_as(C.c_void_p),
        cnt.ctypes.data_as(C.c_void_p), C.c_int32(cap)))
    for r in range(R):
        c_o, o_o, rel_o = orc.detect_onsets_amplitude(xs[r], sr=96000)
        assert ch[r, :cnt[r]].tolist() == c_o and ix[r, :cnt[r]].tolist() == o_o
        assert rel_err(rel[r], rel_o) <= 1e-5
    _lib.check(_lib.lib().ofp_host_release())  # the staging buffers the call keeps for its next use
    _lib.check(_lib.lib().ofp_host_release())  # idempotent


def test_host_entry_time_segments_equal_one_shot(det, monkeypatch):
    """ofp_detect_offline_host feeds the batch in time segments (ofp_detect_continue); whatever the segment
    length, onsets and envelope must be bit-identical to one ofp_detect_offline launch over the whole batch.
    Recording length deliberately not a multiple of the block size."""
    import ctypes as C

    from onset_fingerprinting_b200 import _lib

    xs, _ = synth.drum_batch(5, seconds=1.3, seed=41)
    xs = np.ascontiguousarray(xs[:, : xs.shape[1] - 51])
    R, N, Cn = xs.shape
    assert N % 128 != 0
    d = det.BatchedOnsetDetector(R, 3, 128, sr=96000)
    ch0, ix0, cnt0, rel0 = d.detect_offline(torch.from_numpy(xs).cuda(), warm_n=48000)
    ch0, ix0, cnt0, rel0 = ch0.cpu().numpy(), ix0.cpu().numpy(), cnt0.cpu().numpy(), rel0.cpu().numpy()
    p = det.make_params(3, 128, sr=96000)
    cap = ch0.shape[1]
    for seg, chunk in ((48000, 8192), (50048, 8192), (1 << 20, 8192), (48000, 2)):
        monkeypatch.setenv("OFP_HOST_SEGMENT", str(seg))
        monkeypatch.setenv("OFP_HOST_CHUNK", str(chunk))
        rel = np.full((R, (N // 128) * 128, Cn), np.nan, np.float32)
        ch = np.empty((R, cap), np.int32); ix = np.empty((R, cap), np.int32); cnt = np.empty(R, np.int32)
        _lib.check(_lib.lib().ofp_detect_offline_host(
            C.byref(p), xs.ctypes.data_as(C.c_void_p), C.c_int64(R), C.c_int64(N), C.c_int64(48000),
            rel.ctypes.data_as(C.c_void_p), ch.ctypes.data_as(C.c_void_p), ix.ctypes.data_as(C.c_void_p),
            cnt.ctypes.data_as(C.c_void_p), C.c_int32(cap)))
        assert np.array_equal(cnt, cnt0), (seg, chunk)
        for r in range(R):
            assert np.array_equal(ch[r, :cnt[r]], ch0[r, :cnt[r]]) and np.array_equal(ix[r, :cnt[r]], ix0[r, :cnt[r]])
        assert np.array_equal(rel, rel0), (seg, chunk)
    assert int(cnt0.sum()) >= 5 * 6  # the batch does contain hits
    _lib.check(_lib.lib().ofp_host_release())


def test_backtrack_offline_and_streaming(det, golden_dir):
    g = np.load(golden_dir / "backtrack.npz")
    x, _ = synth.drum_recording(seconds=2.0, seed=8)
    assert sha(x) == str(g["x_sha"])
    ch, on, _ = det.detect_onsets_amplitude(x, sr=96000, backtrack=True, backtrack_buffer_size=128,
                                            backtrack_smooth_size=5)
    assert ch == g["ch_b128"].tolist() and on == g["on_b128"].tolist()
    ch, on, _ = det.detect_onsets_amplitude(x, sr=96000, backtrack=True, backtrack_buffer_size=256,
                                            backtrack_smooth_size=1)
    assert on == g["on_b256s1"].tolist()
    # block-by-block detector with the same settings and warm-up gives the same backtracked onsets
    od = det.AmplitudeOnsetDetector(3, 128, sr=96000, backtrack=True, backtrack_buffer_size=256,
                                    backtrack_smooth_size=1)
    od.init_minmax_tracker(x[:48000])
    got = []
    for i in range(0, len(x) - 127, 128):
        c, d, _ = od(x[i:i + 128])
        got += [i + int(v) for v in d]
    assert got == g["on_b256s1"].tolist()


def test_loud_clipped_signal_exercises_exact_paths(det, orc):
    """Full-scale, clipped audio: dB values near 0 put the followers into the regime where the float32
    shortcut of the follower step is not proven exact, so the kernel re-runs those chunks through its
    exact path; plateaus at +-1.0 hit log10 inputs next to 1.  Results must still equal the oracle."""
    x, _ = synth.drum_recording(seconds=1.2, seed=77)
    x = np.clip(x * 6.0, -1.0, 1.0).astype(np.float32)
    for kw in (dict(), dict(hipass_freq=0, fast_ar=(0.3, 800), slow_ar=(8000, 8000), on_threshold=0.45,
                            off_threshold=0.45)):
        ch, on, rel = det.detect_onsets_amplitude(x, sr=96000, **kw)
        ch_o, on_o, rel_o = orc.detect_onsets_amplitude(x, sr=96000, **kw)
        assert ch == ch_o and on == on_o
        assert rel_err(rel, rel_o) <= 1e-5
        assert float((rel == rel_o).mean()) > 0.9999


@pytest.mark.parametrize("kw", [
    dict(fast_ar=(1.0, 1.5), slow_ar=(1.2, 1.8)),            # release coefficients above 1/2: every chunk on the exact path
    dict(fast_ar=(0.05, 300.0)),                              # attack coefficient 20 (> 16): exact path
    dict(floor=-3.0),                                         # floor above the sliver threshold: exact path
    dict(fast_ar=(0.3, 800), slow_ar=(500.0, 8000.0)),        # attack above 1 on the straight-line path, att != rel slow
    dict(fast_ar=(400.0, 3.0), slow_ar=(2205.0, 100.0)),      # release faster than attack (the sign-flipped max form)
])
def test_follower_coefficient_ranges(kw, det, orc):
    """The straight-line chunk is only enabled for release coefficients in (0, 1/2], attack coefficients in
    (0, 16] and floors in [-180, -4] dB (csrc/onset_detect.cu: fast_ok); outside, and on the boundary cases inside,
    the kernel must still equal the oracle bit for bit."""
    x, _ = synth.drum_recording(seconds=1.5, seed=5)
    ch, on, rel = det.detect_onsets_amplitude(x, sr=96000, **kw)
    ch_o, on_o, rel_o = orc.detect_onsets_amplitude(x, sr=96000, **kw)
    assert ch == ch_o and on == on_o
    assert rel_err(rel, rel_o) <= 1e-5
    assert float((rel == rel_o).mean()) > 0.9999


def test_detector_edge_cases(det, orc):
    """Inputs at the edges of the offline driver (detection.py:55-86): silence (no onset at all), a recording shorter
    than the 0.5 s warm-up, one shorter than a block (no block processed: empty outputs), exactly one block, and a
    batch of one recording; every case equals the oracle's answer, empty lists included."""
    rng = np.random.default_rng(7)
    sr, B = 96000, 128
    full, _ = synth.drum_recording(seconds=1.0, seed=21)
    cases = {
        "silence": np.zeros((sr, 3), np.float32),
        "short_of_warmup": full[48000 + 900:48000 + 900 + 10 * B + 17].copy(),  # a burst inside, < 0.5 s
        "less_than_a_block": (1e-3 * rng.standard_normal((B - 1, 3))).astype(np.float32),
        "one_block": (1e-3 * rng.standard_normal((B, 3))).astype(np.float32),
        "one_recording": full,
    }
    for name, x in cases.items():
        c_o, o_o, rel_o = orc.detect_onsets_amplitude(x, sr=sr)
        ch, ix, cnt, rel = det.detect_onsets_amplitude_batch(x[None], sr=sr)
        k = int(cnt[0])
        assert ch[0, :k].tolist() == list(c_o) and ix[0, :k].tolist() == list(o_o), name
        rel = rel[0].cpu().numpy()
        assert rel.shape == np.asarray(rel_o).shape, (name, rel.shape, np.asarray(rel_o).shape)
        if rel.size:
            assert rel_err(rel, rel_o) <= 1e-5, name
        # the reference's own entry point (lists + ndarray)
        c1, o1, rel1 = det.detect_onsets_amplitude(x, sr=sr)
        assert list(c1) == list(c_o) and list(o1) == list(o_o), name


@pytest.mark.parametrize("n_ch,block", [(1, 128), (2, 64), (5, 128), (7, 96), (32, 32)])
def test_channel_counts_and_lane_packing(n_ch, block, det, orc):
    """Channel counts that pack the warp differently (G = 32, 16, 6, 4, 1 recordings per warp; 32, 32, 30, 28, 32
    active lanes) on a ragged batch of R = 2 G + 1 recordings: onsets identical to the oracle per recording,
    envelope <= 1e-5 relative.  Bursts reach the channels with different delays and gains, so the cross-channel
    off-threshold coupling is exercised for every packing."""
    rng = np.random.default_rng(100 + n_ch)
    sr, n = 96000, 96000 + 4 * block + 5
    R = 2 * (32 // n_ch) + 1 if n_ch < 32 else 3
    R = min(R, 9)
    t = np.arange(3000) / sr
    burst = np.exp(-400 * t) * np.sin(2 * np.pi * 900 * t)
    xs = (1e-4 * rng.standard_normal((R, n, n_ch))).astype(np.float32)
    for r in range(R):
        for s in range(50000 + 37 * r, n - 4000, 11000 + 501 * r):
            for c in range(n_ch):
                a = s + int(rng.integers(0, 60))
                xs[r, a:a + 3000, c] += (rng.uniform(0.05, 0.5) * burst).astype(np.float32)
    ch, ix, cnt, rel = det.detect_onsets_amplitude_batch(xs, block_size=block, sr=sr)
    ch, ix, cnt, rel = ch.cpu().numpy(), ix.cpu().numpy(), cnt.cpu().numpy(), rel.cpu().numpy()
    total = 0
    for r in range(R):
        c_o, o_o, rel_o = orc.detect_onsets_amplitude(xs[r], block_size=block, sr=sr)
        k = int(cnt[r])
        assert ch[r, :k].tolist() == list(c_o) and ix[r, :k].tolist() == list(o_o), (n_ch, r)
        assert rel_err(rel[r], rel_o) <= 1e-5, (n_ch, r)
        total += k
    assert total >= R * n_ch  # every channel of every recording fired at least once
