"""Parity at the sizes BASELINE.json states, CUDA side, through the reference-shaped calls:
configs[0] = one 60 s 3-mic recording through detect_onsets_amplitude -> find_onset_groups -> fix_onsets ->
Multilaterate3D.locate against a golden recorded from the unmodified reference; configs[2] = 16-channel hits with
the exact option set the bench runs.  Bars: onsets / groups / refined onsets bit-exact, envelope 1e-5 relative,
coordinates 1e-4 relative (north_star) -- measured: ~1e-12."""
import hashlib

import numpy as np
import pytest

from onset_fingerprinting_b200 import synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_config0_60s_chain(golden_dir):
    from onset_fingerprinting_b200 import detection as det
    from onset_fingerprinting_b200 import multilateration as ml
    from onset_fingerprinting_b200 import pipeline
    from oracle.make_golden import CONFIG0

    g = np.load(golden_dir / "config0_60s.npz")
    x, _ = synth.drum_recording(**CONFIG0)
    assert sha(x) == str(g["x_sha"])
    ch, on, rel = det.detect_onsets_amplitude(x, sr=96000)
    assert len(on) == 1512 and list(ch) == g["channels"].tolist() and list(on) == g["onsets"].tolist()
    assert rel.shape == tuple(g["rel_shape"])
    ref = g["rel_sub"]
    assert (np.abs(rel[::64] - ref) / np.maximum(np.abs(ref), 1e-6)).max() <= 1e-5
    groups = det.find_onset_groups(on, ch, 1000, 3)
    assert np.array_equal(groups, g["groups"])
    fixed = det.fix_onsets(x, groups)
    assert np.array_equal(fixed, g["fixed"])
    m = ml.Multilaterate3D(synth.SENSORS_3MIC, sr=96000, medium="air")
    want = g["xy"]
    ok = np.isfinite(want[:, 0])
    for h in range(0, len(fixed), 7):  # the streaming call, hit by hit, as the golden was recorded
        m.ongoing = []
        r = None
        for s in np.argsort(fixed[h], kind="stable"):
            r = m.locate(int(s), int(fixed[h][s]))
        assert (r is not None) == bool(ok[h])
        if r is not None:
            assert np.allclose(r, want[h], rtol=1e-9, atol=1e-9)
    # the batched device pipeline on the same recording (R = 1): same groups, refined onsets and positions
    hb = pipeline.HotPath(1, 3, synth.SENSORS_3MIC, medium="air", sr=96000).run(torch.from_numpy(x[None]).cuda(),
                                                                                 return_rel=True)
    assert np.array_equal(hb.onsets.cpu().numpy(), g["groups"])
    assert np.array_equal(hb.fixed.cpu().numpy(), g["fixed"])
    xy, st = hb.xy.cpu().numpy(), hb.loc_status.cpu().numpy()
    assert np.array_equal(st == 0, ok) and int(ok.sum()) == 407
    assert (np.abs(xy[ok] - want[ok]) <= 1e-4 * np.maximum(np.abs(want[ok]), 1e-3)).all()
    assert np.abs(xy[ok] - want[ok]).max() <= 2e-8
    assert (np.abs(hb.rel.cpu().numpy()[0, ::64] - ref) / np.maximum(np.abs(ref), 1e-6)).max() <= 1e-5


def test_hits16_bench_options(golden_dir):
    """configs[2]'s option set (tol 150, cutoff 20, d=1, abs, median 7, L=768): all 240 hits equal to the reference
    with numpy's scalar (stable) argsort, 230 equal to the reference as run with the SIMD sort kernels -- the ten
    that differ all hold tied onsets (H8, oracle/make_golden.py:gen_hits16)."""
    from onset_fingerprinting_b200 import detection as det
    from oracle.make_golden import HITS16_OPTS, hits16_sections

    g = np.load(golden_dir / "hits16_bench_opts.npz")
    xs, on = hits16_sections()
    assert sha(xs) == str(g["x_sha"]) and sha(on) == str(g["on_sha"])
    fixed, lags, st = det.fix_onsets_batch(torch.from_numpy(xs).cuda(), None, torch.from_numpy(on.astype(np.int32)).cuda(),
                                           max_section=xs.shape[1], **HITS16_OPTS)
    fixed, st = fixed.cpu().numpy(), st.cpu().numpy()
    assert np.array_equal(st == 2, g["raised_scalar"] == 1)
    assert np.array_equal(fixed, g["fixed_scalar"])
    same = (fixed == g["fixed"]).all(1)
    assert int(same.sum()) == 230 and g["tied"][~same].all()
    # and hit by hit through the reference-shaped call
    for h in range(0, len(on), 16):
        assert np.array_equal(det.fix_onsets(xs[h], on[h:h + 1], **HITS16_OPTS)[0], g["fixed_scalar"][h])
