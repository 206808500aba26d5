"""oracle/ref_chain.py (the numpy/scipy restatement of the post-detector stages that bench.py times as the
reference's CPU path) against goldens recorded from the unmodified reference."""
import numpy as np

from onset_fingerprinting_b200 import synth
from oracle import ref_chain, ref_style
from oracle.make_golden import CONFIG0, FIX_OPTS, HITS16_OPTS, hits16_sections


def test_chain_on_config0_prefix(golden_dir):
    """First 12 s of the configs[0] recording: detector (reference DLL + numpy block loop), grouping, lag
    refinement and multilateration give the golden's onsets, refined onsets and positions."""
    g = np.load(golden_dir / "config0_60s.npz")
    x, _ = synth.drum_recording(**CONFIG0)
    n = 12 * 96000
    ch, on, _ = ref_style.detect_onsets_amplitude(x[:n], sr=96000)
    k = len(on)
    assert k > 250 and on == g["onsets"][:k].tolist() and ch == g["channels"][:k].tolist()
    groups = ref_chain.find_onset_groups(on, ch, 1000, 3)
    h = len(groups) - 1  # the last group may be cut by the prefix
    assert np.array_equal(groups[:h], g["groups"][:h])
    fixed = ref_chain.fix_onsets(x[:n], groups[:h])
    assert np.array_equal(fixed, g["fixed"][:h])
    loc = ref_chain.Locator(synth.SENSORS_3MIC, sr=96000, medium="air")
    for row, want in zip(fixed, g["xy"][:h]):
        got = loc.locate_hit(row)
        assert (got is None) == bool(np.isnan(want[0]))
        if got is not None:
            assert np.allclose(got, want, rtol=1e-9, atol=1e-9)


def test_fix_onsets_options(golden_dir):
    g = np.load(golden_dir / "fix_3ch.npz")
    x, _ = synth.drum_recording(seconds=3.0, seed=21)
    for name, kw in FIX_OPTS.items():
        assert np.array_equal(ref_chain.fix_onsets(x, g["groups"], **kw), g[f"fixed_{name}"]), name
    g = np.load(golden_dir / "hits16_bench_opts.npz")
    xs, on = hits16_sections()  # the generator's stream depends on n_hits: take the golden's 240
    for h in range(40):
        assert np.array_equal(ref_chain.fix_onsets(xs[h], on[h:h + 1], **HITS16_OPTS)[0], g["fixed_scalar"][h])
