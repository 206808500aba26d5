"""Onset-window extraction (data.py path) on the GPU against what the reference's
FrameExtractor / FastFrameExtractor / batch_cc / MCPOSD returned (tests/golden/frames.npz) and against
numpy's sliding_window_view at batch scale."""
import hashlib

import numpy as np
import pytest

from onset_fingerprinting_b200 import synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir / "frames.npz")


@pytest.fixture(scope="module")
def inputs():
    from oracle.make_golden import frames_inputs

    return frames_inputs()


def test_frame_extractor_variants(gold, inputs):
    from onset_fingerprinting_b200 import data

    x, on = inputs
    assert sha(x) == str(gold["x_sha"])
    assert np.array_equal(data.FrameExtractor(256, 16)(x, on), gold["fe_min"])
    assert np.array_equal(data.FrameExtractor(128, 8, add_pre_samples=True, use_min_onset=False)(x, on), gold["fe_each"])
    assert np.array_equal(data.FrameExtractor(200, 0)(x[:, 1].copy(), on[:, 1]), gold["fe_1d"])
    np.random.seed(5)
    assert np.array_equal(data.FrameExtractor(256, 32, max_shift=10)(x, on), gold["fe_shift"])


def test_fast_extractor_dataset_and_batch_cc(gold, inputs):
    from onset_fingerprinting_b200 import data

    x, on = inputs
    ffe = data.FastFrameExtractor(x, on, 256, 16)
    assert sha(ffe().cpu().numpy()) == str(gold["ffe_sha"])
    shifted = data.FastFrameExtractor(x, on, 256, 16, max_shift=5)()
    assert tuple(shifted.shape) == (40, 3, 256)
    pos = np.random.default_rng(1).uniform(-1, 1, (40, 2))
    ds = data.MCPOSD(x, on, pos, 256, 16)
    xx, yy = ds[0]
    assert sha(xx.cpu().numpy()) == str(gold["ds_x_sha"]) and np.array_equal(yy.cpu().numpy(), gold["ds_y"])
    tr, te = ds.split(0.75)
    assert len(tr.y) == 30 and len(te.y) == 10
    cc = data.batch_cc(np.ascontiguousarray(x[1000:1256].T), np.ascontiguousarray(x[1010:1266].T))
    assert np.allclose(cc, gold["bcc"], rtol=2e-5, atol=1e-6)  # conv1d sums in float32, ours in double


def test_out_of_range_raises_and_negative_wraps():
    from onset_fingerprinting_b200 import data

    x = np.arange(300, dtype=np.float32).reshape(100, 3)
    view = np.lib.stride_tricks.sliding_window_view(x, 10, axis=0)
    fe = data.FrameExtractor(10, 4)
    on = np.array([[2, 5, 6], [50, 51, 52]])
    assert np.array_equal(fe(x, on), view[on.min(1) - 4])  # 2 - 4 = -2 wraps like numpy
    with pytest.raises(IndexError):
        fe(x, np.array([[95, 99, 99]]))


def test_batch_gather_after_hot_path():
    """[R, N, C] batch: windows of every located group cut on the device == numpy per recording."""
    from onset_fingerprinting_b200 import data, pipeline

    xs, _ = synth.drum_batch(6, seconds=1.2, seed=300)
    hp = pipeline.HotPath(6, 3, synth.SENSORS_3MIC, medium="air", sr=96000)
    hb = hp.run(torch.from_numpy(xs).cuda())
    fr = data.extract_frames_batch(torch.from_numpy(xs).cuda(), hb.rec, hb.fixed, 256, 32).cpu().numpy()
    rec, fixed = hb.rec.cpu().numpy(), hb.fixed.cpu().numpy()
    assert len(rec) > 20
    for h in range(len(rec)):
        view = np.lib.stride_tricks.sliding_window_view(xs[rec[h]], 256, axis=0)
        assert np.array_equal(fr[h], view[fixed[h].min() - 32])
