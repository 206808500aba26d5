"""K6 CNN inference vs the plain PyTorch fp32 modules of the same architecture (model.py:52-120).
Floating-point bar: both sides accumulate in float32 (different summation order over the 4096-term
Linear layer and SiLU via __expf): 2e-4 relative to the output scale."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def torch_reference(m, x):
    """The reference forward (model.py:112-117) with stock torch ops in float32, TF32 off."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            h = m.conv_layers(x)
            h = torch.flatten(h, start_dim=1)
            return m.fc(h)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("cfg", [
    dict(input_size=256, output_size=2),                                             # the reference defaults
    dict(input_size=256, output_size=2, layer_sizes=[8, 16, 32], kernel_size=5, padding=2),
    dict(input_size=200, output_size=3, channels=4, layer_sizes=[6, 10], kernel_size=3, padding=0),
    dict(input_size=64, output_size=9, channels=16, layer_sizes=[12], kernel_size=7, padding=3,
         activation=torch.nn.ReLU),
    dict(input_size=128, output_size=2, layer_sizes=[8, 16], kernel_size=3, padding=4, activation=torch.nn.Tanh),
])
def test_cnn_forward_matches_torch(cfg):
    from onset_fingerprinting_b200 import model

    torch.manual_seed(0)
    m = model.CNN(**cfg).cuda()
    x = torch.randn(1037, cfg.get("channels", 3), cfg["input_size"], device="cuda")
    got = m(x)
    want = torch_reference(m, x)
    scale = float(want.abs().max())
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-4 * scale, float((got - want).abs().max()) / scale


def test_state_dict_round_trip_and_views():
    from onset_fingerprinting_b200 import model

    torch.manual_seed(1)
    a = model.CNN(256, 2).cuda()
    b = model.CNN(256, 2).cuda()
    x = torch.randn(64, 3, 256, device="cuda")
    ya = a(x)
    assert float((ya - b(x)).abs().max()) > 1e-3  # different random initialisation
    b.load_state_dict(a.state_dict())
    assert torch.equal(ya, b(x))
    # strided batch (every other window of a larger buffer) and numpy input
    big = torch.randn(128, 3, 256, device="cuda")
    assert torch.equal(a(big[::2]), a(big[::2].contiguous()))
    assert np.array_equal(a.call_np(x.cpu().numpy()), ya.cpu().numpy())


@pytest.mark.parametrize("cfg", [
    dict(input_size=256, output_size=2, batch_norm=True),
    dict(input_size=256, output_size=2, pool=True),
    dict(input_size=256, output_size=2, batch_norm=True, pool=True, layer_sizes=[8, 16, 24]),
    dict(input_size=250, output_size=3, channels=4, layer_sizes=[8, 12], kernel_size=5, padding=2, dilation=3),
    dict(input_size=128, output_size=2, channels=4, layer_sizes=[8, 16], groups=2, pool=True),
    dict(input_size=255, output_size=5, layer_sizes=[6, 9], kernel_size=3, padding=0, dilation=2, pool=True,
         batch_norm=True, activation=torch.nn.ReLU),                                # odd lengths: pooling drops the last
])
def test_cnn_constructor_options_match_torch(cfg):
    """batch_norm / pool / dilation / groups of the reference's CNN (model.py:62-66, 91-108), eval mode, with
    non-trivial BatchNorm running statistics and affine parameters."""
    from onset_fingerprinting_b200 import model

    torch.manual_seed(3)
    m = model.CNN(**cfg).cuda()
    for mod in m.conv_layers:
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.uniform_(-0.3, 0.3)
            mod.running_var.uniform_(0.5, 2.0)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.uniform_(-0.2, 0.2)
    x = torch.randn(515, cfg.get("channels", 3), cfg["input_size"], device="cuda")
    got = m(x)
    want = torch_reference(m, x)
    scale = float(want.abs().max())
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-4 * scale, float((got - want).abs().max()) / scale


def test_unsupported_options_raise():
    from onset_fingerprinting_b200 import model

    with pytest.raises(NotImplementedError):
        model.CNN(256, 2, activation=torch.nn.GELU)
    with pytest.raises(NotImplementedError):
        model.CCCNN(256, 2, batch_norm=True, group=True)


def _fcnn_reference(m, x):
    with torch.no_grad():
        return m.network(x)


@pytest.mark.parametrize("kw", [dict(), dict(hidden_layers=[16, 32, 8, 4], activation=torch.nn.Tanh),
                                dict(batch_norm=False, activation=torch.nn.SiLU), dict(bias=False, eye_init=True)])
def test_fcnn_forward_matches_torch(kw):
    """calibration.FCNN (calibration.py:463-560) in eval mode; BatchNorm with non-trivial running statistics.
    float32 on both sides, BatchNorm folded to scale/shift here: 1e-5 relative to the output scale."""
    from onset_fingerprinting_b200 import calibration

    torch.manual_seed(3)
    m = calibration.FCNN(2, 2, **kw)
    for mod in m.network:
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.normal_(0, 0.5); mod.running_var.uniform_(0.5, 2.0)
            mod.weight.data.normal_(1, 0.2); mod.bias.data.normal_(0, 0.2)
    m = m.cuda().eval()
    x = (torch.randn(5000, 2, device="cuda") * 60).round()
    got, want = m(x), _fcnn_reference(m, x)
    assert float((got - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))
    one = m.call_np((12.0, -7.0))
    assert np.allclose(one, _fcnn_reference(m, torch.tensor([[12.0, -7.0]], device="cuda")).cpu().numpy()[0], rtol=1e-5, atol=1e-6)


def test_multilaterate3d_model_bypass():
    """Multilaterate3D(model=FCNN): trilaterate returns model((d_a1, d_b1)) * 100 (multilateration.py:553-557)
    after the same legality checks; the batched path must agree with the per-hit one and with the solver
    path on which hits are located at all (status 4 = solver failure cannot occur with a model)."""
    from onset_fingerprinting_b200 import calibration
    from onset_fingerprinting_b200 import multilateration as ml

    torch.manual_seed(4)
    net = calibration.FCNN(2, 2).cuda()
    sensors = [(0.9, 140, 75), (0.9, 10, 55), (0.5, 100, 15)]
    a = ml.Multilaterate3D(sensors, sr=96000, medium="air")
    b = ml.Multilaterate3D(sensors, sr=96000, medium="air", model=net)
    rng = np.random.default_rng(0)
    base = rng.integers(1000, 100000, size=(2000, 1))
    on = np.concatenate([base + rng.integers(0, 60, size=(2000, 2)), base], axis=1).astype(np.int32)  # close mic first
    xy_a, st_a = a.locate_batch(on)
    xy_b, st_b = b.locate_batch(on)
    st_a, st_b = st_a.cpu().numpy(), st_b.cpu().numpy()
    assert ((st_a == 0) | (st_a == 4)).sum() == (st_b == 0).sum()
    assert np.array_equal(st_b[st_a != 4], st_a[st_a != 4])
    ok = np.nonzero(st_b == 0)[0]
    assert len(ok) > 100
    xy_b = xy_b.cpu().numpy()
    assert np.isnan(xy_b[st_b != 0]).all()
    for h in ok[:50]:
        order = np.argsort(on[h], kind="stable")
        got = b.trilaterate(([int(s) for s in order], [int(on[h][s]) for s in order]), initial_guess=np.zeros(2))
        assert np.allclose(got, xy_b[h], rtol=1e-6, atol=1e-6)


def test_session_file_to_windows_to_network(tmp_path):
    """The dataset tail of configs[1]/[4]: hot path results -> POSD session on disk (posd.py) -> MCPOSD.from_file
    (data.py:285-311) -> windows on the device -> CNN inference; the windows equal a direct numpy slice."""
    from onset_fingerprinting_b200 import data, model, pipeline, posd, synth

    xs, _ = synth.drum_batch(1, seconds=2.0, seed=21)
    hb = pipeline.HotPath(1, 3, synth.SENSORS_3MIC, medium="air", sr=96000).run(torch.from_numpy(xs).cuda())
    ok = (hb.loc_status == 0).cpu().numpy()
    assert ok.sum() >= 5
    fixed, xy = hb.fixed.cpu().numpy()[ok], hb.xy.cpu().numpy()[ok]
    posd.write_session(tmp_path, "take1", xs[0], 96000, fixed, xy, synth.SENSORS_3MIC, meta={"instrument": "snare"})
    ds = data.MCPOSD.from_file(tmp_path, "take1", frame_length=256, pre_samples=16)
    x, y = ds[0]
    assert tuple(x.shape) == (int(ok.sum()), 3, 256) and tuple(y.shape) == (int(ok.sum()), 2)
    start = fixed.min(1) - 16
    want = np.stack([xs[0][s:s + 256].T for s in start])
    assert np.array_equal(x.cpu().numpy(), want)
    assert np.allclose(y.cpu().numpy(), xy.astype(np.float32))
    torch.manual_seed(5)
    net = model.CNN(256, 2).cuda()
    out = net(x)
    assert tuple(out.shape) == (int(ok.sum()), 2) and bool(torch.isfinite(out).all())


def test_cnn_large_batch_properties(monkeypatch):
    """200 k windows (beyond what the torch reference is run on here): the tensor-core path and the generic FP32
    kernel agree to float32 accuracy, the result does not depend on the batch a window travels in, and it is
    deterministic."""
    from onset_fingerprinting_b200 import model

    torch.manual_seed(11)
    net = model.CNN(256, 2).cuda()
    x = torch.randn(200_000, 3, 256, device="cuda") * 0.2
    y_tc = net(x)
    assert torch.equal(y_tc, net(x))
    monkeypatch.setenv("OFP_K6_NO_TC", "1")
    y_fp = net(x)
    monkeypatch.delenv("OFP_K6_NO_TC")
    scale = float(y_fp.abs().max())
    assert float((y_tc - y_fp).abs().max()) <= 1e-4 * scale
    perm = torch.randperm(x.shape[0], device="cuda")
    assert torch.equal(net(x[perm]), y_tc[perm])
    assert torch.equal(net(x[1234:1234 + 777]), y_tc[1234:1234 + 777])


def test_chunked_host_pipeline_equals_direct_call():
    """hostpipe.run_chunked (uploads overlapped with compute on two streams, pinned results) returns exactly
    what one call on the whole batch returns; ragged last chunk, buffer reuse."""
    from onset_fingerprinting_b200 import hostpipe, model

    torch.manual_seed(2)
    net = model.CNN(256, 2).cuda()
    xh = torch.randn(1003, 3, 256).pin_memory()
    want = net(xh.cuda()).cpu()
    (got,) = hostpipe.run_chunked([xh], net, chunk=128)
    assert torch.equal(got, want)
    (again,) = hostpipe.run_chunked([xh], net, chunk=400, outs=(got,))
    assert again.data_ptr() == got.data_ptr() and torch.equal(again, want)


def cccnn_reference(m, x):
    """model.CCCNN.forward (model.py:512-538, group=False) with stock torch ops, TF32 off."""
    import torch.nn.functional as F

    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            B, Cn, W = x.shape
            if m.group:  # model.py:512-513: grouped convolutions over the C input channels -> (B, C*K, V)
                f = m.conv_layers(x)
                f = f.reshape(B * Cn, f.shape[1] // Cn, f.shape[2])
            else:
                f = m.conv_layers(x.reshape(B * Cn, 1, W))  # the reference vmaps the same stack over the channels
            K, V = f.shape[1:]
            cc = F.conv1d(f.reshape(1, B * Cn * K, V), f.reshape(B * Cn * K, 1, V), groups=B * Cn * K, padding=V - 1)
            cc = cc.view(B * Cn, K, -1).sum(1)
            p = torch.softmax(cc, dim=-1).view(B, -1)
            return m.fc(p)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("cfg", [
    dict(input_size=256, output_size=2),
    dict(input_size=128, output_size=3, channels=4, layer_sizes=[8], kernel_sizes=5, padding=2),
    dict(input_size=64, output_size=2, channels=2, layer_sizes=[6, 12, 8], activation=torch.nn.Tanh),
    dict(input_size=48, output_size=1, channels=3, layer_sizes=[16], activation=torch.nn.ReLU),  # V / 8 not a multiple of 4
    dict(input_size=256, output_size=2, group=True),                                              # own stack per sensor
    dict(input_size=64, output_size=2, channels=4, layer_sizes=[6, 8], group=True, activation=torch.nn.Tanh),
    # the constructor options the reference leaves off by default (model.py:451-457): generic kernel
    dict(input_size=256, output_size=2, pool=True),                                               # V = 64
    dict(input_size=256, output_size=2, batch_norm=True),                                         # GroupNorm(1, K)
    dict(input_size=256, output_size=2, layer_sizes=[8, 16], kernel_sizes=[5, 5], strides=[2, 1], padding=2),  # V = 128
    dict(input_size=131, output_size=3, channels=2, layer_sizes=[6, 8], kernel_sizes=[7, 3], strides=[1, 2], padding=3,
         dilation=2, batch_norm=True, pool=True, activation=torch.nn.ReLU),                       # 125 -> 62 -> 32 -> 16
    dict(input_size=128, output_size=2, channels=3, layer_sizes=[8, 8], pool=True, group=True),
])
def test_cccnn_forward_matches_torch(cfg):
    """The summed auto-correlation runs as F^T F on the tensor cores (3xTF32); softmax turns absolute errors of the
    correlation into relative errors of the probabilities, so the bar is 1e-3 of the output scale (float32 on both
    sides, different summation orders over V*K products).  Input scales from a near-uniform to a near-one-hot lag
    distribution; at least one of them must make the output depend on the window (not a vacuous comparison)."""
    from onset_fingerprinting_b200 import model

    torch.manual_seed(3)
    m = model.CCCNN(**cfg).cuda()
    with torch.no_grad():
        for mod in m.conv_layers:
            if isinstance(mod, torch.nn.GroupNorm):
                # normalised maps of weight ~1 put K V = 4096 into lag 0 and the softmax becomes one-hot whatever the
                # window holds: small weights keep the lag distribution (and the comparison) alive
                mod.weight.uniform_(0.04, 0.10)
                mod.bias.uniform_(-0.01, 0.01)
        m.fc.weight.mul_(30.0)  # default init is 1/sqrt(fan_in): make the lag distribution visible in the output
    spread = 0.0
    for scale in (0.05, 0.15, 0.3, 0.6, 1.2, 2.5):
        x = torch.randn(257, cfg.get("channels", 3), cfg["input_size"], device="cuda") * scale
        if cfg.get("batch_norm"):  # the norm removes the scale: give every window its own tone instead
            t = torch.arange(cfg["input_size"], device="cuda")
            f = torch.rand(257, cfg.get("channels", 3), 1, device="cuda") * 0.2 + 0.01
            x = x + 3.0 * scale * torch.sin(2 * torch.pi * f * t)
        got, want = m(x), cccnn_reference(m, x)
        assert got.shape == want.shape
        err = float((got - want).abs().max())
        assert err <= 1e-3 * max(float(want.abs().max()), 1e-3), (scale, err)
        spread = max(spread, float(want.std(0).max()))
    assert spread > 1e-3, spread
