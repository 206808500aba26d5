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


def test_unsupported_options_raise():
    from onset_fingerprinting_b200 import model

    with pytest.raises(NotImplementedError):
        model.CNN(256, 2, batch_norm=True)
    with pytest.raises(NotImplementedError):
        model.CNN(256, 2, pool=True)
