"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol include/ofp.h declares; the product path fails loudly without CUDA."""
import numpy as np
import pytest
import torch

from onset_fingerprinting_b200 import _lib


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), n
    assert b"sm_100a" in L.ofp_version()


def test_params_struct_layout_matches_header():
    # 5 int32 + 10 float + 10 float = 25 words
    import ctypes

    assert ctypes.sizeof(_lib.DetectorParams) == 25 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    from onset_fingerprinting_b200 import detection

    with pytest.raises(_lib.OfpError):
        detection.detect_onsets_amplitude(np.zeros((4096, 3), np.float32))


def test_argument_validation_without_gpu():
    import ctypes as C

    L = _lib.lib()
    h = C.c_void_p()
    p = _lib.DetectorParams()
    p.n_channels, p.block_size = 40, 128  # > 32 channels is rejected before any CUDA call
    assert L.ofp_detector_create(C.byref(h), C.c_int64(1), C.byref(p)) == -1
    assert b"n_channels" in L.ofp_last_error()


@pytest.mark.parametrize("cfg", [
    dict(input_size=256, output_size=2),
    dict(input_size=256, output_size=2, batch_norm=True, pool=True, layer_sizes=[8, 16, 24]),
    dict(input_size=250, output_size=3, channels=4, layer_sizes=[8, 12], kernel_size=5, padding=2, dilation=3),
    dict(input_size=255, output_size=5, layer_sizes=[6, 9], kernel_size=3, padding=0, dilation=2, pool=True, batch_norm=True),
])
def test_cnn_param_count_matches_the_module(cfg):
    """ofp_cnn_param_count_ex (no GPU needed) agrees with the parameter containers model.CNN builds for the same
    constructor arguments: conv weights padded to 8 output channels + bias (+ norm scale / shift), then the Linear
    layer over the flattened features whose length the C side derives from the same length formula as torch."""
    import ctypes as C

    from onset_fingerprinting_b200 import model

    m = model.CNN(**cfg)
    n, flat = C.c_int64(0), C.c_int32(0)
    sizes = (C.c_int32 * len(m.layer_sizes))(*m.layer_sizes)
    rc = _lib.lib().ofp_cnn_param_count_ex(C.c_int32(m.channels), C.c_int32(m.input_size), C.c_int32(len(m.layer_sizes)),
                                           sizes, C.c_int32(m.kernel_size), C.c_int32(m.padding), C.c_int32(m.dilation),
                                           C.c_int32(m.pool), C.c_int32(m.batch_norm), C.c_int32(m.output_size),
                                           C.byref(n), C.byref(flat))
    assert rc == 0, _lib.lib().ofp_last_error()
    assert flat.value == m.fc.in_features
    with torch.no_grad():  # the module's own answer for the flattened length
        assert m.conv_layers(torch.zeros(1, m.channels, m.input_size)).flatten(1).shape[1] == flat.value
    want, cin = 0, m.channels
    for size in m.layer_sizes:
        cp = (size + 7) // 8 * 8
        want += cin * m.kernel_size * cp + cp + (2 * cp if m.batch_norm else 0)
        cin = size
    want += m.output_size * flat.value + m.output_size
    assert n.value == want


def test_cccnn_param_count_ex_lengths():
    """ofp_cccnn_param_count_ex: the number of lags 2V - 1 follows torch's Conv1d / MaxPool1d length arithmetic."""
    import ctypes as C

    from onset_fingerprinting_b200 import model

    for cfg in (dict(input_size=256, output_size=2, pool=True),
                dict(input_size=256, output_size=2, layer_sizes=[8, 16], kernel_sizes=[5, 5], strides=[2, 1], padding=2),
                dict(input_size=131, output_size=3, channels=2, layer_sizes=[6, 8], kernel_sizes=[7, 3], strides=[1, 2],
                     padding=3, dilation=2, batch_norm=True, pool=True)):
        m = model.CCCNN(**cfg)
        L = len(m.layer_sizes)
        n, lags = C.c_int64(0), C.c_int32(0)
        rc = _lib.lib().ofp_cccnn_param_count_ex(
            C.c_int32(m.channels), C.c_int32(m.input_size), C.c_int32(L), (C.c_int32 * L)(*m.layer_sizes),
            (C.c_int32 * L)(*m.kernel_sizes), (C.c_int32 * L)(*m.strides), C.c_int32(m.padding), C.c_int32(m.dilation),
            C.c_int32(m.pool), C.c_int32(m.batch_norm), C.c_int32(m.output_size), C.c_int32(int(m.group)), C.byref(n),
            C.byref(lags))
        assert rc == 0, _lib.lib().ofp_last_error()
        with torch.no_grad():
            V = m.conv_layers(torch.zeros(1, 1, m.input_size)).shape[-1]
        assert lags.value == 2 * V - 1 == m.n_lags
        assert n.value > m.output_size * m.channels * lags.value
