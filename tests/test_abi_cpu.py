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
