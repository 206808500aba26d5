"""Import the UNMODIFIED Python reference from /root/reference (build container only).

TEST INFRASTRUCTURE.  Used by oracle/make_golden.py (fixture generation) and by
tests that are skipped when /root/reference is absent (i.e. on the GPU box).
Nothing in the product path imports this file.

The reference cannot be imported as shipped:
  * detection.py:5,7 imports ``librosa`` and ``loopmate.circular_array`` -- neither is
    installed; they are only touched by detect_onsets_spectral / backtrack=True /
    Multilaterate3D.locate(rec_audio=...).  Empty stub modules are registered, with a
    small time-ordered ring buffer standing in for loopmate's CircularArray.
  * detection.py:517-519,559-561 ``ctypes.CDLL(Path(__file__).parent/"envelope_follower.so")``
    -- /root/reference is read-only, so the DLL is compiled by oracle/Makefile
    (``make ref``) into oracle/_ref/ and ctypes.CDLL is redirected to it.
"""
from __future__ import annotations

import ctypes
import os
import sys
import types
from pathlib import Path

import numpy as np

REF_ROOT = Path(os.environ.get("OFP_REFERENCE_ROOT", "/root/reference"))
HERE = Path(__file__).resolve().parent
REF_SO = HERE / "_ref" / "envelope_follower.so"


def available() -> bool:
    return (REF_ROOT / "onset_fingerprinting" / "detection.py").exists() and REF_SO.exists()


class RingStub:
    """Stand-in for loopmate.circular_array.CircularArray (package absent).

    Semantics needed by the reference (detection.py:719-721,756,802-803;
    multilateration.py:462-466): ``write(block)`` appends rows, ``N`` is the capacity,
    ``counter`` the total number of rows written, ``ring[-k:]`` the last k rows in time
    order.
    """

    def __init__(self, data, *_, **__):
        self.data = np.zeros_like(data)  # loopmate's ring is absent; rows before the start read 0
        self.N = data.shape[0]
        self.counter = 0
        self.write_counter = 0

    def write(self, block):
        block = np.asarray(block)
        n = len(block)
        self.data = np.concatenate([self.data, block.astype(self.data.dtype)], 0)[-self.N:]
        self.counter += n
        self.write_counter += n

    def __getitem__(self, item):
        return self.data[item]


_loaded = None


def load_reference():
    """Returns (detection, multilateration) modules of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(
            "reference not available (needs /root/reference and `make -C oracle ref`)"
        )
    for name in ("loopmate", "loopmate.circular_array"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if "librosa" not in sys.modules:  # absent: a scipy-based stand-in for the three calls detect_onsets_spectral makes
        from . import librosa_standin

        sys.modules["librosa"] = librosa_standin.module()
        sys.modules["librosa.util"] = sys.modules["librosa"].util
        sys.modules["librosa.filters"] = sys.modules["librosa"].filters
    sys.modules["loopmate.circular_array"].CircularArray = RingStub
    sys.modules["loopmate"].circular_array = sys.modules["loopmate.circular_array"]

    real_cdll = ctypes.CDLL

    class _RedirectCDLL(real_cdll):  # type: ignore[misc,valid-type]
        def __init__(self, name, *a, **k):
            if name is not None and str(name).endswith("envelope_follower.so"):
                name = str(REF_SO)
            super().__init__(str(name) if name is not None else None, *a, **k)

    ctypes.CDLL = _RedirectCDLL
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    import onset_fingerprinting.detection as det  # noqa: E402
    import onset_fingerprinting.multilateration as ml  # noqa: E402

    _loaded = (det, ml)
    return _loaded


def load_online_cc():
    """The reference's CPython extension compiled by `make ref` (c/cross_corr.c)."""
    p = str(HERE / "_ref")
    if p not in sys.path:
        sys.path.insert(0, p)
    import online_cc  # noqa: E402

    return online_cc


class _AnyModule(types.ModuleType):
    """Stub whose every attribute is a do-nothing callable class (audiomentations, soundfile ... are
    absent; data.py builds a module-level AUGMENTATIONS list at import time)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        class _Dummy:
            def __init__(self, *a, **k):
                pass

            def __call__(self, *a, **k):
                raise RuntimeError(f"{self.__class__.__name__} is a stub")

        _Dummy.__name__ = name
        return _Dummy


def load_reference_data():
    """The reference's data.py (FrameExtractor, FastFrameExtractor, batch_cc, MCPOSD) with its absent
    third-party imports stubbed; none of the stubbed packages is touched by those four."""
    load_reference()
    for name in ("audiomentations", "soundfile"):
        if name not in sys.modules:
            sys.modules[name] = _AnyModule(name)
    import onset_fingerprinting.data as data  # noqa: E402

    return data
