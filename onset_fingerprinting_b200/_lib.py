"""ctypes binding of libofp.so (the C ABI declared in include/ofp.h).

There is no CPU fallback: if the CUDA library has not been built, or a compute call is made
without a CUDA device, this raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
(nvcc -gencode arch=compute_100a,code=sm_100a).
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import os

PKG = Path(__file__).resolve().parent
# OFP_LIB selects another build of the same ABI (A/B runs of kernel variants); default: the in-tree library
LIB_PATH = Path(os.environ["OFP_LIB"]).resolve() if os.environ.get("OFP_LIB") else PKG / "libofp.so"
HEADER = PKG.parent / "include" / "ofp.h"

_lib = None


class OfpError(RuntimeError):
    pass


class DetectorParams(C.Structure):
    """ofp_detector_params (include/ofp.h)."""

    _fields_ = [
        ("n_channels", C.c_int32), ("block_size", C.c_int32), ("use_hp", C.c_int32),
        ("manual", C.c_int32), ("cooldown", C.c_int32),
        ("b", C.c_float * 5), ("a", C.c_float * 5),
        ("floor_db", C.c_float),
        ("fast_att", C.c_float), ("fast_rel", C.c_float),
        ("slow_att", C.c_float), ("slow_rel", C.c_float),
        ("on_thr", C.c_float), ("off_thr", C.c_float),
        ("alpha_min", C.c_float), ("alpha_max", C.c_float), ("minmin", C.c_float),
    ]


def declared_symbols() -> list[str]:
    """Every function include/ofp.h declares."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ofp_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise OfpError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built and there is no CPU "
                "fallback.  Run: python -c \"import __graft_entry__ as g; g.build()\""
            )
        L = C.CDLL(str(LIB_PATH))
        L.ofp_last_error.restype = C.c_char_p
        L.ofp_version.restype = C.c_char_p
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise OfpError(f"libofp error {rc}: {lib().ofp_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return C.c_void_p(None if t is None else t.data_ptr())


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise OfpError("no CUDA device: onset_fingerprinting_b200 has no CPU fallback")
    return torch
