"""B200 (sm_100a) implementation of the onset -> lag -> multilateration hot path of onset-fingerprinting.

The submodules mirror the reference's module names (``detection``, ``multilateration``, ``model``, ``data``,
``calibration``, ``online_cc``) so that ``from onset_fingerprinting_b200 import detection`` replaces
``from onset_fingerprinting import detection`` for the functions DESIGN.md §1 lists; everything computes in
``csrc/`` behind the C ABI of ``include/ofp.h`` (``libofp.so``).  Nothing is imported eagerly: the CUDA library is
loaded by the first call that needs it and the import of a submodule never touches a GPU.
"""
__version__ = "0.2.0"

__all__ = ["calibration", "data", "detection", "hostpipe", "model", "multilateration", "online_cc", "parallel",
           "pipeline", "posd", "spectral", "synth"]


def library_path() -> str:
    """Path of the C-ABI shared library this package drives (honours ``OFP_LIB``)."""
    from . import _lib

    return str(_lib.LIB_PATH)
