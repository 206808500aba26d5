"""Drop-in for the reference's ``online_cc`` extension module (c/cross_corr.c) on libofp.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr, stream_ptr


class CrossCorrelation:
    """online_cc.CrossCorrelation(n, block_size) (c/cross_corr.c:195-291).  ``update(a, b)`` returns the
    2n-1-lag cross-correlation of the last n samples; like the reference it hands back the SAME
    ndarray object on every call (c/cross_corr.c:270-272).  ``n_pairs > 1`` batches independent
    stream pairs ([P, block] inputs, [P, 2n-1] output)."""

    def __init__(self, n: int, block_size: int, n_pairs: int = 1):
        self.torch = _lib.require_cuda()
        self.n, self.block_size, self.n_pairs = n, block_size, n_pairs
        self._h = C.c_void_p()
        check(_lib.lib().ofp_ccstream_create(C.byref(self._h), C.c_int32(n_pairs), C.c_int32(n), C.c_int32(block_size)))
        self._out_dev = self.torch.empty((n_pairs, 2 * n - 1), dtype=self.torch.float32, device="cuda")
        self._out = np.zeros(2 * n - 1 if n_pairs == 1 else (n_pairs, 2 * n - 1), np.float32)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().ofp_ccstream_destroy(h)
            except Exception:
                pass

    def update_dev(self, a, b):
        """Device tensors [P, block] -> device tensor [P, 2n-1] (no host copy)."""
        check(_lib.lib().ofp_ccstream_update(self._h, ptr(a), ptr(b), ptr(self._out_dev), stream_ptr()))
        return self._out_dev

    def update(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        torch = self.torch
        ad = torch.from_numpy(np.ascontiguousarray(a, np.float32).reshape(self.n_pairs, self.block_size)).cuda()
        bd = torch.from_numpy(np.ascontiguousarray(b, np.float32).reshape(self.n_pairs, self.block_size)).cuda()
        out = self.update_dev(ad, bd).cpu().numpy()
        self._out[...] = out[0] if self.n_pairs == 1 else out
        return self._out
