"""calibration.FCNN of the reference (calibration.py:463-560) for inference on the GPU.

The reference can hand ``Multilaterate3D`` a small fully connected network that maps the two observed
lags straight to a position (multilateration.py:350, 553-557).  ``FCNN`` here takes the same constructor
arguments and builds the same ``network`` Sequential (so reference checkpoints load unchanged);
``forward`` / ``call_np`` run csrc/multilaterate.cu:k5_fcnn (one thread per row, parameters in shared
memory, BatchNorm1d as its inference affine).  Training helpers (l2_loss, optimize_positions, ...) are
out of scope.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib
from ._lib import check, ptr, stream_ptr

_ACT = {nn.ReLU: 0, nn.Tanh: 1, nn.Sigmoid: 2, nn.SiLU: 3, nn.Identity: 4}


class FCNN(nn.Module):
    def __init__(self, input_size: int, output_size: int, hidden_layers: list[int] = [10, 10, 10],
                 activation=nn.ReLU, dropout: float = 0.0, batch_norm: bool = True, l2_reg: float = 0.0,
                 eye_init=False, eye_noise_floor=0.01, bias=True) -> None:
        super().__init__()
        if activation not in _ACT:
            raise NotImplementedError(f"activation {activation}")
        self.l2_reg = l2_reg
        self.act = _ACT[activation]
        self.widths = [input_size] + list(hidden_layers) + [output_size]
        if max(self.widths) > 32 or len(self.widths) - 1 > 8:
            raise NotImplementedError("k5_fcnn: up to 8 layers of width <= 32")
        layers = []
        sizes = [input_size] + list(hidden_layers)
        for i in range(len(sizes) - 1):
            layer = nn.Linear(sizes[i], sizes[i + 1], bias=bias)
            if eye_init:
                self.init_eye_weights(layer, eye_noise_floor)
            layers.append(layer)
            if batch_norm:
                layers.append(nn.BatchNorm1d(sizes[i + 1]))
            layers.append(activation())
            if dropout > 0:
                layers.append(nn.Dropout(p=dropout))
        layer = nn.Linear(sizes[-1], output_size, bias=bias)
        if eye_init:
            self.init_eye_weights(layer, eye_noise_floor)
        layers.append(layer)
        self.network = nn.Sequential(*layers)  # same module order / state-dict keys as the reference
        self._packed = None
        self.eval()

    def init_eye_weights(self, layer, noise_floor=0.001):
        """calibration.py:542-548."""
        perturbation = torch.randn(layer.out_features, layer.in_features) * noise_floor
        layer.weight.data = torch.eye(layer.out_features, layer.in_features) + perturbation

    def load_state_dict(self, *args, **kw):
        out = super().load_state_dict(*args, **kw)
        self._packed = None
        return out

    def pack(self) -> torch.Tensor:
        """Per layer W [out][in], b [out], scale [out], shift [out] (include/ofp.h: ofp_fcnn_forward)."""
        parts = []
        mods = list(self.network)
        i = 0
        while i < len(mods):
            lin = mods[i]
            assert isinstance(lin, nn.Linear)
            w = lin.weight.detach().double().cpu()
            b = lin.bias.detach().double().cpu() if lin.bias is not None else torch.zeros(w.shape[0], dtype=torch.float64)
            scale, shift = torch.ones(w.shape[0], dtype=torch.float64), torch.zeros(w.shape[0], dtype=torch.float64)
            i += 1
            if i < len(mods) and isinstance(mods[i], nn.BatchNorm1d):
                bn = mods[i]
                scale = bn.weight.detach().double().cpu() / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
                shift = bn.bias.detach().double().cpu() - bn.running_mean.detach().double().cpu() * scale
                i += 1
            while i < len(mods) and not isinstance(mods[i], nn.Linear):
                i += 1  # activation / dropout
            parts += [w.reshape(-1), b, scale, shift]
        self._packed = torch.cat(parts).float().contiguous().cuda()
        return self._packed

    @torch.no_grad()
    def forward_device(self, x, status=None, out_scale: float = 1.0, out_f64=None):
        """x [n, input_size] float32 device tensor -> [n, output_size] float32 (and float64 into out_f64,
        only for rows whose status is 0)."""
        _lib.require_cuda()
        x = x.to(device="cuda", dtype=torch.float32).contiguous()
        if self._packed is None:
            self.pack()
        n = x.shape[0]
        out = torch.zeros((n, self.widths[-1]), dtype=torch.float32, device="cuda")
        widths = (C.c_int32 * len(self.widths))(*self.widths)
        check(_lib.lib().ofp_fcnn_forward(ptr(x), C.c_int64(n), C.c_int32(len(self.widths) - 1), widths,
                                          C.c_int32(self.act), ptr(self._packed), ptr(status),
                                          C.c_float(out_scale), ptr(out), ptr(out_f64), stream_ptr()))
        return out

    def forward(self, x) -> torch.Tensor:
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return self.forward_device(x)

    def call_np(self, lags) -> np.ndarray:
        """calibration.py:550-560: one pair of lags in, one position out (numpy)."""
        return self.forward(torch.tensor([lags], dtype=torch.float32)).cpu().numpy()[0]
