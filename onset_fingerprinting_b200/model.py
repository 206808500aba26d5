"""Onset-window network inference on the GPU (K6, csrc/cnn_infer.cu).

Mirrors ``model.CNN`` of the reference (model.py:52-120) for inference: the constructor takes the same
arguments and builds the same ``conv_layers`` / ``fc`` parameter containers, so a checkpoint of the
reference model loads with ``load_state_dict`` unchanged; ``forward`` runs the fused CUDA kernel
(Conv1d + activation stack -> flatten -> Linear, one warp per window) instead of cuDNN/cuBLAS calls.
The constructor options the reference leaves off by default -- ``batch_norm`` (BatchNorm1d behind every activation,
eval mode), ``pool`` (MaxPool1d(2, 2)), ``dilation``, ``groups`` (model.py:62-66, 91-108) -- run on the generic
kernel; the reference defaults take the tensor-core kernel.  Training (LightningModule hooks, optimisers, plots;
model.py:122-165) is out of scope.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib
from ._lib import check, ptr, stream_ptr

_ACT = {nn.SiLU: 0, nn.ReLU: 1, nn.Tanh: 2, nn.Identity: 3}


def paired_xcorr(x: torch.Tensor, C: int, K: int) -> torch.Tensor:
    """model.paired_xcorr (model.py:12-45): full cross-correlation of every adjacent channel pair (1&2, 2&3, ...) of
    each feature map, averaged over the K maps.  x [B, C*K, V] -> [B, C-1, 2V-1].  The grouped ``F.conv1d`` of the
    reference is one np.correlate(a, b, "full") per (batch, pair, map) row: those rows go through
    ofp_correlate_full (csrc/onset_tools.cu, double accumulation) in one launch; the mean over K is a reshape."""
    from .multilateration import correlate_full

    B, CK, V = x.shape
    assert CK == C * K
    xv = x.reshape(B, C, K, V)
    a = xv[:, :-1].reshape(B * (C - 1) * K, V).float().contiguous()
    b = xv[:, 1:].reshape(B * (C - 1) * K, V).float().contiguous()
    cc = correlate_full(a.cuda(), b.cuda())  # [B*(C-1)*K, 2V-1]
    return cc.view(B, C - 1, K, 2 * V - 1).mean(dim=2)


class CNN(nn.Module):
    def __init__(self, input_size: int, output_size: int, channels: int = 3, layer_sizes: list[int] = [8, 16],
                 kernel_size: int = 3, dropout_rate: float = 0.5, loss=None, batch_norm=False, pool=False,
                 padding=1, dilation=1, groups=1, lr=1e-3, activation=nn.SiLU) -> None:
        super().__init__()
        if activation not in _ACT:
            raise NotImplementedError(f"activation {activation} (supported: {[a.__name__ for a in _ACT]})")
        self.input_size, self.output_size, self.channels = input_size, output_size, channels
        self.layer_sizes, self.kernel_size, self.padding = list(layer_sizes), kernel_size, padding
        self.dilation, self.groups, self.batch_norm, self.pool = int(dilation), int(groups), bool(batch_norm), bool(pool)
        self.act = _ACT[activation]
        self.conv_layers = nn.Sequential()  # same names as the reference: conv1, act1, (bn1,) (pool1,) conv2, ...
        cur, length = channels, input_size
        for i, size in enumerate(self.layer_sizes):
            self.conv_layers.add_module(f"conv{i + 1}", nn.Conv1d(cur, size, kernel_size, padding=padding,
                                                                 dilation=dilation, groups=groups))
            self.conv_layers.add_module(f"act{i + 1}", activation())
            length = length + 2 * padding - dilation * (kernel_size - 1)
            if batch_norm:
                self.conv_layers.add_module(f"bn{i + 1}", nn.BatchNorm1d(size))
            if pool:
                self.conv_layers.add_module(f"pool{i + 1}", nn.MaxPool1d(kernel_size=2, stride=2))
                length //= 2
            cur = size
        self.dropout = nn.Dropout(dropout_rate)
        self.fc = nn.Linear(cur * length, output_size)
        self.flat = cur * length
        self._packed = None
        self.eval()

    # ---- parameter packing (include/ofp.h: ofp_cnn_forward) ----
    def pack(self) -> torch.Tensor:
        parts = []
        for i in range(len(self.layer_sizes)):
            conv = getattr(self.conv_layers, f"conv{i + 1}")
            w = conv.weight.detach().float().cpu()  # [cout, cin / groups, ks]
            b = conv.bias.detach().float().cpu() if conv.bias is not None else torch.zeros(w.shape[0])
            cout = w.shape[0]
            if self.groups != 1:  # dense [cout, cin, ks] with zeros between the groups (0 * x adds nothing)
                cin_g, cout_g = w.shape[1], cout // self.groups
                dense = torch.zeros((cout, cin_g * self.groups, w.shape[2]))
                for g in range(self.groups):
                    dense[g * cout_g:(g + 1) * cout_g, g * cin_g:(g + 1) * cin_g] = w[g * cout_g:(g + 1) * cout_g]
                w = dense
            cp = (cout + 7) // 8 * 8
            wt = torch.zeros((w.shape[1], w.shape[2], cp))
            wt[:, :, :cout] = w.permute(1, 2, 0)
            bp = torch.zeros(cp)
            bp[:cout] = b
            parts += [wt.reshape(-1), bp]
            if self.batch_norm:  # eval mode: (y - mean) / sqrt(var + eps) * weight + bias = y * scale + shift
                bn = getattr(self.conv_layers, f"bn{i + 1}")
                scale = (bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)).cpu()
                shift = bn.bias.detach().double().cpu() - bn.running_mean.detach().double().cpu() * scale
                sp, tp = torch.zeros(cp), torch.zeros(cp)
                sp[:cout], tp[:cout] = scale.float(), shift.float()
                parts += [sp, tp]
        parts += [self.fc.weight.detach().float().cpu().reshape(-1), self.fc.bias.detach().float().cpu()]
        packed = torch.cat(parts).contiguous()
        n = C.c_int64(0)
        sizes = (C.c_int32 * len(self.layer_sizes))(*self.layer_sizes)
        check(_lib.lib().ofp_cnn_param_count_ex(C.c_int32(self.channels), C.c_int32(self.input_size),
                                                C.c_int32(len(self.layer_sizes)), sizes, C.c_int32(self.kernel_size),
                                                C.c_int32(self.padding), C.c_int32(self.dilation),
                                                C.c_int32(self.pool), C.c_int32(self.batch_norm),
                                                C.c_int32(self.output_size), C.byref(n), None))
        assert n.value == packed.numel(), (n.value, packed.numel())
        self._packed = packed.cuda()
        return self._packed

    def load_state_dict(self, *args, **kw):
        out = super().load_state_dict(*args, **kw)
        self._packed = None
        return out

    @torch.no_grad()
    def forward(self, x) -> torch.Tensor:
        """x [B, channels, input_size] float32 (device tensor or numpy) -> [B, output_size] device tensor."""
        _lib.require_cuda()
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        x = x.cuda().float()
        if x.dim() != 3 or x.shape[1] != self.channels or x.shape[2] != self.input_size:
            raise ValueError(f"expected [B, {self.channels}, {self.input_size}], got {tuple(x.shape)}")
        if x.stride(2) != 1 or x.stride(1) != self.input_size:
            x = x.contiguous()
        if self._packed is None:
            self.pack()
        out = torch.empty((x.shape[0], self.output_size), dtype=torch.float32, device="cuda")
        sizes = (C.c_int32 * len(self.layer_sizes))(*self.layer_sizes)
        check(_lib.lib().ofp_cnn_forward_ex(ptr(x), C.c_int64(x.shape[0]), C.c_int64(x.stride(0)),
                                            C.c_int32(self.channels), C.c_int32(self.input_size),
                                            C.c_int32(len(self.layer_sizes)), sizes, C.c_int32(self.kernel_size),
                                            C.c_int32(self.padding), C.c_int32(self.dilation), C.c_int32(self.pool),
                                            C.c_int32(self.batch_norm), C.c_int32(self.act), ptr(self._packed),
                                            C.c_int32(self.output_size), ptr(out), stream_ptr()))
        return out

    def call_np(self, x: np.ndarray) -> np.ndarray:
        """numpy in, numpy out (the shape FCNN.call_np has in the reference, calibration.py:463-560)."""
        return self.forward(x).cpu().numpy()


class CCCNN(nn.Module):
    """``model.CCCNN`` of the reference (model.py:443-538, every constructor option except ``batch_norm`` together with
    ``group=True``) for inference: same constructor arguments
    and parameter names (``conv_layers.conv1`` ..., ``fc``), fused forward in csrc/cnn_cc_infer.cuh -- per sensor
    channel the shared conv stack, the summed auto-correlation of its feature maps as ``F^T F`` on the tensor
    cores with the diagonal sums taken inside the accumulator fragments, softmax over the 2V-1 lags, Linear."""

    def __init__(self, input_size: int, output_size: int, channels: int = 3, layer_sizes: list[int] = [8, 16],
                 kernel_sizes=3, strides=1, dropout_rate: float = 0.5, batch_norm=False, pool=False, padding=1,
                 dilation=1, group: bool = False, activation=nn.SiLU) -> None:
        super().__init__()
        self.layer_sizes = list(layer_sizes)
        if isinstance(kernel_sizes, int):  # model.py:476-479
            kernel_sizes = [kernel_sizes] * len(self.layer_sizes)
        if isinstance(strides, int):
            strides = [strides] * len(self.layer_sizes)
        self.kernel_sizes, self.strides = [int(k) for k in kernel_sizes], [int(t) for t in strides]
        if batch_norm and group:
            # GroupNorm(1, K * channels) of the grouped stack takes its statistics over all sensor channels of a window;
            # the kernel runs the channels one after the other
            raise NotImplementedError("batch_norm together with group=True")
        if activation not in _ACT:
            raise NotImplementedError(f"activation {activation}")
        self.input_size, self.output_size, self.channels = input_size, output_size, channels
        self.padding, self.dilation, self.batch_norm, self.pool = padding, int(dilation), bool(batch_norm), bool(pool)
        # the reference defaults (one kernel size, stride 1, dilation 1, no norm / pool) take the specialised kernels
        self.plain = (len(set(self.kernel_sizes)) == 1 and set(self.strides) == {1} and self.dilation == 1 and
                      not self.batch_norm and not self.pool and self.kernel_sizes[0] in (1, 3, 5, 7))
        self.kernel_size = self.kernel_sizes[0]
        self.act, self.group = _ACT[activation], bool(group)
        self.conv_layers = nn.Sequential()
        g = channels if group else 1  # model.py:470-484: in / out channels times `channels`, groups = channels
        cur, length = g, input_size
        for i, (size, ks, st) in enumerate(zip(self.layer_sizes, self.kernel_sizes, self.strides)):
            self.conv_layers.add_module(f"conv{i + 1}", nn.Conv1d(cur, size * g, ks, padding=padding, dilation=dilation,
                                                                 stride=st, groups=g))
            self.conv_layers.add_module(f"act{i + 1}", activation())
            length = (length + 2 * padding - dilation * (ks - 1) - 1) // st + 1
            if batch_norm:  # model.py:494-498: the `batch_norm` option builds a GroupNorm with one group
                self.conv_layers.add_module(f"bn{i + 1}", nn.GroupNorm(1, size * g))
            if pool:
                self.conv_layers.add_module(f"pool{i + 1}", nn.MaxPool1d(kernel_size=2, stride=2))
                length //= 2
            cur = size * g
        self.dropout = nn.Dropout(dropout_rate)
        self.n_lags = 2 * length - 1
        self.fc = nn.Linear(channels * self.n_lags, output_size)
        self._packed = None
        self.eval()

    def pack(self) -> torch.Tensor:
        parts = []
        for c in range(self.channels if self.group else 1):  # group: one block per sensor channel, channel 0 first
            for i, size in enumerate(self.layer_sizes):
                conv = getattr(self.conv_layers, f"conv{i + 1}")
                w = conv.weight.detach().float().cpu()[c * size:(c + 1) * size]  # [size, c_in per group, k]
                b = conv.bias.detach().float().cpu()[c * size:(c + 1) * size]
                cp = (size + 7) // 8 * 8
                wt = torch.zeros((w.shape[1], w.shape[2], cp))
                wt[:, :, :size] = w.permute(1, 2, 0)
                bp = torch.zeros(cp)
                bp[:size] = b
                parts += [wt.reshape(-1), bp]
                if self.batch_norm:
                    gn = getattr(self.conv_layers, f"bn{i + 1}")
                    gp, tp = torch.zeros(cp), torch.zeros(cp)
                    gp[:size], tp[:size] = gn.weight.detach().float().cpu(), gn.bias.detach().float().cpu()
                    parts += [gp, tp]
        parts += [self.fc.weight.detach().float().cpu().reshape(-1), self.fc.bias.detach().float().cpu()]
        packed = torch.cat(parts).contiguous()
        n = C.c_int64(0)
        sizes = (C.c_int32 * len(self.layer_sizes))(*self.layer_sizes)
        if self.plain:
            check(_lib.lib().ofp_cccnn_param_count(C.c_int32(self.channels), C.c_int32(self.input_size),
                                                   C.c_int32(len(self.layer_sizes)), sizes, C.c_int32(self.kernel_size),
                                                   C.c_int32(self.padding), C.c_int32(self.output_size),
                                                   C.c_int32(int(self.group)), C.byref(n), None))
        else:
            check(_lib.lib().ofp_cccnn_param_count_ex(
                C.c_int32(self.channels), C.c_int32(self.input_size), C.c_int32(len(self.layer_sizes)), sizes,
                (C.c_int32 * len(self.layer_sizes))(*self.kernel_sizes), (C.c_int32 * len(self.layer_sizes))(*self.strides),
                C.c_int32(self.padding), C.c_int32(self.dilation), C.c_int32(self.pool), C.c_int32(self.batch_norm),
                C.c_int32(self.output_size), C.c_int32(int(self.group)), C.byref(n), None))
        assert n.value == packed.numel(), (n.value, packed.numel())
        self._packed = packed.cuda()
        return self._packed

    def load_state_dict(self, *args, **kw):
        out = super().load_state_dict(*args, **kw)
        self._packed = None
        return out

    @torch.no_grad()
    def forward(self, x) -> torch.Tensor:
        """x [B, channels, input_size] float32 -> [B, output_size] device tensor."""
        _lib.require_cuda()
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        x = x.cuda().float()
        if x.dim() != 3 or x.shape[1] != self.channels or x.shape[2] != self.input_size:
            raise ValueError(f"expected [B, {self.channels}, {self.input_size}], got {tuple(x.shape)}")
        if x.stride(2) != 1 or x.stride(1) != self.input_size:
            x = x.contiguous()
        if self._packed is None:
            self.pack()
        out = torch.empty((x.shape[0], self.output_size), dtype=torch.float32, device="cuda")
        sizes = (C.c_int32 * len(self.layer_sizes))(*self.layer_sizes)
        if self.plain:
            check(_lib.lib().ofp_cccnn_forward(ptr(x), C.c_int64(x.shape[0]), C.c_int64(x.stride(0)),
                                               C.c_int32(self.channels), C.c_int32(self.input_size),
                                               C.c_int32(len(self.layer_sizes)), sizes, C.c_int32(self.kernel_size),
                                               C.c_int32(self.padding), C.c_int32(self.act), C.c_int32(int(self.group)),
                                               ptr(self._packed), C.c_int32(self.output_size), ptr(out), stream_ptr()))
        else:
            check(_lib.lib().ofp_cccnn_forward_ex(
                ptr(x), C.c_int64(x.shape[0]), C.c_int64(x.stride(0)), C.c_int32(self.channels),
                C.c_int32(self.input_size), C.c_int32(len(self.layer_sizes)), sizes,
                (C.c_int32 * len(self.layer_sizes))(*self.kernel_sizes), (C.c_int32 * len(self.layer_sizes))(*self.strides),
                C.c_int32(self.padding), C.c_int32(self.dilation), C.c_int32(self.pool), C.c_int32(self.batch_norm),
                C.c_int32(self.act), C.c_int32(int(self.group)), ptr(self._packed), C.c_int32(self.output_size), ptr(out),
                stream_ptr()))
        return out
