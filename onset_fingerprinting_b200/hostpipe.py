"""Chunked host -> device -> host pipeline for the batched entry points whose inputs live in (pinned) host
memory: chunk i+1's upload runs on a second stream while chunk i computes, results return asynchronously into
pinned buffers.  The detector has its own native version of this (ofp_detect_offline_host, time segments);
this is the generic one for per-item kernels (lag refinement over sections, window networks)."""
from __future__ import annotations

from typing import Callable, Sequence

from . import _lib


_STREAMS: dict = {}


def run_chunked(host_inputs: Sequence, fn: Callable, chunk: int, outs: Sequence | None = None):
    """host_inputs: tensors with a common leading dimension n (pinned for asynchronous copies);
    fn(*device_chunks) -> tuple of device tensors with leading dimension = chunk length.
    Returns a tuple of pinned host tensors with leading dimension n; pass them back as `outs` to reuse them
    (pinned allocations cost milliseconds)."""
    torch = _lib.require_cuda()
    n = host_inputs[0].shape[0]
    # the two side streams are kept per device: fresh streams per call come with fresh allocator pools, and every
    # intermediate of fn then falls through to cudaMalloc (single calls 4x slower than the rest were measured)
    dev_idx = torch.cuda.current_device()
    if dev_idx not in _STREAMS:
        _STREAMS[dev_idx] = [torch.cuda.Stream(), torch.cuda.Stream()]
    streams = _STREAMS[dev_idx]
    cur = torch.cuda.current_stream()
    outs = list(outs) if outs is not None else None
    for s in streams:
        s.wait_stream(cur)
    # two device staging sets, one per stream, allocated once: alternating fresh allocations on two streams
    # make the caching allocator fall back to cudaMalloc / cudaFree (hundreds of ms per call)
    stage = [[torch.empty((min(chunk, n),) + tuple(h.shape[1:]), dtype=h.dtype, device="cuda") for h in host_inputs]
             for _ in streams]
    for i, lo in enumerate(range(0, n, chunk)):
        st = streams[i & 1]
        m = min(chunk, n - lo)
        with torch.cuda.stream(st):
            dev = []
            for buf, h in zip(stage[i & 1], host_inputs):
                buf[:m].copy_(h[lo:lo + m], non_blocking=True)
                dev.append(buf[:m])
            res = fn(*dev)
            if not isinstance(res, (tuple, list)):
                res = (res,)
            if outs is None:
                outs = [torch.empty((n,) + tuple(r.shape[1:]), dtype=r.dtype, pin_memory=True) for r in res]
            for o, r in zip(outs, res):
                o[lo:lo + m].copy_(r, non_blocking=True)
                r.record_stream(st)
    for s in streams:
        cur.wait_stream(s)
    cur.synchronize()
    return tuple(outs) if outs is not None else ()
