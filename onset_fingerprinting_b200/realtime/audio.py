"""PlayRec.detect_hits (reference realtime/audio.py:39-74) without the PortAudio stream.

``BlockLocator.detect_hits(block[128, 3])`` is what the reference's audio callback runs per block:
the onset detector with the realtime settings, detections sorted by sample, each fed to
``Multilaterate3D.locate`` until one yields a position.  ``StreamBatch`` runs the detector for many
concurrent streams in one launch (BASELINE config 4: 4096 streams x 3 mics x 128-sample blocks)."""
from __future__ import annotations

from collections import namedtuple

import numpy as np

from .. import detection, multilateration
from . import config

Location = namedtuple("Location", "x y radius")

# realtime/audio.py:39-52
REALTIME_DETECTOR = dict(hipass_freq=0, fast_ar=(0.3, 800), slow_ar=(8000, 8000), on_threshold=0.45,
                         off_threshold=0.45, cooldown=1323, sr=config.SR)


class DeviceRing:
    """The most recent `n_rows` audio rows on the device -- the part of loopmate's CircularArray
    (package absent from the reference tree) that Multilaterate3D.locate uses (multilateration.py:462-466):
    ``write(block)``, ``counter`` (rows written so far), ``N`` and ``ring[-k:]`` (last k rows, time order).
    Rows before the first write read as zero."""

    def __init__(self, n_rows: int, n_channels: int):
        from .. import _lib

        self.torch = _lib.require_cuda()
        self.N = int(n_rows)
        self.data = self.torch.zeros((2 * self.N, n_channels), dtype=self.torch.float32, device="cuda")
        self.counter = 0
        self.write_counter = 0
        self._pos = 0  # rows [pos, pos + N) of the doubled buffer are the ring in time order

    def write(self, block):
        torch = self.torch
        b = torch.as_tensor(block, dtype=torch.float32).to("cuda")
        n = b.shape[0]
        if n >= self.N:
            self.data[: self.N] = b[-self.N:]
            self._pos = 0
        else:
            if self._pos + self.N + n > 2 * self.N:  # slide the window back to the start
                self.data[: self.N] = self.data[self._pos:self._pos + self.N].clone()
                self._pos = 0
            self.data[self._pos + self.N:self._pos + self.N + n] = b
            self._pos += n
        self.counter += n
        self.write_counter += n

    def __getitem__(self, item):
        return self.data[self._pos:self._pos + self.N][item]


class BlockLocator:
    def __init__(self, ml_conf: dict, n_channels: int = config.N_CHANNELS, blocksize: int = config.BLOCKSIZE,
                 detector_kw: dict | None = None, ring_rows: int | None = None):
        """ring_rows: keep that many recent audio rows and let locate() refine every new sensor pair by
        cross-correlation on them, as the reference's callback does with its recording buffer
        (realtime/audio.py:69, 102)."""
        kw = dict(REALTIME_DETECTOR)
        kw.update(detector_kw or {})
        self.current_index = 0
        self.blocksize = blocksize
        self.rec_audio = DeviceRing(ring_rows, n_channels) if ring_rows else None
        self.od = detection.AmplitudeOnsetDetector(n_channels, blocksize, backtrack=False, **kw)
        self.m = multilateration.Multilaterate3D(sensor_locations=ml_conf["sensor_locations"], sr=kw["sr"],
                                                 medium=ml_conf["medium"], c=ml_conf.get("c"))

    def detect_hits(self, audio: np.ndarray):
        """realtime/audio.py:62-74.  Advances current_index by the block length like the callback does
        (realtime/audio.py:120)."""
        if self.rec_audio is not None:
            self.rec_audio.write(audio)  # the callback writes before it detects (realtime/audio.py:102-106)
        c, d, _ = self.od(audio)
        res = None
        if len(c) > 0:
            d = [self.current_index + int(x) for x in d]
            for i in np.argsort(d):
                got = self.m.locate(int(c[i]), d[i], self.rec_audio)
                if got is not None:
                    res = Location(got[0], got[1], self.m.radius)
                    break
        self.current_index += len(audio)
        return res


class StreamBatch:
    """S concurrent realtime detectors advanced by one block per call (one kernel launch)."""

    def __init__(self, n_streams: int, n_channels: int = config.N_CHANNELS, blocksize: int = config.BLOCKSIZE,
                 detector_kw: dict | None = None):
        kw = dict(REALTIME_DETECTOR)
        kw.update(detector_kw or {})
        self.det = detection.BatchedOnsetDetector(n_streams, n_channels, blocksize, **kw)

    def process(self, blocks, return_rel: bool = False):
        """blocks [S, B, C] (device tensor or numpy) -> (channels [S, C], deltas [S, C], counts [S], rel)."""
        return self.det.process_block(blocks, return_rel=return_rel)


class StreamLocatorBatch:
    """detect_hits for S concurrent streams on the device: ``StreamBatch`` (K1, one launch per block)
    followed by ``ofp_stream_locate`` (one thread per stream runs Multilaterate3D.locate's group state
    machine on the block's detections).  BASELINE config 4: 4096 streams x 3 mics x 128-sample blocks,
    two launches per block, nothing returns to the host but the located positions."""

    def __init__(self, n_streams: int, ml_conf: dict, n_channels: int = config.N_CHANNELS,
                 blocksize: int = config.BLOCKSIZE, detector_kw: dict | None = None):
        import ctypes as C

        from .. import _lib

        kw = dict(REALTIME_DETECTOR)
        kw.update(detector_kw or {})
        self.torch = torch = _lib.require_cuda()
        self.S, self.C, self.blocksize = n_streams, n_channels, blocksize
        self.det = detection.BatchedOnsetDetector(n_streams, n_channels, blocksize, **kw)
        self.m = multilateration.Multilaterate3D(sensor_locations=ml_conf["sensor_locations"], sr=kw["sr"],
                                                 medium=ml_conf["medium"], c=ml_conf.get("c"))
        sizes = [C.c_int64() for _ in range(4)]
        _lib.check(_lib.lib().ofp_stream_locate_state_bytes(C.c_int32(n_streams), *[C.byref(v) for v in sizes]))
        self._state = [torch.zeros((v.value,), dtype=torch.uint8, device="cuda") for v in sizes]
        self.current_index = 0

    def reset(self):
        for t in self._state:
            t.zero_()
        self.det.reset()
        self.current_index = 0

    def locate_detections(self, ch, delta, cnt):
        """Feed one block's detections (ofp_detect_block's output) to every stream's state machine.
        Returns (xy [S, 2] float64, found [S] int32)."""
        import ctypes as C

        from .. import _lib
        from .._lib import ptr, stream_ptr

        torch, m = self.torch, self.m
        xy = torch.empty((self.S, 2), dtype=torch.float64, device="cuda")
        found = torch.empty((self.S,), dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib().ofp_stream_locate(
            ptr(m._locs), C.c_int32(m._S), ptr(m._maps), C.c_int32(m._M), ptr(m._mx), ptr(m._mn), ptr(m._mm),
            C.c_double(m.radius), C.c_double(m.samples_per_cm), C.c_double(m.sr), C.c_double(m.c),
            C.c_int32(self.S), C.c_int32(self.C), ptr(ch), ptr(delta), ptr(cnt), C.c_int64(self.current_index),
            ptr(self._state[0]), ptr(self._state[1]), ptr(self._state[2]), ptr(self._state[3]), ptr(xy), ptr(found),
            stream_ptr()))
        return xy, found

    def detect_hits(self, blocks):
        """blocks [S, B, C] -> (xy [S, 2] cm (NaN = no hit), found [S]); advances every stream by one block."""
        ch, dl, cnt, _ = self.det.process_block(blocks, return_rel=False)
        xy, found = self.locate_detections(ch, dl, cnt)
        self.current_index += self.blocksize
        return xy, found


class RealtimeSession:
    """``StreamLocatorBatch.detect_hits`` as one replayed CUDA graph per block (csrc/realtime.cu): the native
    session owns the detector, the locate state machines, the staging and result buffers; a step is one
    asynchronous copy of the block, one ``cudaGraphLaunch`` and one stream synchronisation, and returns
    ``(xy [S, 2] float64, found [S] int32)`` as numpy arrays on the host.  Same results as
    ``StreamLocatorBatch`` block by block (tests/test_gpu_stream_locate.py)."""

    def __init__(self, n_streams: int, ml_conf: dict, n_channels: int = config.N_CHANNELS,
                 blocksize: int = config.BLOCKSIZE, detector_kw: dict | None = None, use_graph: bool = True,
                 ring_rows: int = 0):
        """ring_rows > 0: every stream keeps a ring of its last ring_rows audio rows on the device and each new
        (group, detection) pair goes through the cross-correlation refinement of ``locate(..., rec_audio)``
        (multilateration.py:457-501) -- what PlayRec's callback runs (realtime/audio.py:69, 102).
        ring_rows = 0 is ``locate(..., rec_audio=None)``."""
        import ctypes as C

        from .. import _lib
        from .._lib import ptr

        kw = dict(REALTIME_DETECTOR)
        kw.update(detector_kw or {})
        self.torch = _lib.require_cuda()
        self.S, self.C, self.blocksize = n_streams, n_channels, blocksize
        self.m = m = multilateration.Multilaterate3D(sensor_locations=ml_conf["sensor_locations"], sr=kw["sr"],
                                                     medium=ml_conf["medium"], c=ml_conf.get("c"))
        self._params = detection.make_params(n_channels, blocksize, **kw)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().ofp_rt_create(C.byref(self._h), C.c_int32(n_streams), C.byref(self._params), ptr(m._locs),
                                            C.c_int32(m._S), ptr(m._maps), C.c_int32(m._M), ptr(m._mx), ptr(m._mn),
                                            ptr(m._mm), C.c_double(m.radius), C.c_double(m.samples_per_cm),
                                            C.c_double(m.sr), C.c_double(m.c), C.c_int32(int(use_graph)),
                                            C.c_int32(int(ring_rows))))
        self.ring_rows = int(ring_rows)
        self.xy = np.empty((n_streams, 2), np.float64)
        self.found = np.empty((n_streams,), np.int32)
        self.current_index = 0

    def reset(self):
        from .. import _lib

        _lib.check(_lib.lib().ofp_rt_reset(self._h))
        self.current_index = 0

    def detect_hits(self, blocks):
        """blocks [S, B, C] float32: a device tensor (may be a window into longer per-stream buffers) or a
        C-contiguous numpy array / pinned host tensor."""
        import ctypes as C

        from .. import _lib

        if isinstance(blocks, np.ndarray):
            assert blocks.dtype == np.float32 and blocks.flags.c_contiguous and blocks.shape == (self.S, self.blocksize, self.C)
            p, on_host, stride = blocks.ctypes.data_as(C.c_void_p), 1, 0
        else:
            assert blocks.dtype == self.torch.float32 and tuple(blocks.shape) == (self.S, self.blocksize, self.C)
            assert blocks.stride(2) == 1 and blocks.stride(1) == self.C
            p, on_host, stride = C.c_void_p(blocks.data_ptr()), int(not blocks.is_cuda), blocks.stride(0)
            if blocks.is_cuda:  # the session reads on its own stream: order it after the producer of `blocks`
                _lib.check(_lib.lib().ofp_rt_wait_stream(self._h, _lib.stream_ptr()))
        _lib.check(_lib.lib().ofp_rt_step(self._h, p, C.c_int32(on_host), C.c_int64(stride),
                                          self.xy.ctypes.data_as(C.c_void_p), self.found.ctypes.data_as(C.c_void_p)))
        self.current_index += self.blocksize
        return self.xy, self.found

    def __del__(self):
        try:
            from .. import _lib

            if self._h:
                _lib.lib().ofp_rt_destroy(self._h)
                self._h = None
        except Exception:
            pass
