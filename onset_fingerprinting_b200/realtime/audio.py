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


class BlockLocator:
    def __init__(self, ml_conf: dict, n_channels: int = config.N_CHANNELS, blocksize: int = config.BLOCKSIZE,
                 detector_kw: dict | None = None):
        kw = dict(REALTIME_DETECTOR)
        kw.update(detector_kw or {})
        self.current_index = 0
        self.blocksize = blocksize
        self.od = detection.AmplitudeOnsetDetector(n_channels, blocksize, backtrack=False, **kw)
        self.m = multilateration.Multilaterate3D(sensor_locations=ml_conf["sensor_locations"], sr=kw["sr"],
                                                 medium=ml_conf["medium"], c=ml_conf.get("c"))

    def detect_hits(self, audio: np.ndarray):
        """realtime/audio.py:62-74.  Advances current_index by the block length like the callback does
        (realtime/audio.py:120)."""
        c, d, _ = self.od(audio)
        res = None
        if len(c) > 0:
            d = [self.current_index + int(x) for x in d]
            for i in np.argsort(d):
                got = self.m.locate(int(c[i]), d[i])
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
