"""Block API of the reference's realtime package (realtime/audio.py, realtime/config.py) on libofp.so.
Only the hot-path entry (PlayRec.detect_hits) is mirrored; audio I/O, looper IPC and FX mapping are
out of scope."""
