"""Batched streaming locate (SURVEY 8f rank 1 / BASELINE config 4): one thread per stream runs
Multilaterate3D.locate's group state machine.  Golden: the unmodified reference fed the same per-block
detections stream by stream (tests/golden/stream_fsm.npz: located positions and the length of
`ongoing` after every block)."""
import numpy as np
import pytest

from onset_fingerprinting_b200 import synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ML_CONF = {"sensor_locations": synth.SENSORS_3MIC, "medium": "air", "c": None}


def test_group_state_machine_vs_reference(golden_dir):
    from oracle.make_golden import stream_fsm_events
    from onset_fingerprinting_b200.realtime.audio import StreamLocatorBatch

    g = np.load(golden_dir / "stream_fsm.npz")
    ev, cnt = stream_fsm_events()
    S, NB = cnt.shape
    sl = StreamLocatorBatch(S, ML_CONF)
    n_loc = 0
    for b in range(NB):
        ch = torch.from_numpy(ev[:, b, :, 0].astype(np.int32)).cuda()
        dl = torch.from_numpy(ev[:, b, :, 1].astype(np.int32)).cuda()
        c = torch.from_numpy(cnt[:, b].astype(np.int32)).cuda()
        xy, found = sl.locate_detections(ch, dl, c)
        sl.current_index += 128
        xy, found = xy.cpu().numpy(), found.cpu().numpy()
        want = g["res"][:, b]
        ok = np.isfinite(want[:, 0])
        assert np.array_equal(found == 1, ok), b
        assert (found >= 0).all()
        assert np.allclose(xy[ok], want[ok], rtol=1e-9, atol=1e-9)
        n_groups = sl._state[0].view(torch.int32).cpu().numpy() & 0xff  # upper bits: the group-identity counter
        assert np.array_equal(n_groups, g["n_groups"][:, b]), b
        n_loc += int(ok.sum())
    assert n_loc > 900


def test_detect_hits_batch_matches_single_stream_locator():
    """K1 + state machine for a batch of streams == BlockLocator (the single-stream mirror of
    PlayRec.detect_hits) run on each stream separately."""
    from onset_fingerprinting_b200.realtime.audio import BlockLocator, StreamLocatorBatch

    S, nblk = 5, 600
    xs, _ = synth.drum_batch(S, seconds=nblk * 128 / 96000, seed=77, first_hit=9000)
    xs = xs[:, : nblk * 128]
    sl = StreamLocatorBatch(S, ML_CONF)
    xd = torch.from_numpy(xs).cuda()
    got = {}
    for b in range(nblk):
        xy, found = sl.detect_hits(xd[:, b * 128:(b + 1) * 128])
        f = found.cpu().numpy()
        for s in np.nonzero(f == 1)[0]:
            got[(int(s), b)] = xy[s].cpu().numpy()
    assert len(got) > 10
    for s in range(S):
        bl = BlockLocator(ML_CONF)
        for b in range(nblk):
            r = bl.detect_hits(xs[s, b * 128:(b + 1) * 128])
            if r is None:
                assert (s, b) not in got
            else:
                assert np.array_equal(got[(s, b)], np.asarray([r.x, r.y]))


@pytest.mark.parametrize("use_graph", [True, False])
def test_realtime_session_graph_equals_python_path(use_graph):
    """The native session (csrc/realtime.cu: one replayed CUDA graph per block) gives, block by block, exactly
    what StreamLocatorBatch.detect_hits gives; device windows, numpy blocks and a reset in between."""
    from onset_fingerprinting_b200.realtime.audio import RealtimeSession, StreamLocatorBatch

    S, nblk = 37, 500
    xs, _ = synth.drum_batch(S, seconds=nblk * 128 / 96000, seed=78, first_hit=9000)
    xs = np.ascontiguousarray(xs[:, : nblk * 128])
    xd = torch.from_numpy(xs).cuda()
    sl = StreamLocatorBatch(S, ML_CONF)
    rs = RealtimeSession(S, ML_CONF, use_graph=use_graph)
    for attempt in range(2):
        n_loc = 0
        for b in range(nblk):
            want_xy, want_f = sl.detect_hits(xd[:, b * 128:(b + 1) * 128])
            blk = xd[:, b * 128:(b + 1) * 128] if b % 2 else np.ascontiguousarray(xs[:, b * 128:(b + 1) * 128])
            xy, f = rs.detect_hits(blk)
            want_xy, want_f = want_xy.cpu().numpy(), want_f.cpu().numpy()
            assert np.array_equal(f, want_f), (attempt, b)
            assert np.array_equal(np.isnan(xy), np.isnan(want_xy)) and np.array_equal(np.nan_to_num(xy), np.nan_to_num(want_xy))
            n_loc += int((f == 1).sum())
        assert n_loc >= S  # every stream locates at least one hit on average
        sl.reset(); rs.reset()


@pytest.mark.parametrize("use_graph", [True, False])
def test_ring_refinement_session_vs_reference(use_graph, golden_dir):
    """The realtime path as PlayRec runs it -- locate(sensor, onset, rec_audio), i.e. WITH the ring-buffer
    cross-correlation refinement of every new pair (multilateration.py:457-501) -- for 48 concurrent streams in the
    native session (one warp per stream, per-stream device ring): block by block the same hits at the same
    positions as the unmodified reference run stream by stream (tests/golden/stream_ring.npz)."""
    from oracle.make_golden import STREAM_RING, stream_ring_input
    from onset_fingerprinting_b200.realtime.audio import RealtimeSession

    g = np.load(golden_dir / "stream_ring.npz")
    S = STREAM_RING["n_streams"]
    xs = np.stack([stream_ring_input(s) for s in range(S)])
    nblk = xs.shape[1] // 128
    want = {(int(r[0]), int(r[1])): r[2:] for r in g["rows"]}
    assert len(want) == 289 and (g["raised"] < 0).all()
    rs = RealtimeSession(S, ML_CONF, use_graph=use_graph, ring_rows=STREAM_RING["ring_rows"])
    xd = torch.from_numpy(xs).cuda()
    for attempt in range(2):
        got = {}
        for b in range(nblk):
            xy, f = rs.detect_hits(xd[:, b * 128:(b + 1) * 128])
            assert (f >= 0).all(), (b, f.min())
            for s in np.nonzero(f == 1)[0]:
                got[(int(s), b)] = xy[s].copy()
        assert set(got) == set(want)
        err = max(float(np.abs(got[k] - want[k]).max()) for k in want)
        assert err <= 1e-9 * 20, err
        rs.reset()
    # the refinement changes results: without the ring the same streams locate a different set of hits
    plain = RealtimeSession(S, ML_CONF, use_graph=use_graph)
    got0 = set()
    for b in range(nblk):
        xy, f = plain.detect_hits(xd[:, b * 128:(b + 1) * 128])
        got0 |= {(int(s), b) for s in np.nonzero(f == 1)[0]}
    assert got0 != set(want)
