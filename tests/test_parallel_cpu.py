"""Multi-process path on CPU (gloo, world_size 2): sharding and the gather of per-hit records."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from onset_fingerprinting_b200 import parallel


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 10000, 10001):
        for world in (1, 2, 3, 8):
            got = [parallel.shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C, H = 3, 5 + 2 * rank
    g = torch.Generator().manual_seed(rank)
    rec = torch.arange(H, dtype=torch.int32)
    fixed = torch.randint(0, 1000, (H, C), generator=g, dtype=torch.int32)
    lags = torch.randint(-50, 50, (H, C), generator=g, dtype=torch.int32)
    xy = torch.randn((H, 2), generator=g, dtype=torch.float64)
    xy[0, 0] = float("nan")
    st = torch.zeros(H, dtype=torch.int32)
    mine = parallel.pack_records(rec, fixed, lags, xy, st, st + rank, rec_offset=100 * rank)
    everything = parallel.gather_records(mine)
    d = parallel.unpack_records(everything, C)
    ok = everything.shape[0] == sum(5 + 2 * r for r in range(world))
    lo = sum(5 + 2 * r for r in range(rank))
    ok &= torch.equal(d["fixed"][lo:lo + H], fixed.long())
    ok &= torch.equal(torch.nan_to_num(d["xy"][lo:lo + H]), torch.nan_to_num(xy))
    ok &= bool(torch.isnan(d["xy"][lo, 0]))
    ok &= int(d["rec"][lo]) == 100 * rank and int(d["loc_status"][lo]) == rank
    # the sync-free fixed-capacity form used inside the bench step gives the same table
    blocks, counts = parallel.gather_records_padded(mine, capacity=16)
    ok &= counts.tolist() == [5 + 2 * r for r in range(world)]
    ok &= torch.equal(parallel.compact_gathered(blocks, counts), everything)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gather_records_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29611, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_bind_to_gpu_numa_reports_instead_of_failing():
    """Without a GPU (or with a box that hides NUMA locality) the helper says why it did nothing."""
    from onset_fingerprinting_b200 import parallel

    info = parallel.bind_to_gpu_numa(0)
    assert info["device"] == 0 and isinstance(info["bound"], bool)
    assert info["bound"] or "reason" in info
    assert parallel._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
