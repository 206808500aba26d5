"""Multi-GPU plumbing: one process per GPU, recordings / hits sharded by rank, no data-path
collective; the only exchange is the gather of fixed-size per-hit records at the end
(SURVEY.md section 8e).  The reference has no distributed code at all; this is the B200-box
equivalent of running its notebooks on 8 machines and concatenating the result tables."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index: int) -> dict:
    """Pin the calling process to the host cores local to GPU `device_index` (the PCIe root's
    ``local_cpulist``), so that pinned staging buffers allocated afterwards are first-touched on the GPU's own
    NUMA node and the upload thread runs there: with 8 ranks feeding 8 GPUs out of one node the host side, not
    PCIe, bounds the end-to-end path (round-1 SCALE: 184 GB/s aggregate for 8 x 55 GB/s links).  Returns what was
    found and done; a box that exposes one node (or hides sysfs) is reported as such, not treated as an error."""
    info = {"device": device_index, "bound": False}
    try:
        pr = torch.cuda.get_device_properties(device_index)  # torch >= 2.4: integer domain / bus / device ids
        bdf = f"{int(pr.pci_domain_id):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
    except Exception:
        try:
            import pynvml

            pynvml.nvmlInit()
            bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
            bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
        except Exception as e:  # no way to learn the GPU's PCI address
            info["reason"] = f"pci address unavailable: {e}"
            return info
    bdf = str(bdf).lower()
    if len(bdf.split(":")[0]) == 8:  # nvml prints a 32-bit domain, sysfs a 16-bit one
        bdf = bdf[4:]
    base = f"/sys/bus/pci/devices/{bdf}"
    try:
        node = int(open(f"{base}/numa_node").read())
        local = _parse_cpulist(open(f"{base}/local_cpulist").read())
    except Exception as e:
        info["reason"] = f"sysfs: {e}"
        return info
    allowed = os.sched_getaffinity(0)
    info.update(pci=bdf, numa_node=node, local_cpus=len(local), allowed_cpus=len(allowed))
    use = local & allowed
    if node < 0 or not use:
        info["reason"] = "no NUMA locality exposed" if node < 0 else "the GPU's local cores are outside this process's cpuset"
        return info
    if use == allowed:
        info["reason"] = "already confined to the GPU's node"
        info["bound"] = True
        return info
    os.sched_setaffinity(0, use)
    info["bound"] = True
    return info


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank`; sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_records(rec, fixed, lags, xy, fix_status, loc_status, rec_offset: int = 0) -> torch.Tensor:
    """[H, 2C + 4] int64 records: global recording id, fix status, location status, C onsets, C lags,
    then x and y bit-cast to int64."""
    H, C = fixed.shape
    out = torch.empty((H, 2 * C + 5), dtype=torch.int64, device=fixed.device)
    out[:, 0] = rec.long() + rec_offset
    out[:, 1] = fix_status.long()
    out[:, 2] = loc_status.long()
    out[:, 3:3 + C] = fixed.long()
    out[:, 3 + C:3 + 2 * C] = lags.long()
    out[:, 3 + 2 * C:] = xy.contiguous().view(torch.int64)
    return out


def unpack_records(records: torch.Tensor, n_channels: int) -> dict:
    C = n_channels
    return {
        "rec": records[:, 0], "fix_status": records[:, 1], "loc_status": records[:, 2],
        "fixed": records[:, 3:3 + C], "lags": records[:, 3 + C:3 + 2 * C],
        "xy": records[:, 3 + 2 * C:].contiguous().view(torch.float64),
    }


def gather_records(records: torch.Tensor, group=None) -> torch.Tensor:
    """All ranks receive every rank's records, concatenated in rank order.  One all_gather of the
    counts, one all_gather of the padded record blocks (NCCL over NVLink on the box, gloo in tests)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return records
    world = dist.get_world_size(group)
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=records.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    width = records.shape[1]
    padded = torch.zeros((max(max(counts), 1), width), dtype=records.dtype, device=records.device)
    padded[: records.shape[0]] = records
    blocks = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(blocks, padded, group=group)
    return torch.cat([b[:c] for b, c in zip(blocks, counts)], 0)


def gather_records_padded(records: torch.Tensor, capacity: int, group=None, work: dict | None = None):
    """Sync-free variant for a hot loop: every rank contributes a block padded to a fixed `capacity` rows, so
    neither the counts nor the payload need a host round trip.  Returns (blocks [world, capacity, width],
    counts [world] int64, both on the device; rows past a rank's count are unspecified); `compact_gathered`
    turns them into the concatenated table when a consumer needs it.  records.shape[0] must not exceed
    capacity (checked by the caller, who knows H).  `work` keeps the send / receive buffers between calls."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    width = records.shape[1]
    work = {} if work is None else work
    key = (capacity, width, records.dtype, records.device, world)
    if work.get("key") != key:
        work.clear()
        work["key"] = key
        work["send"] = torch.empty((capacity, width), dtype=records.dtype, device=records.device)
        work["blocks"] = torch.empty((world, capacity, width), dtype=records.dtype, device=records.device)
        work["counts"] = torch.empty((world,), dtype=torch.int64, device=records.device)
        work["n"] = torch.empty((1,), dtype=torch.int64, device=records.device)
    send, blocks, counts, n = work["send"], work["blocks"], work["counts"], work["n"]
    send[: records.shape[0]] = records
    n.fill_(records.shape[0])
    if world == 1:
        blocks[0] = send
        counts.copy_(n)
        return blocks, counts
    try:
        dist.all_gather_into_tensor(blocks.view(world * capacity, width), send, group=group)
        dist.all_gather_into_tensor(counts, n, group=group)
    except (RuntimeError, NotImplementedError):  # a backend without the flat collective
        dist.all_gather(list(blocks.unbind(0)), send, group=group)
        dist.all_gather(list(counts.split(1)), n, group=group)
    return blocks, counts


def compact_gathered(blocks: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """Concatenate the valid rows of `gather_records_padded`'s result in rank order (host sync on counts)."""
    return torch.cat([blocks[r, : int(c)] for r, c in enumerate(counts.tolist())], 0)
