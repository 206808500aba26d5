#!/usr/bin/env python
"""Benchmark of the onset-fingerprinting hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path over one batch of synthetic recordings (BASELINE.json configs[1]:
10k synthetic 3-mic 96 kHz recordings batched on one B200): onset detection (K1) over every sample,
then grouping, lag refinement (K4) and multilateration (K5) of every detected hit when those stages
are enabled.  `value` = channel-samples per second with the batch resident in HBM; `e2e` = the same
metric through the C-ABI host-buffer entry point (pinned host memory in, onsets out).
Multi-GPU: one process per GPU (torchrun), recordings sharded by rank, no data-path collective except
the final gather of per-hit records; value = all ranks' units / max-over-ranks time ("weak" scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

SR = 96000
BLOCK = 128
N_CH = 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--recordings", type=int, default=10000, help="recordings per GPU")
    ap.add_argument("--seconds", type=float, default=5.0, help="length of each recording")
    ap.add_argument("--no-rel", action="store_true", help="onsets-only mode (4 B per channel-sample)")
    ap.add_argument("--e2e-recordings", type=int, default=0,
                    help="recordings of the batch pushed through the host-buffer leg (0 = all that fit in host memory)")
    ap.add_argument("--skip-hits16", action="store_true", help="batch: leave the configs[2] leg out of the line")
    ap.add_argument("--hits16-per-gpu", type=int, default=200000, help="batch: hits per GPU of the configs[2] leg")
    ap.add_argument("--parity-recordings", type=int, default=8,
                    help="recordings of the timed batch checked against the CPU oracle after the timed region")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--k1-only", action="store_true",
                    help="time only the detector kernel (A/B runs and the speed-of-light ladder, scripts/k1_ladder.sh)")
    ap.add_argument("--workload", default="batch", choices=["batch", "hits16", "realtime", "spectral", "cnn"],
                    help="batch = configs[1] (headline); hits16 = configs[2] (16-channel hit mining, K4+K5); "
                         "realtime = configs[3] (4096 concurrent 128-sample block streams)")
    ap.add_argument("--windows", type=int, default=1000000, help="cnn: onset windows per GPU")
    ap.add_argument("--network", default="cnn", choices=["cnn", "cccnn"], help="cnn workload: model.CNN or model.CCCNN")
    ap.add_argument("--hits", type=int, default=1000000, help="hits16: number of hits (over all GPUs)")
    ap.add_argument("--streams", type=int, default=4096, help="realtime: concurrent streams per GPU")
    ap.add_argument("--blocks", type=int, default=1000, help="realtime: consecutive blocks")
    ap.add_argument("--ring-rows", type=int, default=2048, help="realtime: rows of recent audio kept per stream for the "
                                                                "ring-buffer refinement leg")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index=0, period=0.1):
        self.samples, self.reasons, self.maxmhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        self.index, self.period = index, period
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.maxmhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.maxmhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.maxmhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arms: the reference's CPU implementation of the WHOLE chain on the host cores
#   detector  = the reference's compiled DLL (oracle/_ref/envelope_follower.so) driven by a numpy block loop of the
#               reference's granularity (oracle/ref_style.py; bit-identical to detection.detect_onsets_amplitude)
#   later     = find_onset_groups -> fix_onsets -> Multilaterate3D.locate restated with the reference's own numpy /
#               scipy calls per hit (oracle/ref_chain.py; checked against goldens of the unmodified reference)
# The reference's .py files themselves live under /root/reference, which does not exist on the GPU box.
# ------------------------------------------------------------------------------------------------
METRIC = "channel-samples/sec through the onset->lag->multilateration hot path (detect, group, lag-refine, locate)"
_CPU = {}


def _chain_worker(i):
    from oracle import ref_chain

    if "loc" not in _CPU:
        from onset_fingerprinting_b200 import synth

        _CPU["loc"] = ref_chain.Locator(synth.SENSORS_3MIC, sr=SR, medium="air")
    t0 = time.perf_counter()
    n_on, n_hit, n_loc = ref_chain.chain(_CPU["xs"][i], _CPU["loc"], block_size=BLOCK, sr=SR)
    return n_on, n_hit, n_loc, time.perf_counter() - t0


class CpuChain:
    """Worker pool over recordings held by the parent before the fork (no per-step pickling of audio);
    the pool and the DLL handle are created ONCE, outside any timed step."""

    def __init__(self, xs: np.ndarray, cores: int):
        import multiprocessing as mp

        from oracle import ref_style

        ref_style._dll()  # map oracle/_ref/envelope_follower.so in the parent as well (the driver lists loaded .so files)
        self.kind = ref_style.kind()
        _CPU["xs"] = xs
        self.xs, self.cores = xs, cores
        self.pool = mp.get_context("fork").Pool(cores) if cores > 1 else None
        if self.pool is not None:
            self.pool.map(_chain_worker, range(min(cores, len(xs))))  # start-up (imports, lag maps) outside the timing

    def run(self):
        t0 = time.perf_counter()
        res = self.pool.map(_chain_worker, range(len(self.xs)), chunksize=1) if self.pool else \
            [_chain_worker(i) for i in range(len(self.xs))]
        dt = time.perf_counter() - t0
        res = np.asarray(res, np.float64)
        return {"seconds": dt, "rate": self.xs.size / dt, "onsets": int(res[:, 0].sum()), "hits": int(res[:, 1].sum()),
                "located": int(res[:, 2].sum()), "hits_per_sec": float(res[:, 2].sum() / dt)}

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_kind(chain: "CpuChain"):
    return chain.kind if chain.kind == "reference" else "port"


CPU_NOTE = ("detector: the reference's compiled envelope_follower.so under a numpy block loop (oracle/ref_style.py); "
            "grouping / lag refinement / multilateration: the reference's numpy+scipy calls per hit restated in "
            "oracle/ref_chain.py (np.correlate, median_filter, fsolve)")


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the whole path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from onset_fingerprinting_b200 import synth

    cores = host_cores()
    n = int(args.seconds * SR)
    # one step = a bounded sample of the workload: `cores` recordings (one per worker), whole chain each
    n_rec = max(cores, 1)
    xs = np.stack([synth.drum_recording(args.seconds, seed=1000 + r)[0][:n] for r in range(n_rec)])
    chain = CpuChain(xs, cores)
    runs = [chain.run() for _ in range(args.warmup + args.steps)][args.warmup:]
    chain.close()
    ms = 1e3 * float(np.mean([r["seconds"] for r in runs]))
    value = xs.size / (ms / 1e3)
    sample = f"{n_rec} recordings x {args.seconds:g} s x {N_CH} ch per step ({xs.size} channel-samples), whole chain"
    h16 = None
    if args.workload in ("batch", "hits16") and not args.skip_hits16:
        h16 = hits16_reference(cores, args.cpu_seconds)
    line = {
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "channel-samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.recordings),
        "cpu_baseline": {"value": value, "unit": "channel-samples/s", "cores": cores, "kind": cpu_kind(chain),
                         "sample": sample, "note": CPU_NOTE, "onsets": runs[-1]["onsets"], "hits": runs[-1]["hits"],
                         "located": runs[-1]["located"], "localised_hits_per_sec": runs[-1]["hits_per_sec"]},
        "e2e": {"value": value, "unit": "channel-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "localised_hits_per_sec": runs[-1]["hits_per_sec"], "hits16": h16,
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, n_rec):
    return {
        "workload": f"configs[1]: {n_rec} synthetic 3-mic 96 kHz recordings x {args.seconds:g} s per GPU, "
                    f"block {BLOCK}, high-pass 2 kHz, reference defaults",
        "recordings_per_gpu": n_rec, "seconds": args.seconds, "channels": N_CH, "sr": SR, "block_size": BLOCK,
        "mode": "onsets_only" if args.no_rel else "drop_in (rel envelope written to HBM)",
        "l2": "inputs larger than L2 (no flush needed)",
        "pipelining": "stages back to back on one stream",
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from onset_fingerprinting_b200 import _lib, detection, synth

    R, N = args.recordings, int(args.seconds * SR)
    nb = N // BLOCK
    from onset_fingerprinting_b200 import pipeline

    x = synth.drum_batch_device(R, N, seed=1234, rec_offset=rank * R)
    hp = pipeline.HotPath(R, N_CH, synth.SENSORS_3MIC, medium="air", sr=SR, block_size=BLOCK)
    det = hp.det
    warm_n = int(0.5 * SR)
    launches = 0
    LAUNCHES_PER_STEP = 6  # k1_reset, k1_detect, k3_group, k3_compact, k4_fix, k5_locate (+ torch cumsum/clamp)
    last = {}

    def gather(hb):
        """N > 1: every rank receives all per-hit records (the path's only collective)."""
        if dist is not None:
            from onset_fingerprinting_b200 import parallel

            recs = parallel.pack_records(hb.rec, hb.fixed, hb.lags, hb.xy, hb.fix_status, hb.loc_status,
                                         rec_offset=rank * R)
            # fixed-capacity blocks: no host round trip for counts inside the step (H is already on the host)
            cap_h = last.setdefault("gather_capacity", int(recs.shape[0] * 1.05) + 1024)
            if recs.shape[0] <= cap_h:
                last["all_records"] = parallel.gather_records_padded(recs, cap_h, work=last.setdefault("gather_work", {}))
            else:
                last["all_records"] = parallel.gather_records(recs)

    def step():
        nonlocal launches
        last["hits"] = hp.run(x, return_rel=not args.no_rel, warm_n=warm_n)
        gather(last["hits"])
        launches += LAUNCHES_PER_STEP

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        t_wall = time.perf_counter()
        start.record()
        for i in range(args.steps):  # stages back to back on one stream
            last["hits"] = hp.run(x, return_rel=not args.no_rel, warm_n=warm_n, k1_events=ev[i])
            gather(last["hits"])
            launches += LAUNCHES_PER_STEP
        end.record()
        barrier()
        t_wall = time.perf_counter() - t_wall
    total_ms = start.elapsed_time(end)
    k1_ms = float(np.mean([k0.elapsed_time(k1) for k0, k1 in ev]))
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    units_per_step = R * nb * BLOCK * N_CH * world  # channel-samples through the main loop
    value = units_per_step / (ms_per_step / 1e3)

    hb = last["hits"]
    n_onsets = int(hb.onset_counts.sum().item())
    n_hits = int(hb.rec.shape[0])
    n_located = int((hb.loc_status == 0).sum().item())
    if dist is not None:
        t = torch.tensor([n_onsets, n_hits, n_located], device="cuda")
        dist.all_reduce(t)
        n_onsets, n_hits, n_located = (int(v) for v in t.tolist())
    else:
        n_onsets, n_hits, n_located = n_onsets, n_hits, n_located
    peak, peak_src = _peak()
    bytes_per_unit = 4 if args.no_rel else 8
    alg_bytes = R * nb * BLOCK * N_CH * bytes_per_unit
    achieved = alg_bytes / (k1_ms / 1e3) / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes of one k1_detect launch from the committed ncu --set full capture of THIS build at THIS shape
        tj = json.loads((ROOT / "profiles" / "r02_k1_traffic.json").read_text())
        if tj.get("recordings") == R and abs(tj.get("seconds", 0) - args.seconds) < 1e-9 and bool(tj.get("rel")) != args.no_rel:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": alg_bytes, "kernel": "k1_detect",
                "kernel_ms": k1_ms, "peak_source": peak_src, "algorithmic_bytes_per_channel_sample": bytes_per_unit}

    parity = parity_sample(args, torch, x, hp, hb) if args.parity_recordings > 0 else None
    e2e = None if args.skip_e2e else run_e2e(args, torch, hp, x, dist, world, rank)  # every rank feeds its own GPU
    cpu = None
    if rank == 0 and not args.skip_cpu:
        cpu = run_cpu_baseline(args, x)
    del hb
    last.clear()
    hp._out = None
    hp._host = None
    del x
    torch.cuda.empty_cache()
    h16 = None
    if not args.skip_hits16:
        h16 = hits16_leg(args, torch, dist, world, rank, local, args.hits16_per_gpu * world, steps=3, warmup=2)
    line = None
    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value, "unit": "channel-samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, R), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clk.summary(), "onsets_per_step": n_onsets,
            "hits_per_step": n_hits, "located_hits_per_step": n_located,
            "localised_hits_per_sec": n_located / (ms_per_step / 1e3),
            "wall_ms_per_step": 1e3 * t_wall / args.steps, "parity_sample": parity, "hits16": h16,
        }
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def parity_sample(args, torch, x, hp, hb):
    """Self-check of the timed batch: K randomly chosen recordings of the step that was just timed go back to the
    host and through the CPU oracle (detector, grouping, lag refinement, multilateration); the device results of
    those recordings must be identical (onsets, groups, refined onsets, lags, None set), the positions within
    1e-4 relative and the envelope within 1e-5 relative (north_star)."""
    from oracle import oracle as orc
    from onset_fingerprinting_b200 import synth

    R = x.shape[0]
    rng = np.random.default_rng(12345)
    recs = sorted(rng.choice(R, size=min(args.parity_recordings, R), replace=False).tolist())
    ch_d, ix_d, cnt_d, rel_d = hp._out
    rec_all = hb.rec.cpu().numpy()
    out = {"recordings": len(recs), "onsets_equal": True, "groups_equal": True, "fixed_equal": True, "lags_equal": True,
           "none_set_equal": True, "xy_max_rel_err": 0.0, "rel_max_rel_err": 0.0, "rel_bit_equal_frac": 1.0,
           "onsets_checked": 0, "hits_checked": 0, "located_checked": 0, "checked_against": "oracle/oracle_c.c (CPU)"}
    ml = orc.Multilaterate3D(synth.SENSORS_3MIC, sr=SR, medium="air")
    biteq = []
    for r in recs:
        xr = x[r].cpu().numpy()
        c_o, o_o, rel_o = orc.detect_onsets_amplitude(xr, block_size=BLOCK, sr=SR, return_rel=rel_d is not None)
        k = int(cnt_d[r].item())
        out["onsets_equal"] &= (ch_d[r, :k].cpu().tolist() == c_o) and (ix_d[r, :k].cpu().tolist() == o_o)
        out["onsets_checked"] += len(o_o)
        if rel_d is not None:
            got = rel_d[r].cpu().numpy()
            err = np.abs(got - rel_o) / np.maximum(np.abs(rel_o), 1e-6)
            out["rel_max_rel_err"] = max(out["rel_max_rel_err"], float(err.max()))
            biteq.append(float((got == rel_o).mean()))
        groups = orc.find_onset_groups(o_o, c_o, 1000, N_CH)
        sel = rec_all == r
        g_d = hb.onsets.cpu().numpy()[sel]
        if groups is None:
            out["groups_equal"] &= len(g_d) == 0
            continue
        out["groups_equal"] &= np.array_equal(g_d, groups)
        fixed, fstat, lags = orc.fix_onsets(xr, groups, return_status=True)
        out["fixed_equal"] &= np.array_equal(hb.fixed.cpu().numpy()[sel], fixed)
        out["lags_equal"] &= np.array_equal(hb.lags.cpu().numpy()[sel], lags)
        out["hits_checked"] += len(fixed)
        xy_d, st_d = hb.xy.cpu().numpy()[sel], hb.loc_status.cpu().numpy()[sel]
        for h, row in enumerate(fixed):
            got, st = ml.locate_hit([0, 1, 2], row[:3])
            out["none_set_equal"] &= (got is None) == (st_d[h] != 0)
            if got is not None and st_d[h] == 0:
                out["located_checked"] += 1
                e = np.abs(xy_d[h] - np.asarray(got)) / np.maximum(np.abs(np.asarray(got)), 1e-3)
                out["xy_max_rel_err"] = max(out["xy_max_rel_err"], float(e.max()))
    if biteq:
        out["rel_bit_equal_frac"] = float(np.mean(biteq))
    out["ok"] = bool(out["onsets_equal"] and out["groups_equal"] and out["fixed_equal"] and out["lags_equal"]
                     and out["none_set_equal"] and out["xy_max_rel_err"] <= 1e-4 and out["rel_max_rel_err"] <= 1e-5)
    for k_ in ("onsets_equal", "groups_equal", "fixed_equal", "lags_equal", "none_set_equal"):
        out[k_] = bool(out[k_])
    return out


def host_memory_available() -> float:
    """Bytes this job may still take from host memory: what the kernel reports as available, capped by the container's
    cgroup limit minus its current usage (pinned staging buffers beyond that get the job killed, not an exception)."""
    import psutil

    avail = float(psutil.virtual_memory().available)
    for lim, use in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                     ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            limit = open(lim).read().strip()
            if limit != "max":
                avail = min(avail, float(limit) - float(open(use).read().strip()))
        except Exception:
            pass
    return max(avail, 0.0)


def run_e2e(args, torch, hp, x, dist=None, world=1, rank=0):
    """Same metric end to end through the host-buffer entry of the public API (pipeline.HotPath.run_host): the
    batch starts in pinned HOST memory, is uploaded in time segments while the detector runs, goes through
    grouping, lag refinement and multilateration on the device, and the per-hit records (and, in drop-in mode,
    the whole rel envelope) return to pinned host memory inside the timed region.  Both modes are reported;
    `value` is the drop-in one (the mode `value` of the line is quoted in) unless --no-rel.  With N ranks every
    rank pushes its own recordings through its own GPU at the same time; value = all ranks' units over the
    slowest rank's wall time."""
    import psutil

    R, N, Cn = x.shape
    per_rec = N * Cn * 4
    avail = host_memory_available() / max(world, 1)
    want = args.e2e_recordings if args.e2e_recordings > 0 else R
    modes = ["onsets_only"] if args.no_rel else ["onsets_only", "drop_in"]
    fit = int(0.4 * avail / (per_rec * (2 if "drop_in" in modes else 1)))  # pinned memory is locked: stay well inside
    Re = max(1, min(want, R, fit))
    from onset_fingerprinting_b200 import parallel

    numa = parallel.bind_to_gpu_numa(torch.cuda.current_device())  # before the pinned buffers are first touched
    xh = torch.empty((Re, N, Cn), dtype=torch.float32, pin_memory=True)
    xh.copy_(x[:Re])
    torch.cuda.synchronize()
    # Denominator of the end-to-end figure: a plain pinned cudaMemcpyAsync of the same bytes, every rank at the same
    # time (what the host side and the PCIe links sustain with N uploads in flight), H2D alone and H2D + D2H together.
    ceil = {}
    scratch = xd_like = x[:Re]  # the upload lands in the resident device copy itself (same bytes): no extra HBM
    side = torch.cuda.Stream()
    ceil["_buf"] = torch.empty((min(Re, 256), N, Cn), dtype=torch.float32, pin_memory=True)  # allocated outside the timing
    for name, both in (("h2d", False), ("h2d_d2h", True)):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        scratch.copy_(xh, non_blocking=True)
        if both:
            with torch.cuda.stream(side):
                xh2 = ceil["_buf"]
                for lo in range(0, Re, xh2.shape[0]):  # the D2H stream lands in a small pinned buffer (host memory)
                    n = min(xh2.shape[0], Re - lo)
                    xh2[:n].copy_(xd_like[lo:lo + n], non_blocking=True)
        torch.cuda.synchronize()
        dtc = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dtc], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtc = float(t.item())
        ceil[name + "_gbs_per_rank"] = xh.numel() * 4 / dtc / 1e9
        ceil[name + "_ms"] = dtc * 1e3
    ceil.pop("_buf", None)
    hp_e = hp if Re == R else None
    if hp_e is None:
        from onset_fingerprinting_b200 import pipeline, synth

        hp_e = pipeline.HotPath(Re, N_CH, synth.SENSORS_3MIC, medium="air", sr=SR, block_size=BLOCK)
    xd = x[:Re]  # the resident device copy the upload lands in (lag refinement reads its sections from it)
    res = {}
    relh = None
    for mode in modes:
        if mode == "drop_in":
            relh = torch.empty((Re, (N // BLOCK) * BLOCK, Cn), dtype=torch.float32, pin_memory=True)
        out = hp_e.run_host(xh, rel_host=relh, x_dev=xd)  # warm (allocations, pinned result buffers)
        reps = 2
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = hp_e.run_host(xh, rel_host=relh, x_dev=xd)
        dt = (time.perf_counter() - t0) / reps
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        units = world * Re * (N // BLOCK) * BLOCK * N_CH
        d2h = sum(int(v.numel() * v.element_size()) for v in out.values()) + (relh.numel() * 4 if relh is not None else 0)
        res[mode] = {"value": units / dt, "ms": dt * 1e3, "h2d_bytes_per_step": int(world * xh.numel() * 4),
                     "d2h_bytes_per_step": int(world * d2h), "h2d_gbs_per_rank": xh.numel() * 4 / dt / 1e9,
                     "hits": int(out["rec"].shape[0]) * world, "located": int((out["loc_status"] == 0).sum()) * world,
                     "localised_hits_per_sec": int((out["loc_status"] == 0).sum()) * world / dt}
    head = res["onsets_only" if args.no_rel else "drop_in"]
    del xh, relh
    return {"value": head["value"], "unit": "channel-samples/s", "h2d_bytes_per_step": head["h2d_bytes_per_step"],
            "d2h_bytes_per_step": head["d2h_bytes_per_step"], "ms": head["ms"], "recordings": world * Re,
            "recordings_of_batch": f"{Re} of {R} per GPU" + ("" if Re == R else " (host memory bound)"),
            "mode": "onsets_only" if args.no_rel else "drop_in (rel envelope copied back to the host)",
            "modes": res, "numa": numa, "plain_copy_ceiling": ceil,
            "vs_plain_copy": {m: r["h2d_gbs_per_rank"] / ceil["h2d_d2h_gbs_per_rank" if m == "drop_in" else "h2d_gbs_per_rank"]
                              for m, r in res.items()},
            "stages": "upload + detector + grouping + lag refinement + multilateration + results to host",
            "entry": "pipeline.HotPath.run_host (one call per rank, concurrently): ofp_copy2d_async + ofp_detect_offline / "
                     "ofp_detect_continue per time segment, then ofp_group_onsets / ofp_fix_onsets_ex / ofp_locate_hits"}


def run_cpu_baseline(args, x):
    """The reference's CPU path -- whole chain -- on a bounded sample of the SAME device-generated audio."""
    cores = host_cores()
    # ~2.3e6 channel-samples/s/core through the whole chain: size the sample for about args.cpu_seconds of work
    per_rec = x.shape[1] * x.shape[2]
    n_rec = int(max(cores, min(x.shape[0], args.cpu_seconds * 2.3e6 * cores / per_rec)))
    n_rec = max(cores, n_rec // cores * cores)
    xs = x[:n_rec].cpu().numpy()
    chain = CpuChain(xs, cores)
    r = chain.run()
    chain.close()
    return {"value": r["rate"], "unit": "channel-samples/s", "cores": cores, "kind": cpu_kind(chain),
            "sample": f"{n_rec} of the step's recordings ({xs.size} channel-samples, {r['seconds']:.1f} s), whole chain",
            "note": CPU_NOTE, "onsets": r["onsets"], "hits": r["hits"], "located": r["located"],
            "localised_hits_per_sec": r["hits_per_sec"]}


def _peak():
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s"


HITS16 = dict(Cn=16, L=768, tol=150, cut=20)
HITS16_KW = dict(filter_size=7, d=1, take_abs=True, normalization_cutoff=20, onset_tolerance=150)


def run_hits16(args):
    """--workload hits16: configs[2] on its own (1 M hits sharded over the ranks, strong scaling)."""
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = hits16_leg(args, torch, dist, world, rank, local, args.hits, args.steps, args.warmup)
    if rank == 0:
        d.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                  "scaling": "strong", "vs_baseline": None, "data": "synthetic"})
        print(json.dumps(d))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


def hits16_leg(args, torch, dist, world, rank, local, hits_total, steps, warmup):
    """configs[2]: 16-channel mesh hit mining -- K4 (lag refinement, tol 150 / cutoff 20 / d=1 / abs /
    median 7 as in notebooks/refresh.org:1507-1509) + K5 on the first three arrivals; hits sharded by rank.
    Returns the leg's record (rank 0; None elsewhere): value, roofline of k4_fix, cpu_baseline, e2e."""
    from onset_fingerprinting_b200 import detection, multilateration, parallel, synth

    Cn, L, tol, cut = HITS16["Cn"], HITS16["L"], HITS16["tol"], HITS16["cut"]
    lo, hi = parallel.shard_range(hits_total, rank, world)
    H = hi - lo
    look = tol + cut
    # one section per hit: a burst whose wavefront reaches the 16 mesh sensors between look and look+420
    x = synth.drum_batch_device(H, L, sensors=synth.SENSORS_16MESH, medium="drumhead", seed=7, hit_period=10.0,
                                first_hit=look + 4, tail_guard=0, rec_offset=lo)
    g = torch.Generator(device="cuda").manual_seed(rank)
    onsets = (look + 4 + torch.randint(0, 400, (H, Cn), generator=g, device="cuda")).to(torch.int32)
    ml = multilateration.Multilaterate3D(synth.SENSORS_16MESH, sr=SR, medium="drumhead")
    kw = dict(max_section=L, **HITS16_KW)

    def step(events=None):
        if events:
            events[0].record()
        fixed, lags, st = detection.fix_onsets_batch(x, None, onsets, **kw)
        if events:
            events[1].record()
        first3 = torch.argsort(fixed, dim=1, stable=True)[:, :3].to(torch.int32)
        on3 = torch.gather(fixed, 1, first3.long())
        xy, lst = ml.locate_batch(on3, first3)
        return fixed, lags, st, first3, xy, lst

    for _ in range(warmup):
        out = step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1, ek = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), []
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(steps):
            ek.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            fixed, lags, st, first3, xy, lst = step(ek[-1])
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    ms /= steps
    k4_ms = float(np.mean([a.elapsed_time(b) for a, b in ek]))
    peak, src = _peak()
    bytes_per_hit = 4 * Cn * L + 16 * Cn + 48
    achieved = bytes_per_hit * H / (k4_ms / 1e3) / 1e9
    ok = int((st == 0).sum().item()); loc = int((lst == 0).sum().item())
    if dist is not None:
        t = torch.tensor([ok, loc], device="cuda"); dist.all_reduce(t); ok, loc = (int(v) for v in t.tolist())
    # How far fix_onsets moves the REFERENCE (first-arriving) channel's onset over the 15 pairs of a hit:
    # the reason the pairs of a hit cannot share one lag window, i.e. one batched (tensor-core) contraction.
    ref = torch.argmin(onsets, dim=1, keepdim=True)
    dr = (torch.gather(fixed, 1, ref) - torch.gather(onsets, 1, ref)).abs().float().flatten()[st == 0]
    drift = {"mean_abs": float(dr.mean()), "p50": float(dr.median()), "p90": float(dr.quantile(0.9)),
             "max": float(dr.max()), "frac_above_tol": float((dr > tol).float().mean())} if dr.numel() else None
    import ctypes as C
    from onset_fingerprinting_b200 import _lib
    cs = (C.c_uint64 * 4)()
    _lib.check(_lib.lib().ofp_cc_screen_stats(cs, 1))
    screen = {"pairs": int(cs[0]), "single_survivor": int(cs[1]), "few_survivors_exact": int(cs[2]),
              "exact_all_lags": int(cs[3])}
    cpu = e2e = parity = None
    if rank == 0:
        parity = hits16_parity(x, onsets, fixed, lags, st, first3, xy, lst, min(args.parity_recordings * 8, H))
        if not args.skip_cpu:
            cpu = hits16_cpu_baseline(x, onsets, args.cpu_seconds)
        if not args.skip_e2e:
            e2e = hits16_e2e(torch, detection, ml, x, onsets, kw)
    del x
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {
        "metric": "hits/sec through lag refinement + multilateration (16-channel hit mining)",
        "value": hits_total / (ms / 1e3), "unit": "hits/s", "ms_per_step": ms, "dtype": "f64 accumulate / f32 data",
        "localised_hits_per_sec": loc / (ms / 1e3),
        "config": {"workload": f"configs[2]: {hits_total} hits x 16 ch, section {L}, tol {tol}, cutoff {cut}, d=1, abs, "
                               "median 7; K5 on the first three arrivals", "hits_per_gpu": H,
                   "l2": "48.6 GB of sections per 1M hits, larger than L2"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "k4_fix", "kernel_ms": k4_ms, "peak_source": src,
                     "algorithmic_bytes_per_hit": bytes_per_hit,
                     "note": "K4 at 16 ch is instruction bound (median filter, section preparation, float32 "
                             "lag screening, reductions), not HBM bound"},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": 2 * steps, "clocks": clk.summary(),
        "fix_ok": ok, "located": loc, "cc_screening": screen, "parity_sample": parity,
        "ref_onset_drift": drift}


def hits16_parity(x, onsets, fixed, lags, st, first3, xy, lst, n):
    """Self-check of the timed hits: the first n hits of the step through the CPU oracle."""
    from oracle import oracle as orc
    from onset_fingerprinting_b200 import synth

    xs, on = x[:n].cpu().numpy(), onsets[:n].cpu().numpy().astype(np.int64)
    fx, lg, ss = fixed[:n].cpu().numpy(), lags[:n].cpu().numpy(), st[:n].cpu().numpy()
    f3, xyh, lsh = first3[:n].cpu().numpy(), xy[:n].cpu().numpy(), lst[:n].cpu().numpy()
    ml = orc.Multilaterate3D(synth.SENSORS_16MESH, sr=SR, medium="drumhead")
    eq_f = eq_l = eq_s = eq_n = True
    err = 0.0
    nloc = 0
    for h in range(n):
        f_o, s_o, l_o = orc.fix_onsets(xs[h], on[h:h + 1], return_status=True, **HITS16_KW)
        eq_f &= np.array_equal(f_o[0], fx[h]); eq_l &= np.array_equal(l_o[0], lg[h]); eq_s &= int(s_o[0]) == int(ss[h])
        got, stt = ml.locate_hit(f3[h], f_o[0][f3[h]])
        eq_n &= (got is None) == (lsh[h] != 0)
        if got is not None and lsh[h] == 0:
            nloc += 1
            err = max(err, float((np.abs(xyh[h] - np.asarray(got)) / np.maximum(np.abs(np.asarray(got)), 1e-3)).max()))
    return {"hits": n, "fixed_equal": bool(eq_f), "lags_equal": bool(eq_l), "status_equal": bool(eq_s),
            "none_set_equal": bool(eq_n), "located_checked": nloc, "xy_max_rel_err": err,
            "ok": bool(eq_f and eq_l and eq_s and eq_n and err <= 1e-4), "checked_against": "oracle/oracle_c.c (CPU)"}


def _h16_worker(i):
    from oracle import ref_chain

    if "loc16" not in _CPU:
        from onset_fingerprinting_b200 import synth

        _CPU["loc16"] = ref_chain.Locator(synth.SENSORS_16MESH, sr=SR, medium="drumhead")
    xs, on, cores = _CPU["h16"]
    n = loc = 0
    for k in range(i, len(xs), cores):
        try:
            f = ref_chain.fix_onsets(xs[k], on[k:k + 1], **HITS16_KW)[0]
        except ValueError:  # the reference raises on these hits
            n += 1
            continue
        f3 = np.argsort(f, kind="stable")[:3]
        loc += _CPU["loc16"].locate_hit(f[f3], f3) is not None
        n += 1
    return n, loc


def _hits16_cpu(xs, on, cores):
    import multiprocessing as mp

    _CPU["h16"] = (xs, on, cores)
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_h16_worker, [len(xs)] * cores)  # start-up outside the timing (empty slices)
        t0 = time.perf_counter()
        res = pool.map(_h16_worker, range(cores))
        dt = time.perf_counter() - t0
    done, loc = sum(r[0] for r in res), sum(r[1] for r in res)
    return done, loc, dt


def hits16_cpu_baseline(x, onsets, budget_s):
    """The reference's numpy/scipy calls per hit (oracle/ref_chain.py: median_filter + np.correlate per channel pair
    + fsolve) over a bounded sample of the step's hits on all host cores."""
    cores = host_cores()
    n = int(min(x.shape[0], max(cores * 4, 4000 * cores * budget_s / 15.0)))  # ~9e3 hits/s on 16 cores
    xs, on = x[:n].cpu().numpy(), onsets[:n].cpu().numpy().astype(np.int64)
    done, loc, dt = _hits16_cpu(xs, on, cores)
    return {"value": done / dt, "unit": "hits/s", "cores": cores, "kind": "port", "located": loc,
            "sample": f"{done} of the step's hits, lag refinement + multilateration at the reference's granularity "
                      f"(oracle/ref_chain.py), {dt:.1f} s"}


def hits16_reference(cores, budget_s):
    """--impl reference: the same CPU path on host-generated sections (oracle/make_golden.py:hits16_sections)."""
    from oracle.make_golden import hits16_sections

    uniq, reps = 100 * cores, max(1, int(40 * budget_s / 15.0))
    xs, on = hits16_sections(n_hits=uniq, seed=62)
    xs, on = np.concatenate([xs] * reps), np.concatenate([on] * reps)  # the per-hit cost does not depend on uniqueness
    done, loc, dt = _hits16_cpu(xs, on, cores)
    return {"metric": "hits/sec through lag refinement + multilateration (16-channel hit mining)", "value": done / dt,
            "unit": "hits/s", "cores": cores, "kind": "port", "located": loc,
            "sample": f"{done} synthetic 16-channel hits ({uniq} distinct, section 768, tol 150, cutoff 20, d=1, abs, "
                      f"median 7), {dt:.1f} s"}


def hits16_e2e(torch, detection, ml, x, onsets, kw, n=40000):
    """Same metric with the sections in pinned host memory: copy in, K4 + K5, results back."""
    n = min(n, x.shape[0])
    xh = torch.empty((n,) + tuple(x.shape[1:]), dtype=torch.float32, pin_memory=True)
    xh.copy_(x[:n])
    oh = torch.empty((n, onsets.shape[1]), dtype=torch.int32, pin_memory=True)
    oh.copy_(onsets[:n])
    torch.cuda.synchronize()

    from onset_fingerprinting_b200 import hostpipe

    def one(xd, od):
        fixed, lags, st = detection.fix_onsets_batch(xd, None, od, **kw)
        first3 = torch.argsort(fixed, dim=1, stable=True)[:, :3].to(torch.int32)
        xy, lst = ml.locate_batch(torch.gather(fixed, 1, first3.long()), first3)
        return fixed, xy, lst

    keep = {}

    def call():  # uploads of chunk i+1 overlap the kernels of chunk i (hostpipe.run_chunked)
        keep["outs"] = hostpipe.run_chunked([xh, oh], one, chunk=max(1, n // 8), outs=keep.get("outs"))

    call()
    call()
    reps, each = 5, []
    for _ in range(reps):
        t0 = time.perf_counter()
        call()
        each.append(time.perf_counter() - t0)
    dt = float(np.median(each))  # (the first process on a fresh box has shown single calls 4x slower than the rest)
    return {"value": n / dt, "unit": "hits/s", "h2d_bytes_per_step": int(xh.numel() * 4 + oh.numel() * 4),
            "d2h_bytes_per_step": int(n * (onsets.shape[1] * 4 + 16 + 4)), "hits": n, "ms": dt * 1e3,
            "ms_each": [round(1e3 * t, 2) for t in each]}


def run_spectral(args):
    """configs[4], feature half: spectral-flux onset strength (2048-point Hann rFFT every 128 samples of the
    channel mean, recording.py:273-311) over the configs[1] batch; K2 = csrc/spectral_flux.cu."""
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from onset_fingerprinting_b200 import spectral, synth
    from oracle import spectral_np

    R, N = args.recordings, int(args.seconds * SR)
    x = synth.drum_batch_device(R, N, seed=1234, rec_offset=rank * R)
    for _ in range(args.warmup):
        flux = spectral.spectral_flux_batch(x, 2048, 128)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            flux = spectral.spectral_flux_batch(x, 2048, 128)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    ms /= args.steps
    frames = R * (N // 128)
    flops = frames * 5.0 * 2048 * 11  # 5 N log2 N per frame
    peak, src = _peak()
    alg_bytes = R * N * N_CH * 4 + frames * 4
    cpu = None
    if rank == 0 and not args.skip_cpu:
        xs = x[0, : 96000].cpu().numpy()
        t0 = time.perf_counter()
        spectral_np.onset_strength(xs)
        dt = time.perf_counter() - t0
        cpu = {"value": xs.shape[0] / dt, "unit": "samples/s", "cores": 1, "kind": "port",
               "sample": f"1 s of one recording, numpy rfft per hop (oracle/spectral_np.py), {dt:.2f} s"}
    if rank == 0:
        print(json.dumps({
            "metric": "samples/sec through the spectral-flux onset-strength feature", "value": world * R * N / (ms / 1e3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[4] (features): {R} recordings x {args.seconds:g} s, 2048-point rFFT, hop 128, "
                                   "3-channel mean", "l2": "inputs larger than L2"},
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg_bytes / (ms / 1e3) / 1e9 / peak, "traffic": None, "kernel": "k2_flux",
                         "kernel_ms": ms, "peak_source": src,
                         "note": f"FP32-compute bound: {flops / (ms / 1e3) / 1e12:.1f} TFLOP/s of FFT arithmetic"},
            "cpu_baseline": cpu, "e2e": None, "gpu_launches": args.steps, "clocks": clk.summary(),
            "frames_per_step": frames}))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


def run_cnn(args):
    """configs[4], classifier half: model.CNN inference (model.py:52-120, reference defaults: 3 x 256 windows,
    Conv1d 3->8->16 k=3 + SiLU, Linear 4096->2) on --windows onset windows per GPU; K6 = csrc/cnn_infer.cu."""
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from onset_fingerprinting_b200 import model

    torch.manual_seed(7)
    cc = args.network == "cccnn"
    m = (model.CCCNN(256, 2) if cc else model.CNN(256, 2)).cuda()
    n = args.windows if not cc else min(args.windows, 200000)
    g = torch.Generator(device="cuda"); g.manual_seed(100 + rank)
    x = torch.randn((n, 3, 256), device="cuda", generator=g) * 0.1
    for _ in range(args.warmup):
        y = m(x)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            y = m(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    ms /= args.steps
    macs = 256 * (8 * 3 * 3 + 16 * 8 * 3) + 2 * 4096
    if cc:  # per channel: conv 1->8->16, then 16 maps x V^2 / 2 products of the auto-correlation (lags >= 0)
        macs = 3 * (256 * (8 * 3 + 16 * 8 * 3) + 16 * 256 * 257 // 2) + 2 * 3 * 511
    peak, src = _peak()
    alg_bytes = n * (3 * 256 * 4 + 2 * 4)
    cpu = e2e = None
    if rank == 0 and not args.skip_cpu and not cc:
        # the reference runs this network through stock torch modules; same modules on the host cores
        mc = torch.nn.Sequential(m.conv_layers, torch.nn.Flatten(1), m.fc).cpu().eval() if not cc else None
        xs = x[:50000].cpu()
        with torch.no_grad():
            mc(xs[:1000])
            t0 = time.perf_counter(); mc(xs); dt = time.perf_counter() - t0
        cpu = {"value": xs.shape[0] / dt, "unit": "windows/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{xs.shape[0]} windows through torch.nn modules on the host (model.py:112-117), {dt:.2f} s"}
        m.cuda()
    if rank == 0:
        ne = min(n, 200000)
        xh = torch.empty((ne, 3, 256), dtype=torch.float32, pin_memory=True); xh.copy_(x[:ne])
        torch.cuda.synchronize()
        from onset_fingerprinting_b200 import hostpipe

        res = hostpipe.run_chunked([xh], m, chunk=ne // 8)
        t0 = time.perf_counter()
        for _ in range(3):
            res = hostpipe.run_chunked([xh], m, chunk=ne // 8, outs=res)
        dt = (time.perf_counter() - t0) / 3
        e2e = {"value": ne / dt, "unit": "windows/s", "h2d_bytes_per_step": int(xh.numel() * 4),
               "d2h_bytes_per_step": int(ne * 8), "windows": ne, "ms": dt * 1e3}
        print(json.dumps({
            "metric": f"onset windows/sec through model.{'CCCNN' if cc else 'CNN'} inference", "value": world * n / (ms / 1e3),
            "unit": "windows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[4] (classifier): {n} windows x 3 ch x 256 samples per GPU, CNN [8, 16] k=3 SiLU "
                                   "+ Linear 4096->2, random-init weights", "l2": "inputs larger than L2"},
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg_bytes / (ms / 1e3) / 1e9 / peak, "traffic": None, "kernel": "k6_cccnn" if cc else "k6_cnn_tc",
                         "kernel_ms": ms, "peak_source": src,
                         "note": f"compute bound: {2 * macs * n / (ms / 1e3) / 1e12:.1f} TFLOP/s ({macs} MAC per window; conv2 as "
                                 "3xTF32 mma.sync on the tensor cores, conv1 / SiLU / Linear on the FP32 pipe)"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps, "clocks": clk.summary()}))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


def run_realtime(args):
    """configs[3]: S concurrent streams; per 128-sample block one K1 launch (detector) and one
    ofp_stream_locate launch (Multilaterate3D.locate's group state machine, one thread per stream).
    Latency = launch to the located positions being on the host."""
    import torch

    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    from onset_fingerprinting_b200 import synth
    from onset_fingerprinting_b200.realtime import audio as rt

    S, nblk = args.streams, args.blocks
    x = synth.drum_batch_device(S, nblk * BLOCK, seed=3, first_hit=5000)
    conf = {"sensor_locations": synth.SENSORS_3MIC, "medium": "air", "c": None}

    def drive(step, reset):
        """nblk consecutive blocks; latency = call to the host knowing which streams located a hit."""
        for b in range(min(20, nblk)):
            step(b)
        torch.cuda.synchronize()
        reset()
        torch.cuda.synchronize()
        lat, located = [], 0
        t0 = time.perf_counter()
        for b in range(nblk):
            t = time.perf_counter()
            located += step(b)
            lat.append(time.perf_counter() - t)
        total = time.perf_counter() - t0
        lat = np.asarray(lat) * 1e6
        return {"total_s": total, "located": located, "p50": float(np.percentile(lat, 50)),
                "p99": float(np.percentile(lat, 99)), "ms_per_block": 1e3 * total / nblk}

    # (a) Python-driven path: two C-ABI calls per block (ofp_detect_block, ofp_stream_locate) + result read
    sl = rt.StreamLocatorBatch(S, conf)

    def py_step(b):
        xy, found = sl.detect_hits(x[:, b * BLOCK:(b + 1) * BLOCK])
        return int((found.cpu() == 1).sum())

    py = drive(py_step, sl.reset)
    # (b) native session: one replayed CUDA graph per block (csrc/realtime.cu), results in pinned host memory
    rs = rt.RealtimeSession(S, conf)

    def graph_step(b):
        xy, found = rs.detect_hits(x[:, b * BLOCK:(b + 1) * BLOCK])
        return int((found == 1).sum())

    gr = drive(graph_step, rs.reset)
    # (c) the same with the blocks arriving in pinned HOST memory (the audio-callback shape): copy in the step
    xh = torch.empty((nblk, S, BLOCK, N_CH), dtype=torch.float32, pin_memory=True)
    xh.copy_(x.view(S, nblk, BLOCK, N_CH).transpose(0, 1))

    def host_step(b):
        xy, found = rs.detect_hits(xh[b])
        return int((found == 1).sum())

    ho = drive(host_step, rs.reset)
    assert py["located"] == gr["located"] == ho["located"], (py["located"], gr["located"], ho["located"])
    # (d) what the reference's callback actually runs (realtime/audio.py:69 passes rec_audio): every new pair is
    # refined by cross-correlation on the stream's ring of recent audio (multilateration.py:457-501)
    rr = rt.RealtimeSession(S, conf, ring_rows=args.ring_rows)

    def ring_step(b):
        xy, found = rr.detect_hits(x[:, b * BLOCK:(b + 1) * BLOCK])
        return int((found == 1).sum())

    rg = drive(ring_step, rr.reset)

    def ring_host_step(b):
        xy, found = rr.detect_hits(xh[b])
        return int((found == 1).sum())

    rh = drive(ring_host_step, rr.reset)
    units = S * nblk * BLOCK * N_CH
    print(json.dumps({
        "metric": "channel-samples/sec, realtime block streams", "value": units / gr["total_s"],
        "unit": "channel-samples/s", "n_gpus": 1, "steps": nblk, "warmup": 20, "ms_per_step": gr["ms_per_block"],
        "higher_is_better": True, "scaling": "replicas only", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[3]: {S} concurrent 3-mic streams, {BLOCK}-sample blocks, realtime detector "
                               "settings, detector + streaming locate per block, one replayed CUDA graph per block"},
        "latency_us": {"p50": gr["p50"], "p99": gr["p99"], "budget_us": 1e6 * BLOCK / SR},
        "with_ring_refinement": {"note": "locate(..., rec_audio): per-stream device ring of %d rows, bounded-lag CC + "
                                         "adjust_onset for every new pair, one warp per stream" % args.ring_rows,
                                 "value": units / rg["total_s"], "ms_per_step": rg["ms_per_block"],
                                 "latency_us": {"p50": rg["p50"], "p99": rg["p99"]}, "located_hits": rg["located"],
                                 "e2e": {"value": units / rh["total_s"], "latency_us": {"p50": rh["p50"], "p99": rh["p99"]}}},
        "python_driven": {"value": units / py["total_s"], "latency_us": {"p50": py["p50"], "p99": py["p99"]},
                          "ms_per_step": py["ms_per_block"]},
        "located_hits": gr["located"], "localised_hits_per_sec": gr["located"] / gr["total_s"],
        "gpu_launches": 3 * nblk,
        "e2e": {"value": units / ho["total_s"], "unit": "channel-samples/s",
                "h2d_bytes_per_step": S * BLOCK * N_CH * 4, "d2h_bytes_per_step": S * 20,
                "latency_us": {"p50": ho["p50"], "p99": ho["p99"]}},
        "cpu_baseline": None, "roofline": None}))


def run_k1_only(args):
    """Detector kernel alone at the headline shape (A/B runs of kernel variants, the speed-of-light ladder)."""
    import torch

    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    from onset_fingerprinting_b200 import detection, synth

    R, N = args.recordings, int(args.seconds * SR)
    nb = N // BLOCK
    x = synth.drum_batch_device(R, N, seed=1234)
    det = detection.BatchedOnsetDetector(R, N_CH, BLOCK, sr=SR)
    cap = det.default_cap(N)
    out = (torch.empty((R, cap), dtype=torch.int32, device="cuda"), torch.empty((R, cap), dtype=torch.int32, device="cuda"),
           torch.empty((R,), dtype=torch.int32, device="cuda"),
           None if args.no_rel else torch.empty((R, nb * BLOCK, N_CH), dtype=torch.float32, device="cuda"))
    ms = []
    for i in range(args.warmup + args.steps):
        det.reset()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        det.detect_offline(x, int(0.5 * SR), out=out)
        b.record()
        torch.cuda.synchronize()
        if i >= args.warmup:
            ms.append(a.elapsed_time(b))
    k1_ms = float(np.mean(ms))
    peak, src = _peak()
    alg = R * nb * BLOCK * N_CH * (4 if args.no_rel else 8)
    print(json.dumps({"metric": "k1_detect alone", "value": R * nb * BLOCK * N_CH / (k1_ms / 1e3), "unit": "channel-samples/s",
                      "onsets": int(out[2].sum().item()),
                      "roofline": {"kernel_ms": k1_ms, "achieved": alg / (k1_ms / 1e3) / 1e9, "peak": peak,
                                   "frac": alg / (k1_ms / 1e3) / 1e9 / peak, "peak_source": src}}))


def _quiet_stdout():
    """stdout carries exactly one JSON line: libraries that write to file descriptor 1 themselves (NCCL prints its
    version there when NCCL_DEBUG is set in the environment) are pointed at stderr, print() keeps the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real


def main():
    args = parse()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.k1_only:
        run_k1_only(args)
    elif args.workload == "hits16":
        run_hits16(args)
    elif args.workload == "realtime":
        run_realtime(args)
    elif args.workload == "spectral":
        run_spectral(args)
    elif args.workload == "cnn":
        run_cnn(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
