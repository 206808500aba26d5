"""Where a bench step spends its time besides k1_detect: CUDA-event time per stage + host wall time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onset_fingerprinting_b200 import detection, pipeline, synth

R, N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, 480000
x = synth.drum_batch_device(R, N, seed=1234)
hp = pipeline.HotPath(R, 3, synth.SENSORS_3MIC, medium="air", sr=96000)
for _ in range(2):
    hp.run(x, return_rel=True)
torch.cuda.synchronize()
det = hp.det
names = ["reset", "k1", "group", "fix", "locate"]
for rep in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    wall = [time.perf_counter()]
    ev[0].record()
    det.reset(); ev[1].record(); wall.append(time.perf_counter())
    ch, ix, cnt, rel = det.detect_offline(x, 48000, out=hp._out); ev[2].record(); wall.append(time.perf_counter())
    hit_rec, hit_on, _ = detection.find_onset_groups_batch(ch, ix, cnt, 3, **hp.group_kw); ev[3].record(); wall.append(time.perf_counter())
    ms = 1081 if os.environ.get("FIXED_SECTION") else None  # None: sized by the largest onset spread (pipeline default)
    fixed, lags, fstat = detection.fix_onsets_batch(x, hit_rec, hit_on, max_section=ms); ev[4].record(); wall.append(time.perf_counter())
    xy, lstat = hp.ml.locate_batch(fixed); ev[5].record(); wall.append(time.perf_counter())
    torch.cuda.synchronize()
    print(" | ".join(f"{n}: gpu {ev[i].elapsed_time(ev[i + 1]):7.2f} ms host {1e3 * (wall[i + 1] - wall[i]):7.2f} ms" for i, n in enumerate(names)),
          f"| total gpu {ev[0].elapsed_time(ev[5]):.2f} ms")
