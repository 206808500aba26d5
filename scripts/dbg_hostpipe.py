import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onset_fingerprinting_b200 import detection, hostpipe, multilateration as mlm, synth
n, L, C = 16000, 768, 16
xh = (torch.randn(n, L, C) * 0.01).pin_memory()
oh = (torch.randint(200, 500, (n, C), dtype=torch.int32)).pin_memory()
kw = dict(filter_size=7, d=1, take_abs=True, normalization_cutoff=20, onset_tolerance=150, max_section=L)
def one(xd, od):
    fixed, lags, st = detection.fix_onsets_batch(xd, None, od, **kw)
    return fixed, st
def timed(label, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
    print(f"{label}: {1e3 * (time.perf_counter() - t0):.1f} ms"); return r
timed("direct warm", lambda: one(xh.cuda(non_blocking=True), oh.cuda(non_blocking=True)))
timed("direct", lambda: one(xh.cuda(non_blocking=True), oh.cuda(non_blocking=True)))
r = timed("chunked warm", lambda: hostpipe.run_chunked([xh, oh], one, chunk=2000))
for _ in range(3):
    r = timed("chunked", lambda: hostpipe.run_chunked([xh, oh], one, chunk=2000, outs=r))
s = torch.cuda.Stream()
def on_side():
    with torch.cuda.stream(s):
        xd = xh[:2000].cuda(non_blocking=True); od = oh[:2000].cuda(non_blocking=True)
        t0 = time.perf_counter(); res = one(xd, od); t1 = time.perf_counter()
    s.synchronize(); return t1 - t0
for _ in range(3):
    print("side-stream host time of one(): %.2f ms" % (1e3 * on_side()))
