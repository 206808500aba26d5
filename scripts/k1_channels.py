"""k1_detect throughput for other channel counts (lane packing / tile length differ): R recordings x 5 s each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from onset_fingerprinting_b200 import detection

SR, B = 96000, 128
for C, R in ((1, 20000), (2, 15000), (3, 10000), (4, 8000), (8, 4000), (16, 2000), (32, 1000)):
    N = 5 * SR
    x = (1e-4 * torch.randn(R, N, C, device="cuda")).contiguous()
    det = detection.BatchedOnsetDetector(R, C, B, sr=SR)
    nb = N // B
    cap = det.default_cap(N)
    out = (torch.empty((R, cap), dtype=torch.int32, device="cuda"), torch.empty((R, cap), dtype=torch.int32, device="cuda"),
           torch.empty((R,), dtype=torch.int32, device="cuda"), torch.empty((R, nb * B, C), dtype=torch.float32, device="cuda"))
    ms = []
    for i in range(5):
        det.reset()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); det.detect_offline(x, SR // 2, out=out); b.record(); torch.cuda.synchronize()
        if i >= 2:
            ms.append(a.elapsed_time(b))
    t = float(np.mean(ms))
    warps = (R + 32 // C - 1) // (32 // C)
    print(f"C={C:2d} R={R:5d} warps={warps:5d} k1_ms={t:7.2f} ch-samples/s={R * N * C / t * 1e3:.3e} GB/s={R * N * C * 8 / t / 1e6:.0f}")
    del x, out, det
    torch.cuda.empty_cache()
