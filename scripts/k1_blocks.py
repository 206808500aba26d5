"""k1_detect throughput for other block sizes (10 000 x 5 s x 3 ch of noise): the block buffer and the way tiles
split at block ends change with B.  OFP_K1_TILE forces the tile length."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from onset_fingerprinting_b200 import detection

SR, C, R = 96000, 3, 10000
N = 5 * SR
x = (1e-4 * torch.randn(R, N, C, device="cuda")).contiguous()
for B in (32, 64, 128, 256):
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
    print(f"B={B:4d} tile={os.environ.get('OFP_K1_TILE', 'auto'):>4s} k1_ms={float(np.mean(ms)):7.2f}")
    del out, det
