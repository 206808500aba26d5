"""e2e leg only (ofp_detect_offline_host) with phase timings, plus the raw 1-D / 2-D pinned copy bandwidth."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onset_fingerprinting_b200 import _lib, detection, synth

R, N = int(sys.argv[1]) if len(sys.argv) > 1 else 3072, 480000
x = synth.drum_batch_device(R, N, seed=1234)
xh = torch.empty((R, N, 3), dtype=torch.float32, pin_memory=True); xh.copy_(x); del x
torch.cuda.synchronize()
d = torch.empty((R, 49152, 3), dtype=torch.float32, device="cuda")
rt = C.CDLL("libcudart.so.12")
for name, call in (("1-D 1.8 GB", lambda: d.view(-1).copy_(xh.view(-1)[: d.numel()], non_blocking=True)),
                   ("2-D 3072 x 590 KB", lambda: d.copy_(xh[:, :49152], non_blocking=True))):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); call(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: {d.numel() * 4 / dt / 1e9:.1f} GB/s")
p = detection.make_params(3, 128, sr=96000)
cap = 3 * (N // 1323 + 2)
ch = torch.empty((R, cap), dtype=torch.int32, pin_memory=True); ix = torch.empty_like(ch).pin_memory()
cnt = torch.empty((R,), dtype=torch.int32, pin_memory=True)
for i in range(3):
    t0 = time.perf_counter()
    _lib.check(_lib.lib().ofp_detect_offline_host(C.byref(p), C.c_void_p(xh.data_ptr()), C.c_int64(R), C.c_int64(N),
               C.c_int64(48000), None, C.c_void_p(ch.data_ptr()), C.c_void_p(ix.data_ptr()), C.c_void_p(cnt.data_ptr()), C.c_int32(cap)))
    print(f"call {i}: {1e3 * (time.perf_counter() - t0):.1f} ms, onsets {int(cnt.sum())}")
