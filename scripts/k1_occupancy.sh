#!/bin/bash
# K1 throughput vs number of recordings (= warps per scheduler) and the box's pinned H2D bandwidth
run() { python bench.py --steps 3 --warmup 3 --skip-cpu --e2e-recordings 8 "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$*', 'value %.3g'%d['value'], 'k1_ms %.2f'%r['kernel_ms'], 'frac %.3f'%r['frac'])"; }
run --recordings 10000 --seconds 5
run --recordings 10000 --seconds 5 --no-rel
run --recordings 20000 --seconds 5 --no-rel
run --recordings 20000 --seconds 2.5
run --recordings 40000 --seconds 1.25
python - <<'PY'
import torch, time
x = torch.empty(1 << 30, dtype=torch.float32, pin_memory=True)  # 4 GiB
d = torch.empty_like(x, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize()
    print("H2D pinned 4 GiB: %.1f GB/s" % (x.numel() * 4 / (time.perf_counter() - t0) / 1e9))
torch.cuda.synchronize(); t0 = time.perf_counter(); x.copy_(d, non_blocking=True); torch.cuda.synchronize()
print("D2H pinned 4 GiB: %.1f GB/s" % (x.numel() * 4 / (time.perf_counter() - t0) / 1e9))
PY
