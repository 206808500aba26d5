#!/bin/bash
# Speed-of-light ladder of k1_detect at the headline shape: build with scripts/k1_variants.sh ladderN -DOFP_K1_LADDER=N
# (N = 0..4) here, then run this under gpurun.  Prints k1_ms per rung; outputs of rungs < 5 are not the detector's.
run() { python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --k1-only 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', 'k1_ms %.2f'%r['kernel_ms'], 'frac %.3f'%r['frac'])"; }
for n in 0 1 2 3 4; do OFP_LIB=scripts/variants/libofp_k1_ladder$n.so run ladder$n; done
run full
