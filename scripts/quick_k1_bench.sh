#!/bin/bash
python bench.py --steps 3 --warmup 3 --seconds 2 --skip-cpu --e2e-recordings 8 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,1), 'G ch-samp/s  k1_ms', round(d['roofline']['kernel_ms'],2))"
