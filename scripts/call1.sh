#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_tests_a.txt
scripts/k1_ladder.sh > gpurun_out/r02_k1_ladder.txt 2>&1
python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --k1-only --recordings 11840 2>&1 | tail -1 >> gpurun_out/r02_k1_ladder.txt
python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --k1-only --no-rel 2>&1 | tail -1 >> gpurun_out/r02_k1_ladder.txt
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/r02_topo.txt; numactl -H >> gpurun_out/r02_topo.txt 2>&1
cat gpurun_out/r02_tests_a.txt gpurun_out/r02_k1_ladder.txt
