#!/bin/bash
free -g > gpurun_out/r02_mem.txt; nproc >> gpurun_out/r02_mem.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_tests_b.txt
python bench.py > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_b.json 2>> gpurun_out/r02_bench_b.err
cat gpurun_out/r02_mem.txt gpurun_out/r02_tests_b.txt; tail -5 gpurun_out/r02_bench_b.err; cut -c1-1500 gpurun_out/r02_bench_b.json
