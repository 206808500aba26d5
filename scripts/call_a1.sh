#!/bin/bash
# session call 1: ubench + full gpu tests + default bench line
./scripts/ubench/op_rates > gpurun_out/r02_ubench_op_rates.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_a1.log 2>&1; tail -3 gpurun_out/r02_gputest_a1.log
python bench.py > gpurun_out/r02_bench_a1.json 2> gpurun_out/r02_bench_a1.err; tail -c 600 gpurun_out/r02_bench_a1.err
cat gpurun_out/r02_ubench_op_rates.txt
