#!/bin/bash
python bench.py --workload realtime --steps 3 --warmup 3 > gpurun_out/r02_bench_realtime.json 2> gpurun_out/r02_bench_realtime.err; tail -c 400 gpurun_out/r02_bench_realtime.err; cut -c1-1800 gpurun_out/r02_bench_realtime.json
python bench.py --workload hits16 --hits 200000 --steps 3 --warmup 2 > gpurun_out/r02_bench_hits16.json 2> gpurun_out/r02_bench_hits16.err; tail -c 300 gpurun_out/r02_bench_hits16.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_hits16.json')); print('h16', d['value'], d['ref_onset_drift'], d['cc_screening'])"
ncu --set full --clock-control none -k regex:k6_cnn_tc -c 1 -o gpurun_out/r02_prof_k6_tc -f python bench.py --workload cnn --windows 200000 --steps 1 --warmup 0 --skip-cpu --skip-e2e > gpurun_out/r02_ncu_k6.log 2>&1; tail -1 gpurun_out/r02_ncu_k6.log
python bench.py --workload cnn --steps 5 --warmup 3 > gpurun_out/r02_bench_cnn.json 2>> gpurun_out/r02_bench_realtime.err
python bench.py --workload cnn --network cccnn --windows 200000 --steps 3 --warmup 3 > gpurun_out/r02_bench_cccnn.json 2>> gpurun_out/r02_bench_realtime.err
python bench.py --workload spectral --recordings 2000 --steps 3 --warmup 3 > gpurun_out/r02_bench_spectral.json 2>> gpurun_out/r02_bench_realtime.err
ls -la gpurun_out | tail -8
