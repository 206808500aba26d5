#!/bin/bash
# Round measurement set (run under gpurun): every bench workload, then the ncu launch list and one
# --set full capture of the dominant kernel.  Outputs under gpurun_out/, copied to profiles/ by hand.
tag=${1:-r01b}
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_${tag}.json 2>> gpurun_out/bench_${tag}.err
python bench.py --workload hits16 --hits 200000 --steps 3 --warmup 3 > gpurun_out/bench_hits16_${tag}.json 2>> gpurun_out/bench_${tag}.err
python bench.py --workload realtime --steps 3 --warmup 3 > gpurun_out/bench_realtime_${tag}.json 2>> gpurun_out/bench_${tag}.err
python bench.py --workload spectral --recordings 2000 --steps 3 --warmup 3 > gpurun_out/bench_spectral_${tag}.json 2>> gpurun_out/bench_${tag}.err
python bench.py --workload cnn --steps 5 --warmup 3 > gpurun_out/bench_cnn_${tag}.json 2>> gpurun_out/bench_${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu --recordings 2000 > gpurun_out/ncu_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_detect -c 1 -o gpurun_out/prof_k1_${tag} -f \
    python bench.py --steps 1 --warmup 0 --skip-cpu --recordings 2000 --seconds 1 > gpurun_out/ncu_k1_${tag}.log 2>&1
ls -la gpurun_out | tail -12
