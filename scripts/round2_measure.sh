#!/bin/bash
# round-2 measurement set at the headline shape (10 000 x 5 s): tests, both bench arms, K1 ladder, ncu launch list,
# DRAM traffic of k1_detect, one --set full capture
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_final.log 2>&1; tail -3 gpurun_out/r02_gputest_final.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_final.err
python bench.py > gpurun_out/r02_bench_final.json 2>> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.err
scripts/k1_ladder.sh > gpurun_out/r02_k1_ladder.txt 2>&1; cat gpurun_out/r02_k1_ladder.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-hits16 --parity-recordings 0 > gpurun_out/r02_ncu_launches_final.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k1_detect -c 2 --csv --log-file gpurun_out/r02_k1_traffic_final.csv python bench.py --steps 1 --warmup 1 --k1-only > gpurun_out/r02_ncu_k1_traffic_final.log 2>&1
tail -2 gpurun_out/r02_k1_traffic_final.csv | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_detect -c 1 -o gpurun_out/r02_prof_k1_headline -f python bench.py --steps 1 --warmup 0 --k1-only > gpurun_out/r02_ncu_k1_headline.log 2>&1
tail -1 gpurun_out/r02_ncu_k1_headline.log
