#!/bin/bash
# round-2 measurement set at the headline shape (10 000 x 5 s): tests, default bench line, K1 ladder, ncu launch list,
# DRAM traffic of k1_detect, one --set full capture
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_c1.log 2>&1; tail -3 gpurun_out/r02_gputest_c1.log
python bench.py > gpurun_out/r02_bench_c1.json 2> gpurun_out/r02_bench_c1.err; tail -c 300 gpurun_out/r02_bench_c1.err
scripts/k1_ladder.sh > gpurun_out/r02_k1_ladder.txt 2>&1; cat gpurun_out/r02_k1_ladder.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c1.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-hits16 --parity-recordings 0 > gpurun_out/r02_ncu_launches_c1.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k1_detect -c 2 --csv --log-file gpurun_out/r02_k1_traffic_c1.csv python bench.py --steps 1 --warmup 1 --k1-only > gpurun_out/r02_ncu_k1_traffic_c1.log 2>&1
tail -2 gpurun_out/r02_k1_traffic_c1.csv | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_detect -c 1 -o gpurun_out/r02_prof_k1_headline -f python bench.py --steps 1 --warmup 0 --k1-only > gpurun_out/r02_ncu_k1_headline.log 2>&1
tail -2 gpurun_out/r02_ncu_k1_headline.log
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/r02_topo.txt; numactl -H >> gpurun_out/r02_topo.txt 2>&1; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c >> gpurun_out/r02_topo.txt
