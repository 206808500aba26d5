#!/bin/bash
# 2-GPU run of the default bench line (torchrun, NCCL) -- validates the multi-rank path of this build
free -g; cat /sys/fs/cgroup/memory.max 2>/dev/null; cat /sys/fs/cgroup/memory/memory.limit_in_bytes 2>/dev/null; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 3 --cpu-seconds 5 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; tail -c 600 gpurun_out/r02_bench_8gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_8gpu.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'ngpu',d['n_gpus'])
e=d['e2e']; print('e2e',e['value'],{k:(v['value'],v['ms'],v['h2d_gbs_per_rank']) for k,v in e['modes'].items()}); print(e['plain_copy_ceiling'], e['vs_plain_copy'], e['numa'])
print('h16',d['hits16']['value'])
PY
nvidia-smi topo -m | head -8
