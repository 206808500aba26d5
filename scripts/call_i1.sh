#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_i1.log 2>&1; tail -3 gpurun_out/r02_gputest_i1.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_i1.json 2> gpurun_out/r02_bench_i1.err; tail -c 300 gpurun_out/r02_bench_i1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_i1.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'k1',d['roofline']['kernel_ms'])
e=d['e2e']; print('e2e',e['value'],{k:(v['value'],v['ms']) for k,v in e['modes'].items()}, e['plain_copy_ceiling'], e['vs_plain_copy'], e['numa'])
print('h16',d['hits16']['value'],d['hits16']['e2e']['value'],'parity',d['parity_sample']['ok'],d['hits16']['parity_sample']['ok'])"
