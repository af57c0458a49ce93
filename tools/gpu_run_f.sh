#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err
tail -3 gpurun_out/r2f_bench_default.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2f_bench_reference.json 2>> gpurun_out/r2f_bench_default.err
python -c "
import json
d = json.load(open('gpurun_out/r2f_bench_default.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('e2e', d['e2e']); print('roofline', d['roofline']); print('cpu', d.get('cpu_baseline')); print('dfma', d['dfma_probe'])
for k, v in d.get('workloads', {}).items(): print(k, {a: v[a] for a in v if a != 'roofline'}); print('   ', v.get('roofline'))
print(open('gpurun_out/r2f_bench_reference.json').read())
"
