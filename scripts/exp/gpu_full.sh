#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 400 python bench.py 2>gpurun_out/bench_default.err > gpurun_out/r02_bench_default.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
print('cfg3', round(d['value']), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']))
a=d['also']
for k,v in a.items():
    if isinstance(v,dict) and 'value' in v: print(' ',k, round(v['value']), {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('roofline_frac','cuda_core_value','tensor_roofline_frac','algorithmic_tflops')})
print(' pcie', d['e2e'].get('pcie_ceiling'))
print(' clocks', d['clocks'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null > gpurun_out/r02_bench_reference.json; cut -c1-400 gpurun_out/r02_bench_reference.json
timeout 300 python bench.py --workload ddc64 --steps 20 --warmup 3 2>/dev/null > gpurun_out/r02_bench_ddc64.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_ddc64.json')); print('ddc64', round(d['value']), d['roofline'], 'e2e', d['e2e'])"
