#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ddc.py tests/test_gpu_frontend.py tests/test_gpu_pfb.py tests/test_gpu_ingest.py -x -q -m gpu --tb=short 2>&1 | tail -15
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"]*1000), "us/step", round(d["roofline"]["frac"],4), d["gpu_launches"], "e2e", round(d["e2e"]["value"]))'
B="python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 2 --no-ceiling"
timeout 200 $B 2>gpurun_out/bench_cfg1.err | python -c "$P" "cfg1 lone kernel v2" || tail -5 gpurun_out/bench_cfg1.err
timeout 400 python bench.py --no-cpu 2>gpurun_out/bench_default.err > gpurun_out/r02_bench_default_b.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default_b.json'))
print('cfg3', round(d['value']), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])
a=d['also']
for k,v in a.items():
    if isinstance(v,dict) and 'value' in v: print(' ',k, round(v['value']), {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('roofline_frac','cuda_core_value','tensor_roofline_frac','algorithmic_tflops','e2e','e2e_u8')})
print(' clocks', d['clocks'])
PY
