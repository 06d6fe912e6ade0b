#!/bin/bash
# one GPU visit at the end of round 2: all GPU tests, smoke, both bench arms, one bench line per workload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short > gpurun_out/r02_gpu_tests_final.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gpu_tests_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/r02_bench_reference.json
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/bench_default.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
print('cfg3', round(d['value']), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']))
a=d['also']
for k,v in a.items():
    if isinstance(v,dict) and 'value' in v: print(' ',k, round(v['value']), {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('roofline_frac','cuda_core_value','tensor_roofline_frac','algorithmic_tflops')}, {kk: round(v[kk]['value']) for kk in ('e2e','e2e_u8') if kk in v})
print(' pcie', d['e2e'].get('pcie_ceiling'))
print(' clocks', d['clocks'])
PY
for wl in cfg1 cfg2 cfg4 cfg4_16k cfg5 ddc64 cfg3_p16 cfg3_p8; do
  timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --e2e-steps 2 > gpurun_out/r02_bench_$wl.json 2>/dev/null
  python - "$wl" <<'PY'
import json, sys
wl = sys.argv[1]
try:
    d = json.load(open('gpurun_out/r02_bench_%s.json' % wl))
    cb = d.get('cpu_baseline') or {}
    print('%-14s value %9.1f Msps  frac %.4f  %s  e2e %8.1f  cpu %8.1f (%s cores) launches %s' % (wl, d['value'], d['roofline']['frac'], d['roofline'].get('kernel'), d['e2e']['value'], cb.get('value', 0), cb.get('cores'), d['gpu_launches']))
except Exception as e:
    print(wl, 'failed', e)
PY
done
