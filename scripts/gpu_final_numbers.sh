#!/bin/bash
# record one bench line per workload (profiles/r01_bench_<workload>_v6.json) + the reference arm
mkdir -p gpurun_out
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref.json
for wl in cfg3_p16 cfg3_p8 cfg5 cfg2 cfg3_iqfm_p16 cfg4 cfg4_16k cfg1 ddc64; do
  timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --e2e-steps 2 > gpurun_out/bench_$wl.json 2>/dev/null
  python - "$wl" <<'PY'
import json, sys
wl = sys.argv[1]
try:
    d = json.load(open('gpurun_out/bench_%s.json' % wl))
    cb = d.get('cpu_baseline') or {}
    print('%-14s value %9.1f Msps  frac %.4f  e2e %8.1f  cpu %8.1f (%s cores)' % (wl, d['value'], d['roofline']['frac'], d['e2e']['value'], cb.get('value', 0), cb.get('cores')))
except Exception as e:
    print(wl, 'failed', e)
PY
done
