#!/bin/bash
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]))'
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling"
for v in main t1 t2 t3 main t1; do
  if [ $v = main ]; then L=$PWD/radiocapture_rf_b200/libb200chan.so; else L=$PWD/radiocapture_rf_b200/libb200chan_$v.so; fi
  RCB_LIBRARY=$L timeout 120 $B 2>/dev/null | python -c "$P" "cfg3 $v"
done
timeout 200 python bench.py --workload cfg4_16k --steps 10 --warmup 3 --e2e-steps 2 > gpurun_out/r02_bench_cfg4_16k.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg4_16k.json')); print('cfg4_16k', round(d['value']), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['e2e'].get('pcie_ceiling',{}).get('frac'))"
bash scripts/gpu_sanitize_r02b.sh
