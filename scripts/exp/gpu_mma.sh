#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ddc.py -q -m gpu --tb=short -s 2>&1 | grep -v "^  " | tail -25
B="python bench.py --workload ddc64 --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 2 --no-ceiling"
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"]*1000), "us/step", d["gpu_launches"], round(d["e2e"]["value"]))'
timeout 200 $B 2>gpurun_out/bench_ddc64.err | python -c "$P" "ddc64 product" || tail -5 gpurun_out/bench_ddc64.err
for m in 0 1 2 3 4; do
RCB_LIBRARY=$PWD/radiocapture_rf_b200/libb200chan_exp.so RCB_DDC_DBG=$m timeout 200 $B 2>gpurun_out/bench_ddc64.err | python -c "$P" "ddc64 exp dbg=$m" || tail -5 gpurun_out/bench_ddc64.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_ddc64.csv $B > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_launches_ddc64.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i+1; break
ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[start+8:start+20]:
    if len(r)>vi: print(r[ki][:30], r[vi])
PY
