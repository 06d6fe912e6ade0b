#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pfb.py tests/test_gpu_frontend.py -q -m gpu --tb=short -x > gpurun_out/tests_pfb.log 2>&1; tail -5 gpurun_out/tests_pfb.log
run() { python bench.py --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 1 "$@" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f  clocks %s' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step'], d['clocks']))
"; }
echo "== packed (default)"; run
echo "== packed, stores suppressed (experiment)"; RCB_PFB_DEBUG=1 run
echo "== scalar v5 (variant 3)"; RCB_PFB_VARIANT=3 run
echo "== packed cfg2-like N=256 fm (cfg5 uses P=16: generic kernel)"; run --workload cfg5
ncu --set full --clock-control none --import-source on -k regex:pfb_fm_tma -s 3 -c 1 -f -o gpurun_out/prof_pfb_pk python bench.py --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -5
