#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_pfb.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -x > gpurun_out/tests_ws_iq.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/tests_ws_iq.log
run() { timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 "$@" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step']))
"; }
echo "== cfg3_iqfm_p16 ws"; run --workload cfg3_iqfm_p16
echo "== cfg3_iqfm_p16 phase-serial"; RCB_PFB_VARIANT=8 run --workload cfg3_iqfm_p16
