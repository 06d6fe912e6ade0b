#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pfb.py tests/test_gpu_frontend.py tests/test_gpu_abi_errors.py -q -m gpu --tb=short > gpurun_out/tests_pfb.log 2>&1; tail -15 gpurun_out/tests_pfb.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 "$@" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f e2e %.0f' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']))
"; }
echo "== cfg3"; run
echo "== cfg3_p16"; run --workload cfg3_p16
echo "== cfg2 (64 ch, 2 taps/arm, IQ) new kernel"; run --workload cfg2
echo "== cfg2 old kernel"; RCB_PFB_VARIANT=9 run --workload cfg2
echo "== cfg3_iqfm_p16 new"; run --workload cfg3_iqfm_p16
echo "== cfg3_iqfm_p16 old"; RCB_PFB_VARIANT=9 run --workload cfg3_iqfm_p16 --steps 4
echo "== cfg2_p16_iqfm new"; run --workload cfg2_p16_iqfm
echo "== cfg2_p16_iqfm old"; RCB_PFB_VARIANT=9 run --workload cfg2_p16_iqfm --steps 4
echo "== cfg5"; run --workload cfg5
