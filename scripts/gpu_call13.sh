#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pfb.py -q -m gpu --tb=short -x -k "multi_tap" > gpurun_out/tests_13.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/tests_13.log
run() { timeout 180 python bench.py --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 "$@" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f e2e %.0f' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']))
"; }
echo "== cfg3_p16 warp-specialised"; run --workload cfg3_p16
echo "== cfg3_p16 W=16 phase-serial"; RCB_PFB_VARIANT=16 run --workload cfg3_p16
echo "== cfg3 p8 (8 taps/arm)"; run --workload cfg3_p8

timeout 300 ncu --set full --clock-control none --import-source on -k regex:pfb_fm_ws -s 3 -c 1 -f -o gpurun_out/prof_pfb_ws python bench.py --workload cfg3_p16 --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_full_ws.log 2>&1
ls -la gpurun_out | grep ws
