#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pfb.py -q -m gpu --tb=short > gpurun_out/tests_pfb.log 2>&1; tail -12 gpurun_out/tests_pfb.log
run() { python bench.py --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 1 "$@" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f  clocks %s' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step'], d['clocks']))
"; }
echo "== cfg3 packed W=8"; run
echo "== cfg3 packed W=16 paired-lane stores"; RCB_PFB_VARIANT=16 run
echo "== cfg3_p16 (1024 ch, 16 taps/arm) new FIR-phase kernel"; run --workload cfg3_p16
echo "== cfg3_p16 old kernel (variant 9)"; RCB_PFB_VARIANT=9 run --workload cfg3_p16 --steps 5
echo "== cfg5 (8 x 256 ch, 16 taps/arm)"; run --workload cfg5
ncu --set full --clock-control none --import-source on -k regex:pfb_fm_tma -s 3 -c 1 -f -o gpurun_out/prof_pfb_p16 python bench.py --workload cfg3_p16 --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_full_p16.log 2>&1
ls -la gpurun_out | tail -4
