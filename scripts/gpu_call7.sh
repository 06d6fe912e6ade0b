#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ddc.py tests/test_gpu_fft.py tests/test_gpu_frontend.py -q -m gpu --tb=short > gpurun_out/tests_7.log 2>&1; tail -5 gpurun_out/tests_7.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 "$@" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f e2e %.0f' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']))
"; }
echo "== cfg1 (lone-channel blocking)"; run --workload cfg1
echo "== ddc64"; run --workload ddc64
echo "== cfg4 packed"; run --workload cfg4
echo "== cfg4_16k packed"; run --workload cfg4_16k
ncu --set full --clock-control none --import-source on -k regex:pfb_fm_tma -s 3 -c 1 -f -o gpurun_out/prof_pfb_p16w16 python bench.py --workload cfg3_p16 --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_full_p16w16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pfb_fm_tma -s 3 -c 1 -f -o gpurun_out/prof_pfb_ob8 python bench.py --out-block 8 --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_full_ob8.log 2>&1
ls -la gpurun_out | tail -5
