#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fft.py tests/test_gpu_frontend.py -q -m gpu --tb=short > gpurun_out/tests_8.log 2>&1; tail -15 gpurun_out/tests_8.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 "$@" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f e2e %.0f' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']))
"; }
echo "== cfg4 TMA cols"; run --workload cfg4
echo "== cfg4 TMA cols, old rows"; RCB_FFT_VARIANT=2 run --workload cfg4; echo "== cfg4 old cols"; RCB_FFT_VARIANT=1 run --workload cfg4
echo "== cfg4_16k TMA cols"; run --workload cfg4_16k
echo "== cfg4_16k old cols"; RCB_FFT_VARIANT=1 run --workload cfg4_16k
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 12 --csv --log-file gpurun_out/launches_cfg4_v3.csv python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_l4.log 2>&1
grep -E "fft_|fold" gpurun_out/launches_cfg4_v3.csv | awk -F, '{print $5, $NF}' | head -12
