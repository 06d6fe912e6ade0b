#!/bin/bash
python -m pytest tests/test_gpu_fft.py tests/test_gpu_abi_errors.py tests/test_gpu_frontend.py -q -m gpu --tb=short 2>&1 | tail -6
run() { python bench.py --workload $1 --steps 10 --warmup 3 --no-cpu --e2e-steps 2 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f e2e %.0f Msps' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']))
"; }
for mb in 48 96; do echo "== cfg4 scratch $mb MB"; RCB_FFT_SCRATCH_MB=$mb run cfg4; done
echo "== cfg4_16k"; run cfg4_16k
