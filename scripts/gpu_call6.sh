#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pfb.py tests/test_gpu_ddc.py -q -m gpu --tb=short > gpurun_out/tests_6.log 2>&1; tail -8 gpurun_out/tests_6.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 "$@" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f e2e %.0f' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']))
"; }
echo "== cfg3 block 1024"; run
echo "== cfg3 block 8 (OB8 kernel)"; run --out-block 8
echo "== cfg3 block 8, stores suppressed"; RCB_PFB_DEBUG=1 run --out-block 8
echo "== ddc64 packed"; run --workload ddc64
echo "== cfg1"; run --workload cfg1
for mb in 24 48 96; do echo "== cfg4 scratch $mb MB"; RCB_FFT_SCRATCH_MB=$mb run --workload cfg4; done
echo "== cfg4_16k"; run --workload cfg4_16k
