#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pfb.py -q -m gpu --tb=short -x > gpurun_out/tests_pfb.log 2>&1; tail -3 gpurun_out/tests_pfb.log
for v in 0; do
  echo "== variant $v"; RCB_PFB_VARIANT=$v python bench.py --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 1 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f  clocks %s' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step'], d['clocks']))
"
done
