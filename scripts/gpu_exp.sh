#!/bin/bash
python -m pytest tests/test_gpu_pfb.py -q -m gpu --tb=short -x 2>&1 | tail -3
run() { python bench.py --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 1 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step']))
"; }
for v in 0 16; do echo "== RCB_PFB_VARIANT=$v"; RCB_PFB_VARIANT=$v run; done
RCB_PFB_VARIANT=16 python -m pytest tests/test_gpu_pfb.py -q -m gpu --tb=short -x -k "cfg3 or parity" 2>&1 | tail -3
