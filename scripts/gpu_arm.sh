#!/bin/bash
python -m pytest tests/test_gpu_pfb.py -q -m gpu --tb=short 2>&1 | tail -4
run() { python bench.py --workload $1 --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f e2e %.0f Msps' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']))
"; }
for w in cfg2 cfg5; do echo "== $w arm kernel"; run $w; echo "== $w old kernel"; RCB_PFB_VARIANT=9 run $w; done
