#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 "$@" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step']))
"; }
for dbg in 0 4 8 16 12 28; do echo "== cfg3_p16 W=8 debug=$dbg"; RCB_PFB_DEBUG=$dbg run --workload cfg3_p16; done
for dbg in 0 28; do echo "== cfg3_p16 W=16 debug=$dbg"; RCB_PFB_VARIANT=16 RCB_PFB_DEBUG=$dbg run --workload cfg3_p16; done
for dbg in 0 28; do echo "== cfg5 debug=$dbg"; RCB_PFB_DEBUG=$dbg run --workload cfg5; done
RCB_PFB_VARIANT=16 python -m pytest tests/test_gpu_pfb.py -q -m gpu --tb=short -k "multi_tap" 2>&1 | tail -3
