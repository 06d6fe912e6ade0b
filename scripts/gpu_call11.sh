#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 "$@" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f e2e %.0f' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']))
"; }
for wl in cfg2 cfg5 cfg2_p16_iqfm; do echo "== $wl base"; run --workload $wl; done
cp radiocapture_rf_b200/libb200chan.so /tmp/base.so
cp radiocapture_rf_b200/libb200chan_occ.so radiocapture_rf_b200/libb200chan.so
for wl in cfg2 cfg5 cfg2_p16_iqfm; do echo "== $wl occupancy experiment (3-4 CTAs/SM)"; run --workload $wl; done
python -m pytest tests/test_gpu_pfb.py -q -m gpu --tb=short -x 2>&1 | tail -3
cp /tmp/base.so radiocapture_rf_b200/libb200chan.so
