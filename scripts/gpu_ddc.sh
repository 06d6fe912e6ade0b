#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ddc.py tests/test_gpu_frontend.py -q -m gpu --tb=short 2>&1 | tail -6
for w in cfg1 ddc64; do python bench.py --workload $w --steps 10 --warmup 3 --e2e-steps 2 --no-cpu > gpurun_out/bench_$w.json 2>gpurun_out/bench_$w.err; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$w.json'))
print('$w value %.1f Msps frac %.4f ms/step %.3f e2e %.1f Msps launches %s' % (d['value'], d['roofline']['frac'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
"; tail -2 gpurun_out/bench_$w.err; done
