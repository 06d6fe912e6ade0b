#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pfb.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -k "multi or more_than_16 or fullsize or cfg3" 2>&1 | tail -12
for w in cfg5 cfg3_p256; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling 2>gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], round(d['value']), round(d['roofline']['frac'],3), d['gpu_launches'], round(d['e2e']['value']))"
done
