#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pfb.py tests/test_gpu_fullsize.py tests/test_gpu_ingest.py -q -m gpu --tb=short 2>&1 | tail -12
for i in 1 2; do
timeout 200 python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 2 --no-ceiling 2>gpurun_out/bench_cfg3.err | tee gpurun_out/bench_cfg3.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], round(d['value']), round(d['roofline']['frac'],4), d['gpu_launches'], round(d['e2e']['value']))"
done
timeout 200 python bench.py --workload cfg3 --out-block 0 --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 2 --no-ceiling 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('plain', round(d['value']), round(d['roofline']['frac'],4), d['gpu_launches'], round(d['e2e']['value']))"
