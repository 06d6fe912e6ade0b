#!/bin/bash
mkdir -p gpurun_out
for c in 256,1,64 256,16,64 1024,1,64 1024,16,96; do
  timeout 120 python scripts/exp/cl_check.py $c 2>&1 | tail -4
done
timeout 200 python scripts/exp/cl_check.py 2>&1 | tail -14
timeout 600 python -m pytest tests/test_gpu_pfb.py tests/test_gpu_fullsize.py -q -m gpu -x --tb=short 2>&1 | tail -8
for w in cfg3 cfg3_p16 cfg5; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 2>gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], d['value'], d['roofline'])"
done
