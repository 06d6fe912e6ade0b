#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ddc.py -q -m gpu --tb=short -k "pull_all or two_engines" 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_scan -s 2 -c 1 -o gpurun_out/prof_scan python bench.py --workload cfg4 --log2n 25 --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling > gpurun_out/ncu_scan.log 2>&1; tail -2 gpurun_out/ncu_scan.log
for w in cfg1 ddc64; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 5 --no-ceiling 2>gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], round(d['value']), round(d['roofline']['frac'],3), d['gpu_launches'], round(d['e2e']['value']))"
done
