#!/bin/bash
# first GPU contact of ddc_mma_kernel: parity tests under a short timeout (a hang must not cost the box), then the rate
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ddc.py -q -m gpu --tb=short -s -k "tensor_core" 2>&1 | tail -40
for f in "" "--no-tensor-cores"; do
timeout 200 python bench.py --workload ddc64 --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 2 --no-ceiling $f 2>gpurun_out/bench_ddc64.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ddc64 $f', round(d['value']), round(d['roofline']['frac'],4), d['gpu_launches'], round(d['e2e']['value']))" || tail -5 gpurun_out/bench_ddc64.err
done
