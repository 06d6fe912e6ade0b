#!/bin/bash
mkdir -p gpurun_out
echo "== default library"; timeout 300 python scripts/exp/sweep_pfb.py 27 2>&1 | tail -12
echo "== round-1 kernels (RCB_PFB_VARIANT=20)"; RCB_LIBRARY=$PWD/radiocapture_rf_b200/libb200chan_exp.so RCB_PFB_VARIANT=20 timeout 300 python scripts/exp/sweep_pfb.py 27 2>&1 | tail -12
echo "== cluster kernel for 1 tap (VARIANT=21)"; RCB_LIBRARY=$PWD/radiocapture_rf_b200/libb200chan_exp.so RCB_PFB_VARIANT=21 timeout 100 python scripts/exp/sweep_pfb.py 27 1024,1 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_pfb.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -x 2>&1 | tail -8
timeout 200 python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 3 2>gpurun_out/bench_cfg3.err | tee gpurun_out/bench_cfg3.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], round(d['value']), round(d['roofline']['frac'],3), d['e2e'])"
