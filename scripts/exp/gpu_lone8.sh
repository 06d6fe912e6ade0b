#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ddc.py tests/test_gpu_frontend.py -x -q -m gpu --tb=short 2>&1 | tail -6
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"]*1000), "us/step", round(d["roofline"]["frac"],4), d["gpu_launches"], "e2e", round(d["e2e"]["value"]))'
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 2 --no-ceiling"
timeout 200 $B --workload cfg1 2>gpurun_out/bench_cfg1.err | python -c "$P" "cfg1 lone v7" || tail -5 gpurun_out/bench_cfg1.err
bash scripts/gpu_profiles_r02.sh
