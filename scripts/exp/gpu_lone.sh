#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ddc.py tests/test_gpu_frontend.py -q -m gpu --tb=short 2>&1 | tail -8
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"]*1000), "us/step", round(d["roofline"]["frac"],4), d["gpu_launches"], "e2e", round(d["e2e"]["value"]))'
B="python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 2 --no-ceiling"
timeout 200 $B 2>gpurun_out/bench_cfg1.err | python -c "$P" "cfg1 lone kernel" || tail -5 gpurun_out/bench_cfg1.err
RCB_LIBRARY=$PWD/radiocapture_rf_b200/libb200chan_exp.so RCB_DDC_LONE=0 timeout 200 $B 2>gpurun_out/bench_cfg1.err | python -c "$P" "cfg1 tile kernel" || tail -5 gpurun_out/bench_cfg1.err
