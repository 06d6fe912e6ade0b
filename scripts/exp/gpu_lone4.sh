#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ddc.py tests/test_gpu_fft.py tests/test_gpu_frontend.py -x -q -m gpu --tb=short 2>&1 | tail -12
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"]*1000), "us/step", round(d["roofline"]["frac"],4), d["gpu_launches"], "e2e", round(d["e2e"]["value"]))'
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 2 --no-ceiling"
timeout 200 $B --workload cfg1 2>gpurun_out/bench_cfg1.err | python -c "$P" "cfg1 lone v3" || tail -5 gpurun_out/bench_cfg1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ddc_lone_kernel -s 3 -c 1 -o gpurun_out/r02_ddc_lone_v3 $B --workload cfg1 > gpurun_out/ncu_ddc_lone.log 2>&1; tail -1 gpurun_out/ncu_ddc_lone.log
