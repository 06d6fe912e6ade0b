#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ddc.py tests/test_gpu_frontend.py -q -m gpu --tb=short 2>&1 | tail -5
B="python bench.py --workload ddc64 --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 2 --no-ceiling"
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"]*1000), "us/step", d["gpu_launches"], "e2e", round(d["e2e"]["value"]))'
timeout 200 $B 2>gpurun_out/bench_ddc64.err | python -c "$P" "ddc64 mode 1" || tail -5 gpurun_out/bench_ddc64.err
timeout 200 $B --log2n 22 2>gpurun_out/bench_ddc64.err | python -c "$P" "ddc64 mode 1 2^22" || tail -5 gpurun_out/bench_ddc64.err
timeout 200 $B --log2n 22 --no-tensor-cores 2>gpurun_out/bench_ddc64.err | python -c "$P" "ddc64 cuda cores 2^22" || tail -5 gpurun_out/bench_ddc64.err
timeout 200 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 2 --no-ceiling 2>gpurun_out/bench_ddc64.err | python -c "$P" "cfg1" || tail -5 gpurun_out/bench_ddc64.err
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
