#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ddc.py -q -m gpu --tb=short -k tensor_core 2>&1 | tail -3
B="python bench.py --workload ddc64 --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling"
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"]*1000), "us/step", d["gpu_launches"])'
timeout 200 $B 2>gpurun_out/bench_ddc64.err | python -c "$P" "ddc64 product" || tail -5 gpurun_out/bench_ddc64.err
for v in 1000 40 20; do
RCB_LIBRARY=$PWD/radiocapture_rf_b200/libb200chan_exp.so RCB_DDC_KQ=$v timeout 200 $B 2>gpurun_out/bench_ddc64.err | python -c "$P" "ddc64 exp kq=$v" || tail -5 gpurun_out/bench_ddc64.err
done
RCB_LIBRARY=$PWD/radiocapture_rf_b200/libb200chan_exp.so RCB_DDC_DBG=3 timeout 200 $B 2>gpurun_out/bench_ddc64.err | python -c "$P" "ddc64 exp dbg=3 (TMA only)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ddc_mma_kernel -s 3 -c 1 -o gpurun_out/r02_ddc_mma_v3 $B > gpurun_out/ncu_ddc_mma.log 2>&1; tail -1 gpurun_out/ncu_ddc_mma.log
