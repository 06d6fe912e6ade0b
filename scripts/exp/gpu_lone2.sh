#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ddc_lone_kernel -s 3 -c 1 -o gpurun_out/r02_ddc_lone_v1 $B > gpurun_out/ncu_ddc_lone.log 2>&1; tail -1 gpurun_out/ncu_ddc_lone.log
