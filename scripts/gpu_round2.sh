#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short > gpurun_out/tests.log 2>&1; tail -3 gpurun_out/tests.log
for w in cfg4 cfg4_16k; do python bench.py --workload $w --steps 10 --warmup 3 --e2e-steps 2 > gpurun_out/bench_$w.json 2>gpurun_out/bench_$w.err; cut -c1-900 gpurun_out/bench_$w.json; tail -2 gpurun_out/bench_$w.err; done
ncu --set full --clock-control none --import-source on -k regex:fft_ -s 9 -c 3 -f -o gpurun_out/prof_fft python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_fft.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > /dev/null 2>&1
ls -la gpurun_out | tail -8
