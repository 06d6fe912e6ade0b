#!/bin/bash
# one GPU visit: all GPU tests, smoke, both bench arms, launch list of the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
grep -c pfb_fm_tma gpurun_out/launches.csv
