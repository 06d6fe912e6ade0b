#!/bin/bash
# one GPU visit: tests, smoke, bench, launch list, full ncu capture of the top kernel
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short > gpurun_out/tests.log 2>&1; tail -3 gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>>gpurun_out/bench.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pfb_fm_kernel -s 3 -c 2 -f -o gpurun_out/prof_pfb python bench.py --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -12
