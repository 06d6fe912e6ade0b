#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
./scripts/micro/f32x2.bin > gpurun_out/micro_f32x2.txt 2>&1; cat gpurun_out/micro_f32x2.txt
python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/tests.log 2>&1; tail -5 gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
