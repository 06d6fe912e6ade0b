#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ddc.py tests/test_gpu_frontend.py -x -q -m gpu --tb=short 2>&1 | tail -15
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"]*1000), "us/step", round(d["roofline"]["frac"],4), d["gpu_launches"], "e2e", round(d["e2e"]["value"]))'
B="python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 2 --no-ceiling"
timeout 200 $B 2>gpurun_out/bench_cfg1.err | python -c "$P" "cfg1 lone kernel v2" || tail -5 gpurun_out/bench_cfg1.err
timeout 100 python - <<'PY'
import bench, json
from bench import *
r = bench.ddc_lone_side_run(0, "cfg1", 10, 3, 1, None, 0, 6551.0)
print(json.dumps(r))
PY
