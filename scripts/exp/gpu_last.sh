#!/bin/bash
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_pfb.py tests/test_gpu_ingest.py tests/test_gpu_fullsize.py -x -q -m gpu --tb=short -k "1024 or fullsize or full_size or raw" 2>&1 | tail -4
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]))'
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling"
timeout 100 $B 2>/dev/null | python -c "$P" "cfg3 final a"
timeout 100 $B 2>/dev/null | python -c "$P" "cfg3 final b"
