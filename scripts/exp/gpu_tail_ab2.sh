#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_frontend.py -x -q -m gpu --tb=short 2>&1 | tail -4
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]))'
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling"
for v in t4 t5 main; do
  if [ $v = main ]; then L=$PWD/radiocapture_rf_b200/libb200chan.so; else L=$PWD/radiocapture_rf_b200/libb200chan_$v.so; fi
  RCB_LIBRARY=$L timeout 120 $B 2>/dev/null | python -c "$P" "cfg3 $v"
done
