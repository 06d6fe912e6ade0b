#!/bin/bash
mkdir -p gpurun_out
timeout 100 python scripts/exp/cl_check.py 256,16,64 1024,16,96 1024,1,64 1024,1,3000 1024,0.25,8 1024,0.25,5 2>&1 | tail -8
timeout 500 python -m pytest tests/test_gpu_pfb.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -x 2>&1 | tail -12
for w in cfg3 cfg3_p16 cfg3_p8 cfg5; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 2>gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], round(d['value']), round(d['roofline']['frac'],3))"
done
timeout 200 python bench.py --workload cfg3 --out-block 0 --steps 10 --warmup 3 --no-cpu --no-also --e2e-steps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('plain', d['config']['workload'][:40], round(d['value']), round(d['roofline']['frac'],3))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pfb_cl -s 3 -c 1 -o gpurun_out/prof_cl4_p16 python bench.py --workload cfg3_p16 --log2n 26 --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_cl_p16.log 2>&1; tail -2 gpurun_out/ncu_cl_p16.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pfb_fm1 -s 3 -c 1 -o gpurun_out/prof_fm1 python bench.py --workload cfg3 --log2n 26 --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_fm1.log 2>&1; tail -2 gpurun_out/ncu_fm1.log
