#!/bin/bash
mkdir -p gpurun_out
for f in "" "--no-multi"; do
  timeout 200 python bench.py --workload cfg5 $f --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg5 $f', round(d['value']), round(d['roofline']['frac'],3), d['gpu_launches'])"
done
for f in "" "--no-multi"; do
  timeout 200 python bench.py --workload cfg5 $f --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg5 $f', round(d['value']), round(d['roofline']['frac'],3), d['gpu_launches'])"
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash scripts/gpu_sanitize_r02.sh
