#!/bin/bash
mkdir -p gpurun_out
run() { timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-also --e2e-steps 1 "$@" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:300]); continue
    print('value %.1f Gsps  frac %.4f  ms/step %.4f' % (d['value']/1e3, d['roofline']['frac'], d['ms_per_step']))
"; }
for dbg in 0 $((15<<12)) $((13<<12)) $((12<<12)) $((2<<8)) $((8<<8)) $(((15<<12)|(2<<8))); do echo "== cfg3 debug=$dbg"; RCB_PFB_DEBUG=$dbg run; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_cols_tma -s 8 -c 1 -f -o gpurun_out/prof_fft_cols_tma python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_fft_cols.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_rows -s 8 -c 1 -f -o gpurun_out/prof_fft_rows python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 > gpurun_out/ncu_fft_rows.log 2>&1
ls -la gpurun_out | grep prof_fft
