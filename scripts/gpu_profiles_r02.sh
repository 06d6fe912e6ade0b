#!/bin/bash
# one ncu --set full capture per kernel (north_star: "an ncu capture is committed per kernel") + the launch lists of the
# bench commands.  Small blocks (--log2n) keep the ~40 replays per launch short.  The reports (~29 MB each) are
# summarised ON the box (scripts/ncu_summary.py, scripts/make_traffic.py) and deleted: gpurun brings back at most 64 MiB.
mkdir -p gpurun_out/prof
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling"
cap() {  # name, kernel regex, extra bench args, input samples of the captured launch
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o gpurun_out/r02_$1 $B $3 > gpurun_out/ncu_$1.log 2>&1
  if [ -f gpurun_out/r02_$1.ncu-rep ]; then
    python scripts/ncu_summary.py gpurun_out/r02_$1.ncu-rep $4 > gpurun_out/prof/r02_$1_summary.txt 2>/dev/null
    head -2 gpurun_out/prof/r02_$1_summary.txt | tr '\n' ' '; echo
  else
    echo "no capture for $1"; tail -2 gpurun_out/ncu_$1.log
  fi
}
cap pfb_fm1        pfb_fm1_kernel          "--workload cfg3 --log2n 26"           67108864
cap pfb_ws_p16     pfb_fm_ws_kernel        "--workload cfg3_p16 --log2n 26"       67108864
cap pfb_tma_cfg2   pfb_fm_tma_kernel       "--workload cfg2 --log2n 26"           67108864
cap pfb_ws_iqfm    pfb_fm_ws_kernel        "--workload cfg3_iqfm_p16 --log2n 26"  67108864
cap ddc_lone_cfg1  ddc_lone_kernel         "--workload cfg1"                      16777216
cap ddc_tile_64    ddc_tile_kernel         "--workload ddc64 --log2n 22 --no-tensor-cores" 4194304
cap ddc_head       ddc_head_kernel         "--workload ddc64"                     16777216
cap ddc_post       ddc_post_kernel         "--workload cfg1"                      16777216
cap fft_cols       fft_cols_tma_kernel     "--workload cfg4 --log2n 25"           4194304
cap fft_rows       fft_rows_kernel         "--workload cfg4 --log2n 25"           4194304
cap fft_fold       fft_fold_kernel         "--workload cfg4 --log2n 25"           4194304
cap fft_frame      fft_frame_kernel        "--workload cfg4_16k"                  134217728
cap fft_fold_groups fft_fold_groups_kernel "--workload cfg4_16k"                  134217728
cap arm_fir_p256   pfb_arm_fir_kernel      "--workload cfg3_p256 --log2n 24"      16777216
python scripts/make_traffic.py gpurun_out gpurun_out/prof/traffic.json > gpurun_out/prof/traffic.log 2>&1; tail -9 gpurun_out/prof/traffic.log
rm -f gpurun_out/r02_*.ncu-rep
# K4 / K5 / K6 through a small driver script: one summary per kernel of the report
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"quad_demod_rows|convert_iq|post_p25|post_fir_rat|post_fm_deemph|post_squelch" -c 8 -o gpurun_out/r02_k456 python scripts/exp/k456_driver.py > gpurun_out/ncu_k456.log 2>&1; tail -1 gpurun_out/ncu_k456.log
ncu -i gpurun_out/r02_k456.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
h = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size']
print(' | '.join(want))
print(' | '.join(rows[1][h.index(w)] for w in want))
for r in rows[2:]:
    print(' | '.join(r[h.index(w)] for w in want))
" > gpurun_out/prof/r02_k456_summary.txt 2>&1
rm -f gpurun_out/r02_k456.ncu-rep
# launch lists of the bench commands (every launch with its device time; shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/prof/r02_launches_cfg3.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling > gpurun_out/ncu_launches.log 2>&1
grep -c pfb_fm1 gpurun_out/prof/r02_launches_cfg3.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/prof/r02_launches_ddc64.csv python bench.py --workload ddc64 --steps 3 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling > /dev/null 2>&1
grep -c ddc_mma2 gpurun_out/prof/r02_launches_ddc64.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/prof/r02_launches_cfg4_16k.csv python bench.py --workload cfg4_16k --steps 3 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling > /dev/null 2>&1
grep -c fft_frame gpurun_out/prof/r02_launches_cfg4_16k.csv
du -sh gpurun_out
