#!/bin/bash
# one ncu --set full capture per kernel (north_star: "an ncu capture is committed per kernel") + the launch list of the
# default bench command.  Small blocks (--log2n) keep the ~40 replays per launch short.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling"
cap() {  # name, kernel regex, extra bench args
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o gpurun_out/r02_$1 $B $3 > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap pfb_fm1        pfb_fm1_kernel          "--workload cfg3 --log2n 26"
cap pfb_cl_p8      pfb_cl_kernel           "--workload cfg3_p8 --log2n 26"
cap pfb_ws_p16     pfb_fm_ws_kernel        "--workload cfg3_p16 --log2n 26"
cap pfb_tma_cfg5   pfb_fm_tma_multi        "--workload cfg5"
cap pfb_tma_cfg2   pfb_fm_tma_kernel       "--workload cfg2 --log2n 26"
cap pfb_ws_iqfm    pfb_fm_ws_kernel        "--workload cfg3_iqfm_p16 --log2n 26"
cap ddc_lone_cfg1  ddc_lone_kernel         "--workload cfg1"
cap ddc_tile_64    ddc_tile_kernel         "--workload ddc64 --log2n 22 --no-tensor-cores"
cap ddc_mma2       ddc_mma2_kernel         "--workload ddc64"
cap ddc_mma        "ddc_mma_kernel"        "--workload ddc64 --ddc-mode 2"
cap ddc_head       ddc_head_kernel         "--workload ddc64"
cap ddc_post       ddc_post_kernel         "--workload cfg1"
cap fft_cols       fft_cols_tma_kernel     "--workload cfg4 --log2n 25"
cap fft_rows       fft_rows_kernel         "--workload cfg4 --log2n 25"
cap fft_fold       fft_fold_kernel         "--workload cfg4 --log2n 25"
cap fft_frame      fft_frame_kernel        "--workload cfg4_16k"
cap fft_fold_groups fft_fold_groups_kernel "--workload cfg4_16k"
cap arm_fir_p256   pfb_arm_fir_kernel      "--workload cfg3_p256 --log2n 24"
# K4 / K5 / K6 through small driver scripts
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"quad_demod_rows|convert_iq|post_p25|post_fir_rat|post_fm_deemph|post_squelch" -c 8 -o gpurun_out/r02_k456 python scripts/exp/k456_driver.py > gpurun_out/ncu_k456.log 2>&1; tail -1 gpurun_out/ncu_k456.log
# launch list of the bench command (every launch with its device time; shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_cfg3.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling > gpurun_out/ncu_launches.log 2>&1
grep -c pfb_fm1 gpurun_out/r02_launches_cfg3.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_ddc64.csv python bench.py --workload ddc64 --steps 3 --warmup 3 --no-cpu --no-also --e2e-steps 1 --no-ceiling > /dev/null 2>&1
grep -c ddc_mma2 gpurun_out/r02_launches_ddc64.csv
