#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over small parity cases of every new kernel variant
mkdir -p gpurun_out
K="test_pfb_cfg3_literal or (multi_tap_fast_kernel and (1024-16-96 or 256-12 or 64-8)) or (blocked_device_output and (1024-0.25-2-8 or 256-4)) or (iq_fm_parity and (1024-4 or 64-2))"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pfb.py -q -m gpu -x -k "$K" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pfb.py -q -m gpu -x -k "test_pfb_cfg3_literal or (multi_tap_fast_kernel and (1024-3-131 or 64-8)) or (iq_fm_parity and 64-2)" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitize_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fft.py tests/test_gpu_ddc.py -q -m gpu -x > gpurun_out/sanitize_memcheck_fft_ddc.log 2>&1; echo "memcheck fft/ddc rc=$?"; tail -4 gpurun_out/sanitize_memcheck_fft_ddc.log
