#!/bin/bash
# memcheck over the newest kernels: warp-specialised multi-tap kernel (FM and IQ+FM) and the TMA column FFT
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pfb.py tests/test_gpu_fft.py -q -m gpu -x -k "(multi_tap_fast_kernel and 1024-16-96) or (iq_fm_parity and 1024-16) or (logpow_parity and 16384)" > gpurun_out/sanitize_memcheck_ws_fft.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_memcheck_ws_fft.log
