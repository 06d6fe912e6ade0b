#!/bin/bash
# compute-sanitizer memcheck over small parity tests of the kernels added late in round 2: fft_frame_kernel (bulk-copied
# rows, carry / group paths), ddc_lone_kernel (one bulk copy per frame, odd / even decimations), tensor-core DDC bank,
# wire-format DDC input
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -q -m gpu -x --tb=short \
  -k "(frame_resident_kernel_16384 and 7-5-3) or (lone_channel_kernel_shapes and (96-349 or 9-41 or 33-100)) or (wire_format and u8 and 1-False) or (tensor_core_bank_matches and 1-0)" \
  > gpurun_out/r02_sanitize_memcheck_b.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitize_memcheck_b.txt
