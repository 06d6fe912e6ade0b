#!/bin/bash
# compute-sanitizer memcheck over small parity tests of the round-2 kernels (cluster / DSMEM, TMA stores, persistent scan,
# post-demod, fused ingest, multi-stream launch)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -q -m gpu -x --tb=short \
  -k "(cluster_kernel_iteration_edges and (1024-16-17 or 1024-1-33 or 256-16-31)) or (cluster_kernel_blocked_layouts and 1024-16-16) or (persistent_pipeline and 4096) or (p25_c4fm_front and 1-300) or squelch_gate or (raw_input_matches and 1024-0.25-2-u8) or (multi_stream and 256-4-3-64) or (more_than_16 and 1024-24) or pull_all" \
  > gpurun_out/sanitize_memcheck_r02.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/sanitize_memcheck_r02.log
grep -c "ERROR SUMMARY" gpurun_out/sanitize_memcheck_r02.log; grep "ERROR SUMMARY" gpurun_out/sanitize_memcheck_r02.log | sort | uniq -c | head
