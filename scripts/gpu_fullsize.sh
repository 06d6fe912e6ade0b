#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu --tb=short > gpurun_out/tests_fullsize.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/tests_fullsize.log
