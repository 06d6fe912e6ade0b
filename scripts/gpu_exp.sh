#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_frontend.py -q -m gpu --tb=short -x > gpurun_out/tests_frontend.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/tests_frontend.log
