#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_pfb.py -q -m gpu --tb=short -k "raw or more_than_16 or blocked or host_pipeline or ceiling" 2>&1 | tail -8
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; cut -c1-400 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
