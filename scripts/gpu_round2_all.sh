#!/bin/bash
# the last GPU visit of round 2: tests + smoke + bench arms + one line per workload, then the per-kernel captures
bash scripts/gpu_round2_final.sh
bash scripts/gpu_profiles_r02.sh
du -sh gpurun_out
