#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -25
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
