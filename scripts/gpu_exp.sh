#!/bin/bash
timeout 60 python -m pytest tests/test_gpu_fft.py tests/test_gpu_pfb.py -q -m gpu -x -k "(logpow_parity and (4096 or 16384)) or (blocked_device_output and 1024-0.25-2-8) or streaming_split" 2>&1 | tail -2
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
