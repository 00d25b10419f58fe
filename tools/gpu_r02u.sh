#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_tests_gpu.log; cat gpurun_out/r02_tests_gpu.log
timeout 300 python tools/diag_bump_parts.py 8 2>&1 | tail -8 | cut -c1-200
for n in 172; do timeout 300 python tools/lusgs_time.py $n 2>&1 | tail -1; done
timeout 300 python tools/lusgs_time.py bump 1280 1040 2>&1 | tail -1
