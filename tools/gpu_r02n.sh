#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02n_tests.log
cat gpurun_out/r02n_tests.log
bash tools/gpu_r02k.sh 2>&1 | grep -A1 "== n=172\|== n=344\|== bump\|== n=128"
