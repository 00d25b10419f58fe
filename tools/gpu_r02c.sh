#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02c_lusgs_times.log; : > $L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lusgs or bitwise or linearity" 2>&1 | tail -15 > gpurun_out/r02c_tests.log
cat gpurun_out/r02c_tests.log
for n in 64 128 172 344; do
  echo "== n=$n mode=auto" >> $L
  ICSB200_LUSGS_PROF=1 timeout 900 python tools/lusgs_time.py $n >> $L 2>&1
done
echo "== bump 1280x1040 mode=auto" >> $L
ICSB200_LUSGS_PROF=1 timeout 600 python tools/lusgs_time.py bump 1280 1040 >> $L 2>&1
cat $L
