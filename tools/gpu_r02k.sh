#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02k_lusgs.log; : > $L
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lusgs" 2>&1 | tail -3
for n in 64 128 172 344; do
  echo "== n=$n" >> $L
  timeout 900 python tools/lusgs_time.py $n 2>&1 | grep -v "^cells" >> $L
done
echo "== bump" >> $L
timeout 900 python tools/lusgs_time.py bump 1280 1040 2>&1 | grep -v "^cells" >> $L
echo "== prof 172" >> $L
ICSB200_LUSGS_PROF=1 timeout 900 python tools/lusgs_time.py 172 2>&1 | grep -v "^cells" >> $L
cat $L
