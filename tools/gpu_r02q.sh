#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lusgs" 2>&1 | tail -4
L=gpurun_out/r02q_lusgs.log; : > $L
for n in 64 128 172 200; do
  echo "== n=$n" >> $L
  timeout 600 python tools/lusgs_time.py $n 2>&1 | grep -v "^cells" >> $L
  echo "== n=$n colmode 0" >> $L
  ICSB200_LUSGS_COLMODE=0 timeout 600 python tools/lusgs_time.py $n 2>&1 | grep -v "^cells" >> $L
done
echo "== bump" >> $L
timeout 900 python tools/lusgs_time.py bump 1280 1040 2>&1 | grep -v "^cells" >> $L
echo "== bump colmode 0" >> $L
ICSB200_LUSGS_COLMODE=0 timeout 900 python tools/lusgs_time.py bump 1280 1040 2>&1 | grep -v "^cells" >> $L
cat $L
