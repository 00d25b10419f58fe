#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02e_trace.log; : > $L
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lusgs" 2>&1 | tail -3
for n in 172 344; do
  echo "== n=$n" >> $L
  timeout 900 python tools/lusgs_blk_trace.py $n >> $L 2>&1
done
cat $L
