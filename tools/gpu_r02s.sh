#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_tests_gpu.log; cat gpurun_out/r02_tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
L=gpurun_out/r02_lusgs_times.log; : > $L
for n in 64 128 172 200 344; do
  echo "== onera box n=$n" >> $L
  timeout 900 python tools/lusgs_time.py $n 2>&1 >> $L
done
echo "== bump 3x1280x1040" >> $L
timeout 900 python tools/lusgs_time.py bump 1280 1040 2>&1 >> $L
for n in 172 344; do
  echo "== in-kernel cycle profile, n=$n (ICSB200_LUSGS_PROF=1: the instrumented instantiation is ~20 % slower)" >> $L
  ICSB200_LUSGS_PROF=1 timeout 900 python tools/lusgs_time.py $n 2>&1 | grep -v "^cells" >> $L
done
T=gpurun_out/r02_lusgs_trace.log; : > $T
for n in 172 344; do
  echo "== n=$n" >> $T
  timeout 900 python tools/lusgs_blk_trace.py $n >> $T 2>&1
done
grep -A2 "== onera\|== bump" $L | grep "lusgs"
