#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02b_lusgs_times.log; : > $L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lusgs or bitwise or linearity" 2>&1 | tail -8
timeout 300 compute-sanitizer --tool racecheck python tools/lusgs_time.py 16 > gpurun_out/r02b_racecheck.log 2>&1
grep -c hazard gpurun_out/r02b_racecheck.log; grep -m6 -B2 -A8 "hazard" gpurun_out/r02b_racecheck.log | head -60
for n in 128 172 344; do
  echo "== n=$n mode=auto" >> $L
  ICSB200_LUSGS_PROF=1 timeout 900 python tools/lusgs_time.py $n >> $L 2>&1
done
echo "== bump 1280x1040 mode=auto" >> $L
ICSB200_LUSGS_PROF=1 timeout 600 python tools/lusgs_time.py bump 1280 1040 >> $L 2>&1
cat $L
