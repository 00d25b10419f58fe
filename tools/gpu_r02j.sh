#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02j_lusgs.log; : > $L
for n in 128 344; do
  echo "== n=$n" >> $L
  ICSB200_LUSGS_PROF=1 timeout 900 python tools/lusgs_time.py $n 2>&1 | grep -v "^cells" >> $L
done
echo "== bump" >> $L
ICSB200_LUSGS_PROF=1 timeout 900 python tools/lusgs_time.py bump 1280 1040 2>&1 | grep -v "^cells" >> $L
cat $L
