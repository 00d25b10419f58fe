#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02r_colshape.log; : > $L
for cfg in "8 4 8" "8 4 4" "8 4 2" "8 8 4" "8 8 2" "4 4 8" "4 4 4" "16 4 4" "16 2 4" "8 2 8"; do
  set -- $cfg
  for n in 172; do
    echo "== n=$n bside=$1 bside2=$2 depth=$3 (column mode)" >> $L
    ICSB200_LUSGS_COLMODE=1 ICSB200_LUSGS_BSIDE=$1 ICSB200_LUSGS_BSIDE2=$2 ICSB200_LUSGS_DEPTH=$3 timeout 600 python tools/lusgs_time.py $n 2>&1 | grep -v "^cells" | sed 's/.*lusgs/lusgs/' >> $L
  done
done
echo "== n=344 column mode forced, 8 4 8" >> $L
ICSB200_LUSGS_COLMODE=1 timeout 900 python tools/lusgs_time.py 344 2>&1 | grep -v "^cells" >> $L
for cfg in "16 16" "16 8" "32 8" "8 16" "8 32" "32 4"; do
  set -- $cfg
  echo "== bump bside=$1 depth=$2 (column mode)" >> $L
  ICSB200_LUSGS_BSIDE=$1 ICSB200_LUSGS_DEPTH=$2 timeout 600 python tools/lusgs_time.py bump 1280 1040 2>&1 | grep -v "^cells" | sed 's/.*lusgs/lusgs/' >> $L
done
cat $L
