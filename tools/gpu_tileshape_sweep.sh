#!/usr/bin/env bash
# LU-SGS block-tile shape sweep (profiles/r02_lusgs_colshape_sweep.log): cross-section x depth on a 172^3 box and the bump-4M mesh
set -u
mkdir -p gpurun_out
L=gpurun_out/lusgs_colshape.log; : > $L
for cfg in "8 4 8" "8 4 4" "8 4 2" "8 8 4" "8 8 2" "4 4 8" "4 4 4" "16 4 4" "16 2 4" "8 2 8"; do
  set -- $cfg
  echo "== n=172 bside=$1 bside2=$2 depth=$3 (column mode)" >> $L
  ICSB200_LUSGS_COLMODE=1 ICSB200_LUSGS_BSIDE=$1 ICSB200_LUSGS_BSIDE2=$2 ICSB200_LUSGS_DEPTH=$3 timeout 600 python tools/lusgs_time.py 172 2>&1 | grep -v "^cells" | sed 's/.*lusgs/lusgs/' >> $L
done
for cfg in "16 16" "16 8" "32 8" "8 16" "8 32" "32 4"; do
  set -- $cfg
  echo "== bump bside=$1 depth=$2 (column mode)" >> $L
  ICSB200_LUSGS_BSIDE=$1 ICSB200_LUSGS_DEPTH=$2 timeout 600 python tools/lusgs_time.py bump 1280 1040 2>&1 | grep -v "^cells" | sed 's/.*lusgs/lusgs/' >> $L
done
cat $L
