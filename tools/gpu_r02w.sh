#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_lusgs_blk" -s 4 -c 1 -o gpurun_out/r02_full_344_col -f \
    python tools/lusgs_time.py 344 > gpurun_out/r02_ncu_344_col.log 2>&1
tail -3 gpurun_out/r02_ncu_344_col.log
ls -la gpurun_out/r02_full_344_col.ncu-rep
