#!/usr/bin/env bash
# ncu evidence of the round: launch list (160^3, 2 steps) + full-set capture of k_lusgs_blk and k_spmv at bench size (344^3)
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_n160.csv \
    python bench.py --cells-per-dim 160 --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-extra > /dev/null 2> gpurun_out/r02_ncu.err
wc -l gpurun_out/r02_launches_n160.csv
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:k_lusgs_blk|k_spmv" -s 4 -c 2 -o gpurun_out/r02_full_344 -f \
    python tools/lusgs_time.py 344 > gpurun_out/r02_ncu_344.log 2>&1
tail -3 gpurun_out/r02_ncu_344.log
ls -la gpurun_out/r02_full_344.ncu-rep
