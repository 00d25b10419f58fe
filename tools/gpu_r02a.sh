#!/usr/bin/env bash
# round 2, first GPU call: block-tile LU-SGS — parity, sanitizer, kernel times per size and schedule
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lusgs or bitwise or linearity" 2>&1 | tail -15 > gpurun_out/r02a_lusgs_tests.log
tail -5 gpurun_out/r02a_lusgs_tests.log
timeout 300 compute-sanitizer --tool memcheck python tools/lusgs_time.py 16 2>&1 | tail -12 > gpurun_out/r02a_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python tools/lusgs_time.py 16 2>&1 | tail -12 > gpurun_out/r02a_racecheck.log
for n in 128 172 200; do
  for mode in auto level; do
    echo "== n=$n mode=$mode" >> gpurun_out/r02a_lusgs_times.log
    ICSB200_LUSGS_MODE=$mode timeout 600 python tools/lusgs_time.py $n >> gpurun_out/r02a_lusgs_times.log 2>&1
  done
done
echo "== n=344 mode=auto" >> gpurun_out/r02a_lusgs_times.log
timeout 900 python tools/lusgs_time.py 344 >> gpurun_out/r02a_lusgs_times.log 2>&1
echo "== bump 1280x1040 mode=auto" >> gpurun_out/r02a_lusgs_times.log
timeout 600 python tools/lusgs_time_bump.py 1280 1040 >> gpurun_out/r02a_lusgs_times.log 2>&1
cat gpurun_out/r02a_lusgs_times.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_tests.log
tail -5 gpurun_out/r02a_tests.log
