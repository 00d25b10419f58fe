#!/usr/bin/env bash
# One 1-GPU call that produces the round's single-GPU evidence (copy into profiles/):
#   tests, smoke, bench lines of the four workloads + the reference arm, LU-SGS kernel times / in-kernel profile / tile trace
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_tests_gpu.log; cat gpurun_out/r02_tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_onera344_1gpu.json 2> gpurun_out/r02_bench_onera344_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench_onera344_1gpu.err
for w in bump4m forwardstep vki; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02_bench_${w}_1gpu.json 2> gpurun_out/r02_bench_${w}_1gpu.err
done
python tools/show_bench.py gpurun_out/r02_bench_onera344_1gpu.json gpurun_out/r02_bench_bump4m_1gpu.json gpurun_out/r02_bench_forwardstep_1gpu.json gpurun_out/r02_bench_vki_1gpu.json
cut -c1-700 gpurun_out/r02_bench_reference.json
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
grep -A1 "== onera\|== bump" $L | grep -v "^--"
