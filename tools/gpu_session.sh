#!/usr/bin/env bash
# One 1-GPU call that produces the single-GPU evidence of a round (run from the repo root on the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash tools/gpu_session.sh r02 [ncu]'
# Outputs under gpurun_out/ (copy the ones to keep into profiles/):
#   <tag>_tests_gpu.log                 pytest -m gpu (full suite), then __graft_entry__.smoke()
#   <tag>_bench_<workload>_1gpu.json    bench.py lines of the four named workloads, <tag>_bench_reference.json the CPU arm
#   <tag>_lusgs_times.log               tools/lusgs_time.py: LU-SGS / SpMV kernel times + the in-kernel cycle profile of k_lusgs_blk
#   <tag>_lusgs_trace.log               tools/lusgs_blk_trace.py: per-tile timeline
# with a second argument `ncu` also
#   <tag>_launches_n160.csv             ncu launch list (gpu__time_duration.sum) of two steps at 160^3
#   <tag>_full_344.ncu-rep              ncu --set full of k_lusgs_blk and k_spmv at bench size (344^3)
# Numbers printed under ncu are never bench values.
set -u
tag=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${tag}_tests_gpu.log; cat gpurun_out/${tag}_tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_onera344_1gpu.json 2> gpurun_out/${tag}_bench_onera344_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench_onera344_1gpu.err
for w in bump4m forwardstep vki; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/${tag}_bench_${w}_1gpu.json 2> gpurun_out/${tag}_bench_${w}_1gpu.err
done
python tools/show_bench.py gpurun_out/${tag}_bench_onera344_1gpu.json gpurun_out/${tag}_bench_bump4m_1gpu.json gpurun_out/${tag}_bench_forwardstep_1gpu.json gpurun_out/${tag}_bench_vki_1gpu.json
L=gpurun_out/${tag}_lusgs_times.log; : > $L
for n in 64 128 172 200 344; do
  echo "== onera box n=$n" >> $L
  timeout 900 python tools/lusgs_time.py $n >> $L 2>&1
done
echo "== bump 3x1280x1040" >> $L
timeout 900 python tools/lusgs_time.py bump 1280 1040 >> $L 2>&1
for n in 172 344; do
  echo "== in-kernel cycle profile, n=$n (ICSB200_LUSGS_PROF=1: the instrumented instantiation is ~20 % slower)" >> $L
  ICSB200_LUSGS_PROF=1 timeout 900 python tools/lusgs_time.py $n 2>&1 | grep -v "^cells" >> $L
done
T=gpurun_out/${tag}_lusgs_trace.log; : > $T
for n in 172 344; do
  echo "== n=$n" >> $T
  timeout 900 python tools/lusgs_blk_trace.py $n >> $T 2>&1
done
if [ "${2:-}" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches_n160.csv \
      python bench.py --cells-per-dim 160 --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-extra > /dev/null 2>> gpurun_out/${tag}_bench_onera344_1gpu.err
  timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:k_lusgs_blk|k_spmv" -s 4 -c 2 -o gpurun_out/${tag}_full_344 -f \
      python tools/lusgs_time.py 344 > gpurun_out/${tag}_ncu_344.log 2>&1
fi
