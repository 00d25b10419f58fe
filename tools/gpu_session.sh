#!/usr/bin/env bash
# One GPU call that produces everything a round needs (run from the repo root on the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh r02 [kernel-regex-for-ncu-full]'
# Outputs under gpurun_out/ (copy the ones to keep into profiles/):
#   <tag>_tests.log            pytest -m gpu (full suite, no -x so one failure does not hide the rest)
#   <tag>_bench_default.json   bench.py at the default workload (344^3), <tag>_bench_reference.json the CPU arm
#   <tag>_launches_n160.csv    ncu launch list (gpu__time_duration.sum) of two steps at 160^3
#   <tag>_full.ncu-rep         ncu --set full of the kernels matching the regex (default: k_jac|k_flux_faces|k_lusgs_tma), 160^3
# Numbers printed under ncu are never bench values.
set -u
tag=${1:-rXX}
regex=${2:-"k_jac|k_flux_faces|k_lusgs_tma"}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${tag}_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches_n160.csv \
    python bench.py --cells-per-dim 160 --steps 2 --warmup 1 --skip-cpu --skip-e2e > /dev/null 2>> gpurun_out/${tag}_bench_default.err
ncu --set full --clock-control none --import-source on -k "regex:${regex}" -c 12 -o gpurun_out/${tag}_full -f \
    python bench.py --cells-per-dim 160 --steps 1 --warmup 1 --skip-cpu --skip-e2e > /dev/null 2>> gpurun_out/${tag}_bench_default.err
tail -3 gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_bench_default.json
