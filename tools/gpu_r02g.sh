#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02g_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench_default.json 2> gpurun_out/r02g_bench_default.err
tail -5 gpurun_out/r02g_tests.log
cat gpurun_out/r02g_bench_default.json
tail -5 gpurun_out/r02g_bench_default.err
