#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_onera344_1gpu.json 2> gpurun_out/r02_bench_onera344_1gpu.err
python bench.py --workload bump4m --steps 10 --warmup 3 > gpurun_out/r02_bench_bump4m_1gpu.json 2> gpurun_out/r02_bench_bump4m_1gpu.err
python tools/show_bench.py gpurun_out/r02_bench_onera344_1gpu.json gpurun_out/r02_bench_bump4m_1gpu.json
