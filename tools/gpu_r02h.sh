#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for w in bump4m forwardstep vki; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02h_bench_$w.json 2> gpurun_out/r02h_bench_$w.err
  tail -3 gpurun_out/r02h_bench_$w.err; cut -c1-1200 gpurun_out/r02h_bench_$w.json
done
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02h_bench_onera344.json 2> gpurun_out/r02h_bench_onera344.err
tail -3 gpurun_out/r02h_bench_onera344.err; cat gpurun_out/r02h_bench_onera344.json
