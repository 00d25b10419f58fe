#!/usr/bin/env bash
set -u
N=$1; tag=$2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 10 --warmup 3 --workload bump4m > gpurun_out/${tag}_bench_bump4m_${N}gpu.json 2> gpurun_out/${tag}_bench_bump4m_${N}gpu.err
python tools/show_bench.py gpurun_out/${tag}_bench_bump4m_${N}gpu.json
