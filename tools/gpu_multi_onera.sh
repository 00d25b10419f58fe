#!/usr/bin/env bash
# usage: gpu_multi_onera.sh N tag — C4 bench under torchrun on N GPUs (run under `gpurun --gpus N`)
set -u
N=$1; tag=$2
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${tag}_bench_onera344_${N}gpu.json 2> gpurun_out/${tag}_bench_onera344_${N}gpu.err
tail -2 gpurun_out/${tag}_bench_onera344_${N}gpu.err | cut -c1-300
python tools/show_bench.py gpurun_out/${tag}_bench_onera344_${N}gpu.json
