#!/usr/bin/env bash
# N-GPU evidence: /usr/local/graft/bin/gpurun --gpus N --timeout 2400 -- 'bash tools/gpu_multi.sh N tag [tests] [workloads...]'
#   tests: pytest tests/test_gpu_multi.py + the full tests/multi_gpu_check.py log of every variant (parity vs the P-rank oracle world)
#   then bench.py under torchrun for each workload (default: onera344 bump4m); every line carries the `parity` record of its pre-flight
set -u
N=$1; tag=$2; shift 2
tests=no
if [ "${1:-}" = "tests" ]; then tests=tests; shift; fi
workloads=${*:-onera344 bump4m}
mkdir -p gpurun_out
if [ "$tests" = "tests" ]; then
  timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -15 > gpurun_out/${tag}_multi_tests.log
  : > gpurun_out/${tag}_multi_parity.log
  for v in "" globaldt mrf hb fullvisc vki decomposed; do
    echo "== variant '${v}'" >> gpurun_out/${tag}_multi_parity.log
    ICS_MULTI_VARIANT=$v ICS_MULTI_MU=$([ "$v" = fullvisc ] && echo 0.3 || echo 0) timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29561 tests/multi_gpu_check.py 2>&1 | grep -E "history|rel err|MULTI_GPU_PARITY" >> gpurun_out/${tag}_multi_parity.log
  done
  cat gpurun_out/${tag}_multi_tests.log; grep -c "PARITY OK" gpurun_out/${tag}_multi_parity.log
fi
for w in $workloads; do
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 10 --warmup 3 --workload $w > gpurun_out/${tag}_bench_${w}_${N}gpu.json 2> gpurun_out/${tag}_bench_${w}_${N}gpu.err
  python tools/show_bench.py gpurun_out/${tag}_bench_${w}_${N}gpu.json
done
