#!/usr/bin/env bash
# usage: gpu_multi.sh N tag [tests]   — run under `gpurun --gpus N`
set -u
N=$1; tag=$2; tests=${3:-no}
mkdir -p gpurun_out
if [ "$tests" = "tests" ]; then
  timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -15 > gpurun_out/${tag}_multi_tests.log
  for v in "" globaldt mrf hb fullvisc vki decomposed; do
    echo "== variant '${v}'" >> gpurun_out/${tag}_multi_parity.log
    ICS_MULTI_VARIANT=$v ICS_MULTI_MU=$([ "$v" = fullvisc ] && echo 0.3 || echo 0) timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29561 tests/multi_gpu_check.py 2>&1 | grep -E "history|rel err|MULTI_GPU_PARITY" >> gpurun_out/${tag}_multi_parity.log
  done
  cat gpurun_out/${tag}_multi_tests.log; grep -c "PARITY OK" gpurun_out/${tag}_multi_parity.log
fi
for w in onera344 bump4m; do
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 10 --warmup 3 --workload $w > gpurun_out/${tag}_bench_${w}_${N}gpu.json 2> gpurun_out/${tag}_bench_${w}_${N}gpu.err
  tail -2 gpurun_out/${tag}_bench_${w}_${N}gpu.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_${w}_${N}gpu.json"))
    print("$w N=$N value", d["value"], "ms/step", d["ms_per_step"], "parity", d.get("parity"), "e2e", d["e2e"])
    print({k:(v["ms_per_step"],v.get("frac")) for k,v in d["roofline"]["kernel_classes"].items()})
except Exception as e: print("no line", e)
PY
done
