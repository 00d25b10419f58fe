#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_tutorials.py -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_r02k.sh 2>&1 | tail -16
timeout 600 python bench.py --cells-per-dim 200 --steps 3 --warmup 2 --skip-cpu --skip-e2e --skip-extra 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print({k:(v['ms_per_step'],v.get('frac')) for k,v in d['roofline']['kernel_classes'].items()})"
