#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_model_gpu.py -x -q -m gpu > $O/c35_tests.log 2>&1; echo "model tests rc=$?"; tail -4 $O/c35_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c35_bench.json 2> $O/c35.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c35_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), d['gpu_launches'], d['extra'].get('infer_ms_per_step'))
PY
