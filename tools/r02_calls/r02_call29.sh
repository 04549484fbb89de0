#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 200 python tools/bn_bwd_probe.py > $O/c29_bn_probe.log 2>&1; tail -3 $O/c29_bn_probe.log
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "affine or bn or norm" > $O/c29_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/c29_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c29_bench.json 2> $O/c29_a.err; echo "train rc=$?"
python - <<'PY'
import json
for f in ['c29_bench']:
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d['gpu_launches'], round(d['e2e']['value'],1))
PY
