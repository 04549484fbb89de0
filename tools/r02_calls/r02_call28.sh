#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_metrics.py -x -q -m gpu > $O/c28_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c28_tests.log
timeout 400 python bench.py --workload eval --clips 1024 --steps 3 --warmup 3 --no-cpu-baseline > $O/r02_bench_eval_1024clips.json 2> $O/c28_eval.err; echo "eval rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c28_bench.json 2> $O/c28_a.err; echo "train rc=$?"
python - <<'PY'
import json
for f in ['r02_bench_eval_1024clips','c28_bench']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d['gpu_launches'], round(d['e2e']['value'],1))
    except Exception as e:
        print(f, 'ERR', e)
PY
