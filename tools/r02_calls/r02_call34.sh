#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu > $O/c34_tests.log 2>&1; echo "model tests rc=$?"; tail -4 $O/c34_tests.log
for B in 1 0; do
SAP3D_BRANCHES=$B timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c34_bench_br$B.json 2> $O/c34_br$B.err; echo "branches=$B rc=$?"; tail -2 $O/c34_br$B.err
done
python - <<'PY'
import json
for b in (1,0):
    try:
        d=json.loads(open('gpurun_out/c34_bench_br%d.json'%b).read().strip().splitlines()[-1])
        print('branches',b, round(d['value'],1), round(d['ms_per_step'],3), d['gpu_launches'], d['extra'].get('infer_ms_per_step'))
    except Exception as e:
        print(b, 'ERR', e)
PY
