#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_gn_gpu.py -x -q -m gpu > $O/c32_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c32_tests.log
for S in 1 2 3 4; do
SAP3D_WGRAD_STREAMS=$S timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c32_bench_s$S.json 2> $O/c32_s$S.err; echo "streams=$S rc=$?"
done
python - <<'PY'
import json
for s in (1,2,3,4):
    try:
        d=json.loads(open('gpurun_out/c32_bench_s%d.json'%s).read().strip().splitlines()[-1])
        print('streams',s, round(d['value'],1), round(d['ms_per_step'],3), d['gpu_launches'])
    except Exception as e:
        print(s, 'ERR', e)
PY
