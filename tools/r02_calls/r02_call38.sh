#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_gn_gpu.py -x -q -m gpu -k "not branches" > $O/c38_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c38_tests.log
for E in 1 0; do
SAP3D_EARLY_ADAM=$E timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c38_bench_e$E.json 2> $O/c38_e$E.err; echo "early=$E rc=$?"
done
python - <<'PY'
import json
for e in (1,0):
    try:
        d=json.loads(open('gpurun_out/c38_bench_e%d.json'%e).read().strip().splitlines()[-1])
        print('early adam',e, round(d['value'],1), round(d['ms_per_step'],3), d['gpu_launches'], round(d['e2e']['value'],1))
    except Exception as ex:
        print(e, 'ERR', ex)
PY
tail -3 $O/c38_e1.err
