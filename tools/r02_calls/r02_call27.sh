#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_gn_gpu.py -x -q -m gpu -k "not opt_in" > $O/c27_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c27_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c27_bench.json 2> $O/c27_a.err; echo "train rc=$?"
SAP3D_CONV_CTA_ROWS=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c27_bench_tilerows.json 2> $O/c27_b.err; echo "train tile rows rc=$?"
python - <<'PY'
import json
for f in ['c27_bench','c27_bench_tilerows']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
