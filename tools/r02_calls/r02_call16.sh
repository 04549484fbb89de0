#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_conv_gpu.py -x -q -m gpu > $O/c16_conv_tests.log 2>&1; echo "conv tests rc=$?"; tail -4 $O/c16_conv_tests.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_flash_gpu.py -x -q -m gpu > $O/c16_model_tests.log 2>&1; echo "model tests rc=$?"; tail -4 $O/c16_model_tests.log
timeout 300 python tools/run_dominant_kernel.py fwd > $O/c16_dom.log 2>&1; echo "dominant: $(tail -1 $O/c16_dom.log)"
timeout 300 python tools/run_dominant_kernel.py dgrad2 > $O/c16_dom_dgrad2.log 2>&1; echo "dgrad2: $(tail -1 $O/c16_dom_dgrad2.log)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c16_bench.json 2> $O/c16_train.err; echo "train rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --batch 32 --no-cpu-baseline > $O/c16_bench_b32.json 2> $O/c16_train_b32.err; echo "train b32 rc=$?"
python - <<'PY'
import json
for f in ['c16_bench','c16_bench_b32']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
