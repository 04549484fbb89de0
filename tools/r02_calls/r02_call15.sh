#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -x -q -m gpu -k "swap or repack or split_backward or training_step" > $O/c15_tests.log 2>&1; echo "tests rc=$?"; tail -6 $O/c15_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c15_bench.json 2> $O/c15_train.err; echo "train rc=$?"
python - <<'PY'
import json
for f in ['c15_bench']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 300 python tools/trace_step.py --out $O/r02_ingraph_trace_train_ds_b8.txt > $O/c15_trace.log 2>&1; tail -3 $O/c15_trace.log; head -30 $O/r02_ingraph_trace_train_ds_b8.txt
