#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -x -q -m gpu -k "not opt_in and not config" > $O/c26_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c26_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c26_bench.json 2> $O/c26_a.err; echo "train rc=$?"
SAP3D_CONV_NARROW=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c26_bench_wide.json 2> $O/c26_b.err; echo "train wide rc=$?"
python - <<'PY'
import json
for f in ['c26_bench','c26_bench_wide']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 300 python tools/trace_step.py --out $O/c26_trace.txt > $O/c26_trace.log 2>&1; grep -n "conv_tc_persist\|span" $O/c26_trace.txt | head
