#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for I in 4 8; do SAP3D_NOB_INFLIGHT=$I timeout 200 python tools/bn_bwd_probe.py > $O/c19_bn_probe_$I.log 2>&1; cat $O/c19_bn_probe_$I.log | tail -3; done
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -x -q -m gpu -k "affine or bn or norm or repack or training_step" > $O/c19_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c19_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c19_bench.json 2> $O/c19_train.err; echo "train rc=$?"
SAP3D_NOB_INFLIGHT=4 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c19_bench_if4.json 2> $O/c19_train4.err; echo "train if4 rc=$?"
python - <<'PY'
import json
for f in ['c19_bench','c19_bench_if4']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'))
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 300 python tools/trace_step.py --out $O/c19_trace.txt > $O/c19_trace.log 2>&1; grep -n "pack_multi\|adam_kernel\|apply_bwd_reduce_nob\|apply_bwd_nob\|span" $O/c19_trace.txt | head
