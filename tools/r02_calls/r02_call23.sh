#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "finished_inside" > $O/c23_fuse_tests.log 2>&1; echo "fuse tests rc=$?"; tail -3 $O/c23_fuse_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c23_bench.json 2> $O/c23_train.err; echo "train rc=$?"
SAP3D_CONV_FUSE_BN=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c23_bench_nofuse.json 2> $O/c23_trainb.err; echo "train nofuse rc=$?"
python - <<'PY'
import json
for f in ['c23_bench','c23_bench_nofuse']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 300 python tools/trace_step.py --out $O/c23_trace.txt > $O/c23_trace.log 2>&1; head -12 $O/c23_trace.txt
