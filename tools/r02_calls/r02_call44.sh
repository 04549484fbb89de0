#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_metrics.py -x -q -m gpu -k "per_clip or per_sample or stem_cache or video" > $O/c44_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/c44_tests.log
for F in 1 0; do
SAP3D_SAMPLE_NORM_FUSED=$F timeout 300 python bench.py --workload eval --clips 1024 --steps 3 --warmup 3 --no-cpu-baseline > $O/c44_eval_f$F.json 2> $O/c44_f$F.err; echo "fused=$F rc=$?"
done
python - <<'PY'
import json
for f in (1,0):
    try:
        d=json.loads(open('gpurun_out/c44_eval_f%d.json'%f).read().strip().splitlines()[-1])
        print('fused',f, round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], d['config']['metric_means'])
    except Exception as e:
        print(f,'ERR',e)
PY
