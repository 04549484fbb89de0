#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_metrics.py -x -q -m gpu > $O/c9_metrics_tests.log 2>&1; echo "metrics tests rc=$?"; tail -2 $O/c9_metrics_tests.log
bash tools/r02_evidence.sh > $O/c9_evidence.log 2>&1; echo "evidence rc=$?"; tail -22 $O/c9_evidence.log
timeout 600 python bench.py --workload gn160 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_gn160_b16.json 2> $O/c9_gn.err; echo "gn160 rc=$?"
timeout 600 python bench.py --workload eval --clips 1024 --steps 5 > $O/r02_bench_eval_1024clips.json 2> $O/c9_eval.err; echo "eval rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_train_ds_b8.json 2> $O/c9_train.err; echo "train rc=$?"
timeout 600 python bench.py --batch 32 --steps 8 --warmup 3 --no-cpu-baseline > $O/r02_bench_train_ds_b32.json 2> $O/c9_b32.err; echo "b32 rc=$?"
python tools/trace_step.py --out $O/r02_ingraph_trace_train_ds_b8.txt > /dev/null 2> $O/c9_trace.err; echo "trace rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('in_graph',{}).get('frac'), d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
