#!/bin/bash
# final 1-GPU pass: whole GPU test suite, evidence captures, bench lines
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/c24_tests.log 2>&1; echo "gpu tests rc=$?"; tail -6 $O/c24_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/r02_bench_train_ds_b8.json 2> $O/c24_b8.err; echo "b8 rc=$?"
timeout 400 python bench.py --steps 5 --warmup 3 --batch 32 --no-cpu-baseline > $O/r02_bench_train_ds_b32.json 2> $O/c24_b32.err; echo "b32 rc=$?"
timeout 400 python bench.py --steps 5 --warmup 3 --batch 16 --no-cpu-baseline > $O/r02_bench_train_ds_b16.json 2> $O/c24_b16.err; echo "b16 rc=$?"
timeout 400 python bench.py --workload eval --clips 1024 --steps 3 --warmup 3 --no-cpu-baseline > $O/r02_bench_eval_1024clips.json 2> $O/c24_eval.err; echo "eval rc=$?"
timeout 600 python bench.py --workload gn160 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_gn160_b16.json 2> $O/c24_gn.err; echo "gn160 rc=$?"
python - <<'PY'
import json,glob
for f in ['r02_bench_train_ds_b8','r02_bench_train_ds_b32','r02_bench_train_ds_b16','r02_bench_eval_1024clips','r02_bench_gn160_b16']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d['gpu_launches'], round(d['e2e']['value'],1))
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 1500 bash tools/r02_evidence.sh > $O/c24_evidence.log 2>&1; echo "evidence rc=$?"; tail -12 $O/c24_evidence.log
timeout 300 python tools/trace_step.py --out $O/r02_ingraph_trace_train_ds_b8.txt > $O/c24_trace.log 2>&1; head -3 $O/r02_ingraph_trace_train_ds_b8.txt
