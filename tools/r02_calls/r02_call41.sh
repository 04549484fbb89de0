#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for cfg in "8 " "32 " "8 --per-sample-bn" "32 --per-sample-bn"; do
  set -- $cfg
  timeout 400 python bench.py --workload eval --clips 1024 --batch $1 ${2:-} --steps 3 --warmup 3 --no-cpu-baseline > $O/c41_eval_b$1${2:+_ps}.json 2> $O/c41_b$1${2:+_ps}.err; echo "eval batch=$1 ${2:-} rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c41_eval_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'])
    except Exception as e:
        print(f,'ERR',e)
PY
