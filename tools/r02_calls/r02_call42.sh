#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 8 --workload eval --clips 1024 --steps 3 --warmup 3 > $O/r02_bench_eval_1024clips_8gpu.json 2> $O/c42_eval.err; echo "eval8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_eval_1024clips_8gpu.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['config']['workload'][-120:])
PY
