#!/bin/bash
# 8-GPU confirmation of the final data-parallel default
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02_bench_train_ds_8gpu.json 2> $O/c25_train.err; echo "train8 rc=$?"
timeout 400 $TR bench.py --gpus 8 --workload eval --clips 1024 --steps 3 --warmup 3 > $O/r02_bench_eval_1024clips_8gpu.json 2> $O/c25_eval.err; echo "eval8 rc=$?"
python - <<'PY'
import json
for f in ['r02_bench_train_ds_8gpu','r02_bench_eval_1024clips_8gpu']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('extra',{}).get('replicas_identical'), d['gpu_launches'], round(d['e2e']['value'],1))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -2 $O/c25_train.err
