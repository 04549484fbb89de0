#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for C in 32 1; do
CUDA_DEVICE_MAX_CONNECTIONS=$C timeout 600 $TR tools/dp_timeline.py 8 112 10 > $O/c18_dp_timeline_conn$C.log 2>&1; echo "== CUDA_DEVICE_MAX_CONNECTIONS=$C"; grep "dp_timeline" $O/c18_dp_timeline_conn$C.log
done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c18_bench_conn32.json 2> $O/c18_a.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c18_bench_default.json 2> $O/c18_b.err
python - <<'PY'
import json
for f in ['c18_bench_conn32','c18_bench_default']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'))
    except Exception as e:
        print(f, 'ERR', e)
PY
