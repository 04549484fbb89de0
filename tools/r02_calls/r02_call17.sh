#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dp_timeline.py 8 112 10 > $O/c17_dp_timeline.log 2>&1; grep "dp_timeline" $O/c17_dp_timeline.log
timeout 900 python -m pytest tests/test_zz_dp_nccl_gpu.py -x -q -m gpu > $O/c17_tests.log 2>&1; echo "nccl tests rc=$?"; tail -3 $O/c17_tests.log
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02_bench_train_ds_2gpu.json 2> $O/c17_2gpu.err; echo "2gpu rc=$?"
python - <<'PY'
import json
for f in ['r02_bench_train_ds_2gpu']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('extra',{}).get('replicas_identical'), d['gpu_launches'], d.get('roofline',{}).get('frac'))
    except Exception as e:
        print(f, 'ERR', e)
PY
