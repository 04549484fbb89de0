#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_zz_dp_nccl_gpu.py -x -q -m gpu > $O/c39_tests.log 2>&1; echo "nccl tests rc=$?"; tail -3 $O/c39_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02_bench_train_ds_2gpu.json 2> $O/c39_2gpu.err; echo "2gpu rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_train_ds_2gpu.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), d.get('extra',{}).get('replicas_identical'), d['gpu_launches'], round(d['e2e']['value'],1))
PY
