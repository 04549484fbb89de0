#!/bin/bash
# 2-GPU call: NCCL data-parallel correctness + the overlapped multi-segment exchange, A/B against the single split
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_zz_dp_nccl_gpu.py tests/test_model_gpu.py -x -q -m gpu -k "nccl or split_backward" -s > $O/c12_tests.log 2>&1; echo "tests rc=$?"; tail -6 $O/c12_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c12_bench_1gpu.json 2> $O/c12_1gpu.err; echo "1gpu rc=$?"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02_bench_train_ds_2gpu.json 2> $O/c12_2gpu.err; echo "2gpu rc=$?"
SAP3D_DP_MARKS=1 timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/c12_bench_2gpu_onesplit.json 2> $O/c12_2gpu_b.err; echo "2gpu one split rc=$?"
NCCL_MAX_CTAS=8 timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/c12_bench_2gpu_maxctas8.json 2> $O/c12_2gpu_c.err; echo "2gpu maxctas rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload eval --clips 256 --steps 3 --warmup 3 > $O/r02_bench_eval_2gpu.json 2> $O/c12_eval.err; echo "eval rc=$?"
python - <<'PY'
import json
for f in ['c12_bench_1gpu','r02_bench_train_ds_2gpu','c12_bench_2gpu_onesplit','c12_bench_2gpu_maxctas8','r02_bench_eval_2gpu']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('replicas_identical'), d.get('extra',{}).get('replicas_identical'), d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 $O/c12_2gpu.err
