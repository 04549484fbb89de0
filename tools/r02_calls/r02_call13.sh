#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -k "halo or swap or linearity" > $O/c13_tests.log 2>&1; echo "swap tests rc=$?"; tail -15 $O/c13_tests.log
for S in 0 1; do
  SAP3D_CONV_SWAP=$S timeout 300 python tools/run_dominant_kernel.py fwd > $O/c13_dom_s$S.log 2>&1; echo "swap=$S: $(tail -1 $O/c13_dom_s$S.log)"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c13_bench.json 2> $O/c13_train.err; echo "train rc=$?"
SAP3D_CONV_SWAP=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c13_bench_noswap.json 2> $O/c13_train_b.err; echo "train(noswap) rc=$?"
python - <<'PY'
import json
for f in ['c13_bench','c13_bench_noswap']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dp_timeline.py 8 112 10 > $O/c13_dp_timeline.log 2>&1; grep dp_timeline -A1 $O/c13_dp_timeline.log
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:conv_tc_swap -s 1 -c 1 -o $O/r02_full_conv_swap python tools/run_dominant_kernel.py fwd > $O/c13_ncu.log 2>&1; tail -1 $O/c13_ncu.log
