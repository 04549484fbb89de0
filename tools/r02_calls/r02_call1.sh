#!/bin/bash
# round-2 GPU call 1: validate the merged r02 branches (batched attention GEMMs, multicast conv, split/finalize tweaks),
# A/B bench lines, in-graph trace.  Risky kernels (multicast clusters) run last under their own timeout.
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_zz_conv_multicast_gpu.py > $O/c1_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/c1_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c1_bench_base.json 2> $O/c1_bench_base.err; echo "bench base rc=$?"
SAP3D_ATTN_BATCHED=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c1_bench_attnb.json 2> $O/c1_bench_attnb.err; echo "bench attnb rc=$?"
python tools/trace_step.py --out $O/c1_trace_train.txt > /dev/null 2> $O/c1_trace.err; echo "trace rc=$?"
SAP3D_ATTN_BATCHED=1 python tools/trace_step.py --out $O/c1_trace_train_attnb.txt > /dev/null 2>> $O/c1_trace.err; echo "trace attnb rc=$?"
timeout 400 python -m pytest tests/test_zz_conv_multicast_gpu.py -x -q > $O/c1_mc.log 2>&1; echo "multicast test rc=$?"
tail -5 $O/c1_mc.log
if grep -q "2 passed" $O/c1_mc.log; then
  SAP3D_CONV_MULTICAST=2 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c1_bench_mc2.json 2> $O/c1_bench_mc2.err; echo "bench mc2 rc=$?"
  SAP3D_CONV_MULTICAST=4 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c1_bench_mc4.json 2> $O/c1_bench_mc4.err; echo "bench mc4 rc=$?"
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c1_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
