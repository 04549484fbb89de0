#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "affine_act or bn_backward" > $O/c6_ops.log 2>&1; echo "bn ops tests rc=$?"; tail -2 $O/c6_ops.log
timeout 120 python tools/bn_chain_probe.py > $O/c6_bn_chain_slab.txt 2>&1; cat $O/c6_bn_chain_slab.txt
timeout 120 python tools/conv_phase_probe.py > $O/c6_conv_phase_probe.txt 2>&1; echo "phase probe rc=$?"
grep -E "^==|partials received|epilogue|exit" $O/c6_conv_phase_probe.txt | head -40
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c6_bench.json 2> $O/c6_bench.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c6_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['extra']['infer_ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
