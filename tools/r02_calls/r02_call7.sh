#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "affine_act or bn_backward" > $O/c7_ops.log 2>&1; echo "bn ops tests rc=$?"; tail -2 $O/c7_ops.log
timeout 120 python tools/bn_chain_probe.py > $O/c7_bn_chain_slab.txt 2>&1; cat $O/c7_bn_chain_slab.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c7_bench.json 2> $O/c7_bench.err; echo "bench rc=$?"
SAP3D_BN_BWD_SLAB=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c7_bench_noslab.json 2> $O/c7_bench_noslab.err; echo "bench noslab rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c7_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['extra']['infer_ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 900 python -m pytest tests/test_metrics.py tests/test_zz_network_surface_gpu.py -x -q -m gpu > $O/c7_new_tests.log 2>&1; echo "new tests rc=$?"; tail -6 $O/c7_new_tests.log
