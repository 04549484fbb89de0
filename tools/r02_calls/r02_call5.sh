#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "affine_act or bn_backward" > $O/c5_ops.log 2>&1; echo "bn ops tests rc=$?"; tail -2 $O/c5_ops.log
SAP3D_BN_BWD_SLAB=0 timeout 120 python tools/bn_chain_probe.py > $O/c5_bn_chain_coop.txt 2>&1; cat $O/c5_bn_chain_coop.txt
timeout 120 python tools/bn_chain_probe.py > $O/c5_bn_chain_slab.txt 2>&1; cat $O/c5_bn_chain_slab.txt
SAP3D_BN_BWD_SLAB=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c5_bench_noslab.json 2> $O/c5_bench_noslab.err; echo "bench noslab rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c5_bench.json 2> $O/c5_bench.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c5_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['extra']['infer_ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_zz_config_parity_gpu.py > $O/c5_pytest.log 2>&1; echo "gpu suite (w/o config parity) rc=$?"; tail -4 $O/c5_pytest.log
