#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_conv_gpu.py -x -q -m gpu > $O/c21_conv_tests.log 2>&1; echo "conv tests rc=$?"; tail -3 $O/c21_conv_tests.log
for S in 0 1; do SAP3D_WGRAD_PAIR=$S timeout 300 python tools/run_dominant_kernel.py wgrad > $O/c21_wg_$S.log 2>&1; echo "pair=$S: $(tail -1 $O/c21_wg_$S.log)"; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c21_bench.json 2> $O/c21_train.err; echo "train rc=$?"
SAP3D_WGRAD_PAIR=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c21_bench_nopair.json 2> $O/c21_trainb.err; echo "train nopair rc=$?"
python - <<'PY'
import json
for f in ['c21_bench','c21_bench_nopair']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'))
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:wgrad_tc_pair -s 1 -c 1 -o $O/r02_full_wgrad_pair python tools/run_dominant_kernel.py wgrad > $O/c21_ncu.log 2>&1; tail -1 $O/c21_ncu.log
