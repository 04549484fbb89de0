#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_conv_gpu.py -q -m gpu -k "halo or linearity" > $O/c11_halo_tests.log 2>&1; echo "halo tests rc=$?"; tail -15 $O/c11_halo_tests.log
timeout 1200 python -m pytest tests/test_conv_gpu.py tests/test_ops_gpu.py -x -q -m gpu -k "not halo" > $O/c11_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/c11_tests.log
for H in 0 1; do for B in 0 1; do
  SAP3D_CONV_HALO=$H SAP3D_CONV_BALANCED=$B timeout 300 python tools/run_dominant_kernel.py fwd > $O/c11_dom_h${H}_b${B}.log 2>&1; echo "halo=$H balanced=$B: $(tail -1 $O/c11_dom_h${H}_b${B}.log)"
done; done
for H in 0 1; do SAP3D_CONV_HALO=$H timeout 300 python tools/run_dominant_kernel.py dgrad2 > $O/c11_dom_dgrad2_h$H.log 2>&1; echo "dgrad2 halo=$H: $(tail -1 $O/c11_dom_dgrad2_h$H.log)"; done
timeout 300 python tools/run_dominant_kernel.py wgrad > $O/c11_dom_wgrad.log 2>&1; echo "wgrad: $(tail -1 $O/c11_dom_wgrad.log)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c11_bench.json 2> $O/c11_train.err; echo "train rc=$?"
SAP3D_CONV_HALO=0 SAP3D_CONV_BALANCED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c11_bench_off.json 2> $O/c11_train_off.err; echo "train(off) rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/c11_bench.json','gpurun_out/c11_bench_off.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
NCUF="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCUF -k regex:conv_tc_persist -s 1 -c 1 -o $O/r02_full_conv_halo python tools/run_dominant_kernel.py fwd > $O/c11_ncu_conv.log 2>&1; tail -1 $O/c11_ncu_conv.log
timeout 300 $NCUF -k regex:wgrad_tc -s 1 -c 1 -o $O/r02_full_wgrad_dominant python tools/run_dominant_kernel.py wgrad > $O/c11_ncu_wg.log 2>&1; tail -1 $O/c11_ncu_wg.log
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -f --kernel-name-base mangled -k regex:conv_tc_kernelILi128ELi4ELi1ELi4E -s 60 -c 1 -o $O/r02_full_conv_splitk python tools/profile_step.py p3d_unetplusplus_ds 8 112 train > $O/c11_ncu_splitk.log 2>&1; tail -1 $O/c11_ncu_splitk.log
ls -la $O/*.ncu-rep | tail -5
