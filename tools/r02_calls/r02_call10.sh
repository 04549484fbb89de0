#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gn_gpu.py tests/test_ops_gpu.py tests/test_metrics.py tests/test_zz_network_surface_gpu.py -x -q -m gpu > $O/c10_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/c10_tests.log
NCUF="ncu --profile-from-start off --set full --clock-control none --import-source on -f"
NCUS="ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none -f"
G=p3d_unetplusplus_ds
timeout 300 $NCUF --kernel-name-base demangled -k 'regex:conv_tc_kernel<128, 4, 1, 4>' -s 60 -c 1 -o $O/r02_full_conv_splitk python tools/profile_step.py $G 8 112 train > $O/c10_ncu_splitk.log 2>&1; tail -2 $O/c10_ncu_splitk.log
timeout 300 $NCUS -k 'regex:bn_apply_fused|bn_bwd_coop|bn_finalize|apply_kernel|maxpool|pack_multi|adam_kernel|flash_' -c 40 -o $O/r02_sol_step_kernels python tools/profile_step.py $G 8 112 train > $O/c10_ncu_sol.log 2>&1
timeout 400 $NCUS -k 'regex:cbam_|sample_channel|gn_' -c 60 -o $O/r02_sol_gn_cbam python tools/profile_step.py gn:inference_p3d 16 160 train > $O/c10_ncu_gn.log 2>&1
timeout 200 $NCUF -k 'regex:metrics_kernel' -c 1 -o $O/r02_full_metrics python tools/profile_metrics.py > $O/c10_ncu_metrics.log 2>&1
timeout 600 python bench.py --workload gn160 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_gn160_b16.json 2> $O/c10_gn.err; echo "gn160 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c10_bench.json 2> $O/c10_train.err; echo "train rc=$?"
python - <<'PY'
import json,glob
for f in ['gpurun_out/r02_bench_gn160_b16.json','gpurun_out/c10_bench.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d.get('roofline',{}).get('frac'), d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
ls -la $O/r02_*.ncu-rep
