#!/bin/bash
# round-2 ncu evidence (run under gpurun; one report per kernel family so every file stays small).  Every profiled program has
# already exited 0 without ncu in this round's earlier calls (bench.py, tests); the plain runs below repeat that check.
set -u
O=gpurun_out
mkdir -p $O
NCUF="ncu --profile-from-start off --set full --clock-control none --import-source on -f"
NCUS="ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none -f"
G=p3d_unetplusplus_ds
python tools/profile_step.py $G 8 112 train > $O/ev_plain_train.log 2>&1 || { echo "plain train failed"; exit 1; }
# launch lists (cold, serialised) of one training step and one inference step
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_train.csv python tools/profile_step.py $G 8 112 train > $O/ev_ncu_train.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_infer.csv python tools/profile_step.py $G 8 112 infer > $O/ev_ncu_infer.log 2>&1
# dominant conv timed alone
python tools/run_dominant_kernel.py fwd > $O/ev_plain_dom.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc_ -s 1 -c 1 -f -o $O/r02_full_conv_dominant python tools/run_dominant_kernel.py fwd > $O/ev_ncu_dom.log 2>&1
python tools/run_dominant_kernel.py wgrad > $O/ev_plain_wg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 1 -c 1 -f -o $O/r02_full_wgrad_dominant python tools/run_dominant_kernel.py wgrad > $O/ev_ncu_wg.log 2>&1
# inside the training step
timeout 300 $NCUF --kernel-name-base mangled -k regex:conv_tc_kernelILi128ELi4ELi1ELi4E -s 60 -c 1 -o $O/r02_full_conv_splitk python tools/profile_step.py $G 8 112 train > $O/ev_ncu_splitk.log 2>&1
timeout 300 $NCUF -k 'regex:bn_bwd_slab' -s 40 -c 1 -o $O/r02_full_bn_slab python tools/profile_step.py $G 8 112 train > $O/ev_ncu_slab.log 2>&1
timeout 300 $NCUF -k 'regex:apply_bwd_reduce_nob|apply_bwd_nob' -s 2 -c 2 -o $O/r02_full_bn_nob python tools/profile_step.py $G 8 112 train > $O/ev_ncu_nob.log 2>&1
timeout 300 $NCUS -k 'regex:bn_apply_fused|bn_bwd_coop|bn_finalize|apply_kernel|maxpool|pack_multi|adam_kernel|flash_' -c 40 -o $O/r02_sol_step_kernels python tools/profile_step.py $G 8 112 train > $O/ev_ncu_sol.log 2>&1
# GroupNorm + CBAM graph at configs[2] size
python tools/profile_step.py gn:inference_p3d 16 160 train > $O/ev_plain_gn.log 2>&1 && timeout 400 $NCUS -k 'regex:cbam_|sample_channel|gn_' -c 60 -o $O/r02_sol_gn_cbam python tools/profile_step.py gn:inference_p3d 16 160 train > $O/ev_ncu_gn.log 2>&1
python tools/profile_metrics.py > $O/ev_plain_metrics.log 2>&1 && timeout 200 $NCUF -k 'regex:metrics_kernel' -c 1 -o $O/r02_full_metrics python tools/profile_metrics.py > $O/ev_ncu_metrics.log 2>&1
ls -la $O/r02_*.ncu-rep $O/r02_launches_*.csv
du -sh $O
