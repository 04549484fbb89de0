#!/bin/bash
# ncu --set full captures of the bandwidth-bound kernels (norm apply / BN backward at decoder size, GroupNorm + CBAM of the
# 160 x 160 graph, saliency metrics): achieved DRAM bytes / duration against the HBM peak.  Run under gpurun; every profiled
# program first runs plain (the BN training step has been exercised by the bench already).
set -u
O=gpurun_out
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k 'regex:apply_bwd_kernel|apply_bwd_reduce' -c 8 -o $O/full_bw_norm python tools/profile_step.py p3d_unetplusplus_ds 8 112 train > $O/ncu_bw_norm.log 2>&1
python tools/profile_step.py gn:inference_p3d 16 160 train > $O/plain_gn.log 2>&1 &&
timeout 200 $NCU -k 'regex:cbam_|sample_channel|gn_finalize' -c 14 -o $O/full_bw_gn python tools/profile_step.py gn:inference_p3d 16 160 train > $O/ncu_bw_gn.log 2>&1
timeout 200 $NCU -k 'regex:gn_bwd_reduce|gn_bwd_apply|cbam_bwd_pos|cbam_bwd_apply' -c 8 -o $O/full_bw_gn_bwd python tools/profile_step.py gn:inference_p3d 16 160 train > $O/ncu_bw_gn_bwd.log 2>&1
python tools/profile_metrics.py > $O/plain_metrics.log 2>&1 &&
timeout 200 $NCU -k 'regex:metrics_kernel|auc_|resize_bilinear' -c 8 -o $O/full_bw_metrics python tools/profile_metrics.py > $O/ncu_bw_metrics.log 2>&1
tail -2 $O/plain_gn.log $O/plain_metrics.log
ls -la $O/full_bw_*.ncu-rep
