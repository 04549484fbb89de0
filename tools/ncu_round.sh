#!/bin/bash
# ncu evidence of one round (run under gpurun): launch lists of the training / inference step and --set full captures of the
# dominant kernels.  Every profiled command first runs plain (exit code checked) as the profiling recipe requires.
set -u
O=gpurun_out
G=${1:-p3d_unetplusplus_ds}
python tools/profile_step.py $G 8 112 train > $O/plain_train.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train.csv \
    python tools/profile_step.py $G 8 112 train > $O/ncu_train.log 2>&1
python tools/profile_step.py $G 8 112 infer > $O/plain_infer.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_infer.csv \
    python tools/profile_step.py $G 8 112 infer > $O/ncu_infer.log 2>&1
python tools/run_dominant_kernel.py fwd > $O/plain_dom.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_ -s 1 -c 2 -f -o $O/full_conv_fwd \
    python tools/run_dominant_kernel.py fwd > $O/ncu_dom.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:flash_ -c 3 -f -o $O/full_flash \
    python tools/profile_step.py $G 8 112 train > $O/ncu_flash.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:bn_bwd_coop -s 100 -c 2 -f -o $O/full_bn_coop \
    python tools/profile_step.py $G 8 112 train > $O/ncu_coop.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:conv_tc_kernel.*1, 4>' -s 60 -c 2 -f -o $O/full_conv_splitk \
    python tools/profile_step.py $G 8 112 train > $O/ncu_splitk.log 2>&1
ls -la $O/*.ncu-rep
