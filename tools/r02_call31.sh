#!/bin/bash
# compute-sanitizer over the new kernels (halo / swap / dynamic scheduler / fused BN / pair wgrad): memcheck, then racecheck on one case
set -u
O=gpurun_out
mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --print-limit 20 --error-exitcode 3 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "swap and 784 or halo and 625 or finished_inside and conv3 or finished_inside and convS or persistent and 128" > $O/c31_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/c31_memcheck.log | head -10
timeout 600 $CS --tool racecheck --racecheck-report analysis --print-limit 20 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "swap and 784" > $O/c31_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard|passed|failed" $O/c31_racecheck.log | head -10
