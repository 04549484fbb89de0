#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 480 python -m pytest tests -x -q -m gpu > $O/c46_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 $O/c46_tests.log
