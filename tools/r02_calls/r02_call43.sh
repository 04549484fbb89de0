#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_zz_config_parity_gpu.py tests/test_zz_network_surface_gpu.py -x -q -m gpu > $O/c43_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/c43_tests.log
