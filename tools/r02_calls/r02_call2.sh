#!/bin/bash
# round-2 GPU call 2: parity at the benchmark's configurations, conv phase probe, bench with the new roofline fields
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_zz_config_parity_gpu.py -x -q -s > $O/c2_config_parity.log 2>&1; echo "config parity rc=$?"
grep -E "oracle|CUDA|tap |passed|failed|Error|error" $O/c2_config_parity.log | head -80
python tools/conv_phase_probe.py > $O/c2_conv_phase_probe.txt 2>&1; echo "phase probe rc=$?"
python bench.py --steps 10 --warmup 3 > $O/c2_bench.json 2> $O/c2_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/c2_bench_ref.json 2> $O/c2_bench_ref.err; echo "bench ref rc=$?"
tail -c 1500 $O/c2_bench.json
