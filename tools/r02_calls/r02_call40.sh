#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "not opt_in" > $O/c40_tests.log 2>&1; echo "conv tests rc=$?"; tail -3 $O/c40_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c40_bench.json 2> $O/c40.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c40_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), d['gpu_launches'], d['roofline']['frac'])
PY
