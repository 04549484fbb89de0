#!/bin/bash
# round-2 GPU call 4: slab BN backward (A/B), epilogue-vector staging under the K loop, full GPU suite
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_conv_gpu.py -x -q > $O/c4_ops.log 2>&1; echo "ops+conv tests rc=$?"; tail -3 $O/c4_ops.log
timeout 120 python tools/conv_phase_probe.py > $O/c4_conv_phase_probe.txt 2>&1; echo "phase probe rc=$?"
grep -E "^==|staged|peers ready|partials received|acc ready|epilogue done" $O/c4_conv_phase_probe.txt | head -30
SAP3D_BN_BWD_SLAB=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c4_bench_noslab.json 2> $O/c4_bench_noslab.err; echo "bench noslab rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c4_bench.json 2> $O/c4_bench.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c4_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['extra']['infer_ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
python tools/trace_step.py --out $O/c4_trace_train.txt > /dev/null 2> $O/c4_trace.err; echo "trace rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > $O/c4_pytest.log 2>&1; echo "full gpu suite rc=$?"; tail -5 $O/c4_pytest.log
