#!/bin/bash
# round-2 GPU call 3: DSMEM bulk-copy split-K exchange + programmatic dependent launch (A/B), new parity / surface tests
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q > $O/c3_conv.log 2>&1; echo "conv tests rc=$?"; tail -3 $O/c3_conv.log
timeout 120 python tools/conv_phase_probe.py > $O/c3_conv_phase_probe.txt 2>&1; echo "phase probe rc=$?"
grep -E "^==|cluster sync|acc ready|epilogue done" $O/c3_conv_phase_probe.txt | head -40
SAP3D_PDL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c3_bench_nopdl.json 2> $O/c3_bench_nopdl.err; echo "bench nopdl rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c3_bench_pdl.json 2> $O/c3_bench_pdl.err; echo "bench pdl rc=$?"
tail -3 $O/c3_bench_pdl.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c3_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['extra']['infer_ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 900 python -m pytest tests/test_zz_network_surface_gpu.py tests/test_zz_config_parity_gpu.py -q -s > $O/c3_parity.log 2>&1; echo "parity rc=$?"
grep -E "^\S.*training step|oracle|CUDA|passed|failed|Error|error|assert" $O/c3_parity.log | cut -c1-400 | head -60
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -x -q > $O/c3_model.log 2>&1; echo "model/ops tests rc=$?"; tail -3 $O/c3_model.log
python tools/trace_step.py --out $O/c3_trace_train.txt > /dev/null 2> $O/c3_trace.err; echo "trace rc=$?"
