#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "affine_act or bn_backward" > $O/c8_ops.log 2>&1; echo "bn ops tests rc=$?"; tail -2 $O/c8_ops.log
timeout 120 python tools/bn_chain_probe.py > $O/c8_bn_chain_slab.txt 2>&1; head -6 $O/c8_bn_chain_slab.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/c8_bench.json 2> $O/c8_bench.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c8_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['extra']['infer_ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 900 python -m pytest tests/test_metrics.py tests/test_zz_network_surface_gpu.py -x -q -m gpu > $O/c8_new_tests.log 2>&1; echo "new tests rc=$?"; tail -6 $O/c8_new_tests.log
python tools/profile_metrics.py > $O/c8_plain_metrics.log 2>&1 && timeout 200 ncu --profile-from-start off --set full --clock-control none -k 'regex:metrics_kernel' -c 2 -f -o $O/r02_bw_metrics python tools/profile_metrics.py > $O/c8_ncu_metrics.log 2>&1
ncu -i $O/r02_bw_metrics.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]
for r in rows[2:]:
    g=lambda n: r[h.index(n)]+' '+u[h.index(n)]
    print(r[h.index('Kernel Name')][:40], g('gpu__time_duration.sum'), g('dram__bytes_read.sum'), g('dram__bytes_write.sum'), g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'))
"
