"""aggregate a tools/trace_step.py *_list.txt by kernel name:  python tools/agg_trace_list.py gpurun_out/x_list.txt [top]"""
import collections
import re
import sys

agg = collections.defaultdict(lambda: [0.0, 0])
tot = 0.0
span = 0.0
for l in open(sys.argv[1]):
    m = re.match(r'\s*([\d.]+) gap=\s*(-?[\d.]+) dur=\s*([\d.]+)\s+(.*)', l)
    if not m:
        continue
    d = float(m.group(3))
    nm = m.group(4).replace('void ', '').replace('(anonymous namespace)::', '').replace('sap3d::', '')
    nm = re.split(r'\((?!anonymous)', nm)[0][:70]
    agg[nm][0] += d
    agg[nm][1] += 1
    tot += d
    span = max(span, float(m.group(1)) + d)
print(f"span {span / 1e3:.3f} ms, busy sum {tot / 1e3:.3f} ms, {sum(c for _, c in agg.values())} activities")
for nm, (d, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{d / 1e3:8.3f} ms {100 * d / tot:5.1f}% n={c:4d} avg={d / c:7.1f}us {nm}")
