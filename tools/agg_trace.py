"""aggregate gpurun_out/trace_step_list.txt by kernel name"""
import collections
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_step_list.txt"
agg = collections.defaultdict(lambda: [0.0, 0])
span = 0.0
for l in open(path):
    m = re.match(r"\s*([\d.]+) gap=\s*([-\d.]+) dur=\s*([\d.]+)\s+(.*)", l)
    t, dur, name = float(m[1]), float(m[3]), m[4]
    span = max(span, t + dur)
    name = name.replace("void ", "").replace("(anonymous namespace)::", "").replace("sap3d::", "")
    name = name.split("(")[0] if name.startswith(("conv_tc", "wgrad", "flash")) else re.split(r"[<(]", name)[0]
    agg[name][0] += dur
    agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"span {span / 1e3:.3f} ms, busy sum {tot / 1e3:.3f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v[0] / 1e3:8.3f} ms {100 * v[0] / tot:5.1f}% n={v[1]:4d} avg={v[0] / v[1]:7.1f} us  {k}")
