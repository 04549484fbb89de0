"""Summarises an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        v *= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e6
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    tot[name] += v
    cnt[name] += 1
    total += v
print(f"total {total/1e6:.3f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{v/1e6:10.3f} ms {100*v/total:6.2f}% n={cnt[k]:5d} avg={v/cnt[k]/1e3:9.1f} us  {k[:110]}")
