"""Per-kernel device durations INSIDE the replayed CUDA graphs of one training step (torch.profiler / CUPTI activity
records): warm, in-order, with real inter-kernel gaps — complements the ncu launch list (cold, serialised).
    python tools/trace_step.py [--graph NAME] [--batch 8] [--infer]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import sap3d_tensorflow_b200 as sp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--graph", default="p3d_unetplusplus_ds")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=112)
ap.add_argument("--infer", action="store_true")
ap.add_argument("--gn", action="store_true")
ap.add_argument("--out", default="gpurun_out/trace_step.txt")
args = ap.parse_args()
B, size = args.batch, args.size
xin = sp.placeholder([B, 16, size, size, 3], dtype="bf16", training_graph=not args.infer)
if args.gn:
    from sap3d_tensorflow_b200.gn import p3d_gn
    head = getattr(p3d_gn, args.graph)(xin, 0.5, B, not args.infer)
else:
    head = getattr(sp.p3d, args.graph)(xin, 0.5, B, not args.infer)
sess = sp.Session(head)
x = torch.randn(B, 16, size, size, 3, device="cuda")
y = torch.rand(B, 16, size, size, device="cuda")
step = (lambda: sess.run(x, graph=True)) if args.infer else (lambda: sess.train_step(x, y, graph=True))
for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
n = len(evs) // 3
last = evs[2 * n:]                      # third replay
t0, t1 = last[0].time_range.start, last[-1].time_range.end
agg = collections.defaultdict(lambda: [0.0, 0])
busy = 0.0
for e in last:
    d = e.time_range.end - e.time_range.start
    nm = e.name.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("sap3d::", "").split("(")[0]
    agg[nm][0] += d
    agg[nm][1] += 1
    busy += d
lines = [f"{args.graph} B={B} {'infer' if args.infer else 'train'}: span {(t1 - t0) / 1e3:.3f} ms, kernel-busy sum {busy / 1e3:.3f} ms, {len(last)} device activities"]
for nm, (d, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    lines.append(f"{d / 1e3:9.3f} ms {100 * d / busy:6.2f}% n={c:5d} avg={d / c:8.1f} us  {nm[:90]}")
os.makedirs(os.path.dirname(args.out), exist_ok=True)
open(args.out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
# the raw in-order list, for locating a specific layer
with open(args.out.replace(".txt", "_list.txt"), "w") as f:
    prev = t0
    for e in last:
        f.write(f"{(e.time_range.start - t0):10.1f} gap={(e.time_range.start - prev):7.1f} dur={(e.time_range.end - e.time_range.start):8.1f}  {e.name[:110]}\n")
        prev = e.time_range.end
# the same list with the CUDA stream of every activity (from the chrome-trace export): "<start us> <dur us> <stream> <name>"
try:
    import json
    jpath = args.out.replace(".txt", "_chrome.json")
    prof.export_chrome_trace(jpath)
    tr = json.load(open(jpath))
    ks = [ev for ev in tr["traceEvents"] if ev.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in ev]
    ks.sort(key=lambda ev: ev["ts"])
    n3 = len(ks) // 3
    lastk = ks[2 * n3:]
    k0 = lastk[0]["ts"]
    with open(args.out.replace(".txt", "_streams.txt"), "w") as f:
        for ev in lastk:
            f.write(f"{ev['ts'] - k0:10.1f} {ev.get('dur', 0):8.1f} {ev.get('args', {}).get('stream', -1):4d} {ev['name'][:100]}\n")
    os.remove(jpath)
except Exception as ex:  # noqa: BLE001  (developer tool: the aggregate above is the product)
    print("stream list not written:", ex)

