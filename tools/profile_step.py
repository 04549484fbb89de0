"""One eager training step (and one forward) between cudaProfilerStart/Stop, for ncu launch lists:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sap3d_tensorflow_b200 as sp  # noqa: E402

graph = sys.argv[1] if len(sys.argv) > 1 else "p3d_unetplusplus_nonsa"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
size = int(sys.argv[3]) if len(sys.argv) > 3 else 112
mode = sys.argv[4] if len(sys.argv) > 4 else "train"
xin = sp.placeholder([B, 16, size, size, 3], dtype="bf16", training_graph=(mode == "train"))
if graph.startswith("gn:"):     # GroupNorm + CBAM graphs: gn:inference_p3d, gn:inference_p3d_decoder_block ...
    from sap3d_tensorflow_b200.gn import p3d_gn
    head = getattr(p3d_gn, graph[3:])(xin, 0.5, B, mode == "train")
else:
    head = getattr(sp.p3d, graph)(xin, 0.5, B, mode == "train")
sess = sp.Session(head)
x = torch.randn(B, 16, size, size, 3, device="cuda") * 0.3
y = torch.rand(B, 16, size, size, device="cuda")
for _ in range(2):
    if mode == "train":
        sess.train_step(x, y)
    else:
        sess.run(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
if mode == "train":
    sess.train_step(x, y)
else:
    sess.run(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", mode, graph, B, size)
