"""In-graph time per kernel category of the training step: captures the step as CUDA graphs with only ONE category of
ABI entry points enabled (the others return 0 without launching) and times the replay.  Unlike the ncu launch list
(cold cache, serialised) these are warm, back-to-back, in-graph durations; their sum against the full step shows
the overlap / gaps.   python tools/ablate_step.py [--graph p3d_unetplusplus_ds] [--batch 8]"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sap3d_tensorflow_b200 as sp  # noqa: E402
from sap3d_tensorflow_b200 import _abi as A  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--graph", default="p3d_unetplusplus_ds")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=112)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--gn", action="store_true")
args = ap.parse_args()

CATS = {
    "conv_fwd": ["sap3d_conv_fwd"], "conv_dgrad": ["sap3d_conv_dgrad"], "conv_wgrad": ["sap3d_conv_wgrad"],
    "bn_finalize": ["sap3d_bn_finalize"], "affine_act": ["sap3d_affine_act"], "affine_act_bwd": ["sap3d_affine_act_bwd"],
    "gn_fwd": ["sap3d_sample_channel_partials", "sap3d_gn_finalize", "sap3d_cbam_fwd", "sap3d_cbam_merge", "sap3d_concat_channels"],
    "gn_bwd": ["sap3d_gn_act_bwd", "sap3d_cbam_tail_bwd", "sap3d_split_channels"],
    "maxpool_fwd": ["sap3d_maxpool3d_fwd"], "maxpool_bwd": ["sap3d_maxpool3d_bwd"],
    "head+loss": ["sap3d_head_fwd", "sap3d_head_bwd", "sap3d_loss_smooth_l1"],
    "attn_gemm": ["sap3d_gemm_nt", "sap3d_gemm_tn"], "attn_softmax": ["sap3d_softmax_rows", "sap3d_softmax_bwd_rows"],
    "attn_glue": ["sap3d_transpose", "sap3d_pad_channels", "sap3d_cast", "sap3d_gate_fwd", "sap3d_gate_bwd", "sap3d_attention_fwd",
                  "sap3d_attention_bwd"],
    "dropout": ["sap3d_dropout"], "adam": ["sap3d_adam_step", "sap3d_step_increment"], "pack": ["sap3d_pack_multi"],
}
enabled = set()
real = {}
small_only = {"v": None}   # None: all convs; True: only small; False: only large
FLOP_SPLIT = 20e9


def conv_flops(d):
    d = d.contents if hasattr(d, "contents") else d._obj
    cin = d.cin[0] + d.cin[1]
    if d.transposed:
        pos = d.N * d.D * d.H * d.W
    else:
        pos = d.N * -(-d.D // d.sd) * -(-d.H // d.sh) * -(-d.W // d.sw)
    return 2.0 * pos * d.kd * d.kh * d.kw * cin * d.cout


def make_wrapper(name, fn):
    def w(*a):
        if name not in enabled:
            return 0
        if small_only["v"] is not None and name.startswith("sap3d_conv_"):
            if (conv_flops(a[0]) < FLOP_SPLIT) != small_only["v"]:
                return 0
        return fn(*a)
    return w


for cat, names in CATS.items():
    for n in names:
        real[n] = getattr(A.lib, n)
        setattr(A.lib, n, make_wrapper(n, real[n]))
all_names = set(real)
enabled |= all_names

B, size = args.batch, args.size
xin = sp.placeholder([B, 16, size, size, 3], dtype="bf16", training_graph=True)
if args.gn:
    from sap3d_tensorflow_b200.gn import p3d_gn
    head = getattr(p3d_gn, args.graph)(xin, 0.5, B, True)
else:
    head = getattr(sp.p3d, args.graph)(xin, 0.5, B, True)
sess = sp.Session(head)
x = torch.randn(B, 16, size, size, 3, device="cuda")
y = torch.rand(B, 16, size, size, device="cuda")
for _ in range(2):
    sess.train_step(x, y)
torch.cuda.synchronize()


def time_graphs(label):
    sess.graph_train = None
    sess.capture(train=True)
    graphs = [sess.graph_train[0], *sess.graph_train[2], sess.graph_train[1]]
    for _ in range(2):
        for g in graphs:
            g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.iters):
        for g in graphs:
            g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"{label:28s} {ms:8.3f} ms", flush=True)
    return ms


res = {"full": time_graphs("full step")}
enabled.clear()
res["empty (fills/memsets only)"] = time_graphs("empty (fills/memsets only)")
for cat, names in CATS.items():
    enabled.clear()
    enabled |= set(names)
    if cat.startswith("conv_"):
        for lab, v in (("small", True), ("large", False)):
            small_only["v"] = v
            res[f"{cat}/{lab}"] = time_graphs(f"{cat}/{lab}")
        small_only["v"] = None
    else:
        res[cat] = time_graphs(cat)
tot = sum(v for k, v in res.items() if k not in ("full",)) - (len(res) - 2) * res["empty (fills/memsets only)"]
print(f"sum of categories (minus repeated fills): {tot:.3f} ms vs full {res['full']:.3f} ms")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/ablate_step.json", "w"), indent=1)
