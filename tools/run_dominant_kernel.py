"""Runs the FLOP-dominant launch (x_1_2 / x_1_3 decoder conv: 3x3x3, 128+128 -> 128 at B x 8 x 56 x 56) a few
times, for `ncu --set full -k regex:conv_tc_kernel`.  Optional arg: which = fwd | wgrad | dgrad | dgrad2.  Also prints the
stand-alone time per call (CUDA events, L2 flushed)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sap3d_tensorflow_b200 import _abi as A  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
N, D, H, W = 8, 8, 56, 56
dev = torch.device("cuda")
torch.manual_seed(0)
xs = [torch.randn(N, D, H, W, 128, device=dev).to(torch.bfloat16) for _ in range(2)]
w = torch.randn(3, 3, 3, 256, 128, device=dev) * 0.02
b = torch.zeros(128, device=dev)
d = A.make_conv_desc(A.BF16, N, D, H, W, [128, 128], 128, (3, 3, 3), (1, 1, 1), False, True, False, A.IMPL_TC)
y = torch.empty(N, D, H, W, 128, device=dev, dtype=torch.bfloat16)
dy = torch.randn(N, D, H, W, 128, device=dev).to(torch.bfloat16)
dx = torch.empty_like(xs[0])
dw = torch.zeros_like(w)
stats = torch.zeros(A.lib.sap3d_conv_stats_rows(C.byref(d)), 2, 128, device=dev)
wf = torch.zeros(A.lib.sap3d_conv_packed_elems(C.byref(d), 0), device=dev, dtype=torch.bfloat16)
wd = torch.zeros(A.lib.sap3d_conv_packed_elems(C.byref(d), 1), device=dev, dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
A.check(A.lib.sap3d_conv_pack_weights(C.byref(d), A.ptr(w), A.ptr(wf), A.ptr(wd), st), "pack")
dx1 = torch.empty_like(xs[1])


def launch():
    if which == "fwd":
        A.check(A.lib.sap3d_conv_fwd(C.byref(d), A.ptr(xs[0]), A.ptr(xs[1]), A.ptr(w), A.ptr(wf), A.ptr(b), A.ptr(y), A.ptr(stats), st), "fwd")
    elif which == "dgrad":
        A.check(A.lib.sap3d_conv_dgrad(C.byref(d), 0, A.ptr(dy), A.ptr(w), A.ptr(wd), A.ptr(dx), 0, st), "dgrad")
    elif which == "dgrad2":
        A.check(A.lib.sap3d_conv_dgrad2(C.byref(d), A.ptr(dy), A.ptr(w), A.ptr(wd), A.ptr(dx), 0, A.ptr(dx1), 0, st), "dgrad2")
    elif which == "wgrad":
        A.check(A.lib.sap3d_conv_wgrad(C.byref(d), A.ptr(xs[0]), A.ptr(xs[1]), A.ptr(dy), A.ptr(dw), None, None, st), "wgrad")
    else:
        raise SystemExit("which = fwd | dgrad | dgrad2 | wgrad")


for _ in range(4):
    launch()
torch.cuda.synchronize()
# timed alone, L2 flushed between launches (the bench's stand-alone figure)
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
tot, iters = 0.0, 20
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
ms = tot / iters
flops = 2.0 * N * D * H * W * 27 * 256 * 128 * (1.0 if which != "dgrad" else 0.5)
print("done %s: %.1f us per call, %.0f TFLOP/s, halo launches %d" % (which, ms * 1e3, flops / ms / 1e9, A.lib.sap3d_debug_conv_halo_launches()))
