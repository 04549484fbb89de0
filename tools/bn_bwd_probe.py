"""Stand-alone timing of the big-tensor BatchNorm backward (reduce + finalize + apply, the `nob` kernels) at the decoder's sizes:
CUDA events, L2 flushed between calls.  SAP3D_NOB_INFLIGHT=4|8 selects the loads-in-flight depth.
    python tools/bn_bwd_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sap3d_tensorflow_b200 import _abi as A  # noqa: E402

dev = torch.device("cuda")
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
for name, shape in (("level 0: 8x16x112x112x64", (8, 16, 112, 112, 64)), ("level 1: 8x8x56x56x128", (8, 8, 56, 56, 128)),
                    ("level 2: 8x4x28x28x256", (8, 4, 28, 28, 256))):
    C = shape[-1]
    P = 1
    for d in shape[:-1]:
        P *= d
    a = torch.randn(*shape, device=dev).to(torch.bfloat16)
    dy = torch.randn(*shape, device=dev).to(torch.bfloat16)
    da = torch.empty_like(a)
    s1, t1 = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    m1, r1 = torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev) + 0.5
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    ws = torch.zeros(A.lib.sap3d_affine_act_bwd_workspace(C) // 4 + 16, device=dev)

    def call():
        A.check(A.lib.sap3d_affine_act_bwd(A.BF16, A.ptr(dy), A.ptr(a), A.ptr(s1), A.ptr(t1), A.ptr(m1), A.ptr(r1), 1, None, None, None, None, None, 0, 0,
                                           P, C, A.ptr(da), 0, None, 0, A.ptr(dg), A.ptr(db), None, None, A.ptr(ws), st), "bwd")

    for _ in range(3):
        call()
    tot, iters = 0.0, 10
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / iters
    nbytes = P * C * 2 * 5   # reduce reads dy, a; apply reads dy, a and writes da
    print(f"{name}: {ms * 1e3:.1f} us for reduce + finalize + apply, {nbytes / ms / 1e9:.2f} TB/s over 5 tensor passes "
          f"(SAP3D_NOB_INFLIGHT={os.environ.get('SAP3D_NOB_INFLIGHT', 'default 8')})")
