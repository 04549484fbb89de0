"""Per-launch time of the backbone's BatchNorm kernels in a dependent chain replayed from a CUDA graph (what the training step
sees): forward bn_apply_fused and the backward (slab kernel, or the cooperative kernel with SAP3D_BN_BWD_SLAB=0) on stage-2 /
stage-3 tensor shapes.   python tools/bn_chain_probe.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sap3d_tensorflow_b200 import _abi as A  # noqa: E402

dev = "cuda"
CASES = [("stage3 inner [8,2,7,7,256]", (8, 2, 7, 7, 256)), ("stage3 tail [8,2,7,7,1024]", (8, 2, 7, 7, 1024)),
         ("stage2 inner [8,4,14,14,128]", (8, 4, 14, 14, 128)), ("stage2 tail [8,4,14,14,512]", (8, 4, 14, 14, 512))]


def chain(fn, n=50):
    for _ in range(3):
        fn(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=torch.cuda.Stream()):
        for _ in range(n):
            fn(torch.cuda.current_stream().cuda_stream)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (4 * n)


print("SAP3D_BN_BWD_SLAB =", os.environ.get("SAP3D_BN_BWD_SLAB", "1"), " SAP3D_PDL =", os.environ.get("SAP3D_PDL", "1"))
for name, shape in CASES:
    N, D, H, W, Cc = shape
    P = N * D * H * W
    bf = torch.bfloat16
    a = torch.randn(shape, device=dev).to(bf)
    b = torch.randn(shape, device=dev).to(bf)
    dy = torch.randn(shape, device=dev).to(bf)
    da, db = torch.empty_like(a), torch.empty_like(a)
    f = lambda: torch.rand(Cc, device=dev) + 0.5  # noqa: E731
    s1, t1, m1, r1, g1, b1 = f(), f(), f(), f(), f(), f()
    dg, dbt = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
    ws = torch.zeros(A.lib.sap3d_affine_act_bwd_workspace(Cc) // 4 + 16, device=dev)
    for tag, use_b in (("relu(bn(a))", False), ("relu(bn(a) + b)", True)):
        def bwd(st):
            A.check(A.lib.sap3d_affine_act_bwd(A.BF16, A.ptr(dy), A.ptr(a), A.ptr(s1), A.ptr(t1), A.ptr(m1), A.ptr(r1), 0 if use_b else 1,
                                               A.ptr(b) if use_b else None, None, None, None, None, 0, 1 if use_b else 0, P, Cc, A.ptr(da), 0,
                                               A.ptr(db) if use_b else None, 0, A.ptr(dg), A.ptr(dbt), None, None, A.ptr(ws), st), "bwd")
        print(f"{name:32s} backward {tag:18s}: {chain(bwd):6.2f} us per launch")
print("done")
