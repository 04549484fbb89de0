"""GPU probe: conv family (tcgen05 + SIMT) against the torch oracle.  Usage (on a B200 box):
    python tests/probe_conv.py > gpurun_out/probe_conv.log 2>&1
Development aid; the real parity tests live in tests/ (pytest -m gpu).
"""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tf_semantics as tfs  # noqa: E402
from sap3d_tensorflow_b200 import _abi as A  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
stream = torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item(), (a - b).abs().max().item()


def run_case(name, N, D, H, W, cin, cout, k, s, transposed=False, bias=True, impl=A.IMPL_AUTO, dtype=A.BF16,
             test_bwd=True, timeit=False):
    tdt = torch.bfloat16 if dtype == A.BF16 else torch.float32
    cin_t = sum(cin)
    xs = [torch.randn(N, D, H, W, c, device=dev).to(tdt) for c in cin]
    if transposed:
        w = torch.randn(*k, cout, cin_t, device=dev) / (cin_t * k[0] * k[1] * k[2]) ** 0.5
    else:
        w = torch.randn(*k, cin_t, cout, device=dev) / (cin_t * k[0] * k[1] * k[2]) ** 0.5
    b = torch.randn(cout, device=dev) if bias else None
    d = A.make_conv_desc(dtype, N, D, H, W, cin, cout, k, s, transposed, bias, False, impl)
    Do, Ho, Wo = A.conv_out_dims(d)
    y = torch.full((N, Do, Ho, Wo, cout), float("nan"), device=dev, dtype=tdt)
    rows = A.lib.sap3d_conv_stats_rows(C.byref(d))
    stats = torch.zeros(rows, 2, cout, device=dev)
    nf = A.lib.sap3d_conv_packed_elems(C.byref(d), 0)
    nd = A.lib.sap3d_conv_packed_elems(C.byref(d), 1)
    wf = torch.zeros(nf, device=dev, dtype=torch.bfloat16)
    wd = torch.zeros(nd, device=dev, dtype=torch.bfloat16)
    A.check(A.lib.sap3d_conv_pack_weights(C.byref(d), A.ptr(w), A.ptr(wf), A.ptr(wd), stream), "pack")
    x1 = xs[1] if len(xs) > 1 else None
    A.check(A.lib.sap3d_conv_fwd(C.byref(d), A.ptr(xs[0]), A.ptr(x1), A.ptr(w), A.ptr(wf), A.ptr(b), A.ptr(y),
                                 A.ptr(stats), stream), "fwd")
    torch.cuda.synchronize()
    # oracle on the same (rounded) inputs, fp32 math; weights rounded to bf16 for the TC path comparison
    xcat = torch.cat([t.float() for t in xs], dim=-1).requires_grad_(True)
    wq = (w.to(tdt).float() if dtype == A.BF16 else w).clone().requires_grad_(True)
    if transposed:
        ref = tfs.conv3d_transpose_same(xcat, wq, s, b)
    else:
        ref = tfs.conv3d_same(xcat, wq, s, b)
    r, m = rel(y, ref)
    sref1 = ref.sum(dim=(0, 1, 2, 3))
    sref2 = (ref * ref).sum(dim=(0, 1, 2, 3))
    s1 = stats[:, 0].double().sum(0).float()
    s2 = stats[:, 1].double().sum(0).float()
    rs1 = ((s1 - sref1).norm() / (sref1.norm() + 1e-20)).item()
    rs2 = ((s2 - sref2).norm() / (sref2.norm() + 1e-20)).item()
    ok = r < 1e-2 and rs2 < 1e-2
    line = f"{name:34s} fwd rel={r:.2e} max={m:.2e} stats1={rs1:.1e} stats2={rs2:.1e} rows={rows} nan={int(torch.isnan(y.float()).sum())}"
    if test_bwd:
        dy = torch.randn_like(ref).to(tdt)
        ref.backward(dy.float())
        # dgrad per segment
        off = 0
        for si, c in enumerate(cin):
            dx = torch.full_like(xs[si], float("nan"))
            A.check(A.lib.sap3d_conv_dgrad(C.byref(d), si, A.ptr(dy), A.ptr(w), A.ptr(wd), A.ptr(dx), 0, stream), "dgrad")
            torch.cuda.synchronize()
            rr, mm = rel(dx, xcat.grad[..., off:off + c])
            ok = ok and rr < 1e-2
            line += f" | dgrad{si} rel={rr:.2e}"
            # accumulate variant
            dx2 = dx.clone()
            A.check(A.lib.sap3d_conv_dgrad(C.byref(d), si, A.ptr(dy), A.ptr(w), A.ptr(wd), A.ptr(dx2), 1, stream), "dgrad acc")
            torch.cuda.synchronize()
            rr2, _ = rel(dx2, 2 * xcat.grad[..., off:off + c])
            ok = ok and rr2 < 2e-2
            off += c
        dw = torch.zeros_like(w)
        db = torch.zeros(cout, device=dev)
        A.check(A.lib.sap3d_conv_wgrad(C.byref(d), A.ptr(xs[0]), A.ptr(x1), A.ptr(dy), A.ptr(dw), A.ptr(db), None, stream), "wgrad")
        torch.cuda.synchronize()
        rw, _ = rel(dw, wq.grad)
        rb, _ = rel(db, dy.float().sum(dim=(0, 1, 2, 3)))
        ok = ok and rw < 1e-2 and rb < 1e-2
        line += f" | wgrad rel={rw:.2e} db={rb:.1e}"
    if timeit:
        for _ in range(3):
            A.lib.sap3d_conv_fwd(C.byref(d), A.ptr(xs[0]), A.ptr(x1), A.ptr(w), A.ptr(wf), A.ptr(b), A.ptr(y), A.ptr(stats), stream)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            A.lib.sap3d_conv_fwd(C.byref(d), A.ptr(xs[0]), A.ptr(x1), A.ptr(w), A.ptr(wf), A.ptr(b), A.ptr(y), A.ptr(stats), stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        taps = k[0] * k[1] * k[2]
        pos = N * D * H * W if transposed else N * Do * Ho * Wo
        fl = 2.0 * pos * taps * cin_t * cout
        line += f" | {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s"
        dy = torch.randn_like(y)
        dw = torch.zeros_like(w)
        dxs = [torch.empty_like(t) for t in xs]
        for what in ("dgrad", "wgrad"):
            def run():
                if what == "dgrad":
                    for si in range(len(cin)):
                        A.lib.sap3d_conv_dgrad(C.byref(d), si, A.ptr(dy), A.ptr(w), A.ptr(wd), A.ptr(dxs[si]), 0, stream)
                else:
                    A.lib.sap3d_conv_wgrad(C.byref(d), A.ptr(xs[0]), A.ptr(x1), A.ptr(dy), A.ptr(dw), None, None, stream)
            for _ in range(2):
                run()
            e0.record()
            for _ in range(5):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            line += f" | {what} {ms:.3f} ms {fl / ms / 1e9:.1f} TF/s"
    print(("PASS " if ok else "FAIL ") + line, flush=True)
    return ok


def main():
    print("device ok:", A.lib.sap3d_device_ok(), torch.cuda.get_device_name(0), flush=True)
    allok = True
    T = A.IMPL_TC
    cases = [
        # name, N,D,H,W, cin, cout, k, s, transposed
        ("tc 1x1x1 flat 64->64", 2, 4, 8, 8, [64], 64, (1, 1, 1), (1, 1, 1), False),
        ("tc 1x1x1 flat 256->64", 2, 8, 28, 28, [256], 64, (1, 1, 1), (1, 1, 1), False),
        ("tc 1x1x1 64->256", 2, 8, 28, 28, [64], 256, (1, 1, 1), (1, 1, 1), False),
        ("tc 1x1x1 s(1,2,2) 256->128", 2, 4, 28, 28, [256], 128, (1, 1, 1), (1, 2, 2), False),
        ("tc S 1x3x3 64", 2, 8, 28, 28, [64], 64, (1, 3, 3), (1, 1, 1), False),
        ("tc T 3x1x1 64", 2, 8, 28, 28, [64], 64, (3, 1, 1), (1, 1, 1), False),
        ("tc T 3x1x1 256 D=2", 2, 2, 7, 7, [256], 256, (3, 1, 1), (1, 1, 1), False),
        ("tc S 1x3x3 256 7x7", 2, 2, 7, 7, [256], 256, (1, 3, 3), (1, 1, 1), False),
        ("tc 3x3x3 concat 64+128->128", 1, 8, 56, 56, [64, 128], 128, (3, 3, 3), (1, 1, 1), False),
        ("tc 3x3x3 concat 256+256->256", 1, 4, 28, 28, [256, 256], 256, (3, 3, 3), (1, 1, 1), False),
        ("tc 2x3x3 concat 512+512->512", 1, 2, 14, 14, [512, 512], 512, (2, 3, 3), (1, 1, 1), False),
        ("tc deconv 3x3x3 s2 256->128", 1, 4, 28, 28, [256], 128, (3, 3, 3), (2, 2, 2), True),
        ("tc deconv 2x3x3 s2 512->256", 1, 2, 14, 14, [512], 256, (2, 3, 3), (2, 2, 2), True),
        ("tc deconv 1x3x3 s2 1024->512", 1, 1, 7, 7, [1024], 512, (1, 3, 3), (2, 2, 2), True),
        ("tc 1x1x1 128->16 (attn f)", 1, 8, 56, 56, [128], 16, (1, 1, 1), (1, 1, 1), False),
    ]
    for c in cases:
        try:
            allok &= run_case(*c, impl=T)
        except Exception as e:  # noqa: BLE001
            allok = False
            print(f"ERROR {c[0]}: {e}", flush=True)
    simt = [
        ("simt stem 1x7x7 s2 3->64", 1, 4, 32, 32, [3], 64, (1, 7, 7), (1, 2, 2), False),
        ("simt 3x3x3 concat 8+16->24", 1, 3, 9, 10, [8, 16], 24, (3, 3, 3), (1, 1, 1), False),
        ("simt deconv 3x3x3 s2 16->1", 1, 4, 10, 10, [16], 1, (3, 3, 3), (2, 2, 2), True),
        ("simt deconv 3x3x3 s4 16->8", 1, 1, 5, 5, [16], 8, (3, 3, 3), (4, 4, 4), True),
        ("simt deconv 3 s1 16->8", 1, 3, 6, 6, [16], 8, (3, 3, 3), (1, 1, 1), True),
        ("simt 2x3x3 8->8", 2, 2, 7, 7, [8], 8, (2, 3, 3), (1, 1, 1), False),
    ]
    for c in simt:
        for dt in (A.F32, A.BF16):
            try:
                allok &= run_case(c[0] + (" f32" if dt == A.F32 else " bf16"), *c[1:], impl=A.IMPL_SIMT, dtype=dt)
            except Exception as e:  # noqa: BLE001
                allok = False
                print(f"ERROR {c[0]}: {e}", flush=True)
    # timing of the FLOP-dominant decoder shapes (B=8)
    big = [
        ("time x_1_2 3x3x3 128+128->128 B8", 8, 8, 56, 56, [128, 128], 128, (3, 3, 3), (1, 1, 1), False),
        ("time x_2_2 3x3x3 256+256->256 B8", 8, 4, 28, 28, [256, 256], 256, (3, 3, 3), (1, 1, 1), False),
        ("time x_3_1 2x3x3 512+512->512 B8", 8, 2, 14, 14, [512, 512], 512, (2, 3, 3), (1, 1, 1), False),
        ("time upx_2 deconv 256->128 B8", 8, 4, 28, 28, [256], 128, (3, 3, 3), (2, 2, 2), True),
        ("time stage1 S 64 B8", 8, 8, 28, 28, [64], 64, (1, 3, 3), (1, 1, 1), False),
        ("time stage3 expand 256->1024 B8", 8, 2, 7, 7, [256], 1024, (1, 1, 1), (1, 1, 1), False),
    ]
    for c in big:
        try:
            allok &= run_case(*c, impl=T, test_bwd=False, timeit=True)
        except Exception as e:  # noqa: BLE001
            allok = False
            print(f"ERROR {c[0]}: {e}", flush=True)
    print("ALL PASS" if allok else "SOME FAILED", flush=True)
    return 0 if allok else 1


if __name__ == "__main__":
    t0 = time.time()
    rc = main()
    print(f"elapsed {time.time() - t0:.1f}s")
    sys.exit(rc)
