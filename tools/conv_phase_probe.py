"""Where do the ~12 us of a small backbone convolution go?  Per-CTA phase time stamps of conv_tc_kernel (clock64 +
globaltimer written by the kernel itself, sap3d_debug_conv_timing) for the stage-2 / stage-3 layer shapes, next to the
per-launch time of a dependent chain of the same launch replayed from a CUDA graph (what the training step sees).

    python tools/conv_phase_probe.py > gpurun_out/conv_phase_probe.txt
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sap3d_tensorflow_b200 import _abi as A  # noqa: E402

dev = torch.device("cuda")
PH = ["entry", "setup done", "first TMA issued", "first stage full (MMA starts)", "last MMA issued", "acc ready (split: ship)",
      "partials received", "epilogue done", "exit", "last TMA issued", "acc ready (no split)", "partials staged (split)",
      "peers ready (split)", "epilogue: last chunk loaded + partials added", "epilogue: last chunk statistics done", "epilogue: stores issued"]

CASES = [
    # name, N, D, H, W, cin, cout, kernel
    ("stage3 conv1 1x1x1 1024->256", 8, 2, 7, 7, 1024, 256, (1, 1, 1)),
    ("stage3 convS 1x3x3 256->256", 8, 2, 7, 7, 256, 256, (1, 3, 3)),
    ("stage3 convT 3x1x1 256->256", 8, 2, 7, 7, 256, 256, (3, 1, 1)),
    ("stage3 conv3 1x1x1 256->1024", 8, 2, 7, 7, 256, 1024, (1, 1, 1)),
    ("stage2 conv1 1x1x1 512->128", 8, 4, 14, 14, 512, 128, (1, 1, 1)),
    ("stage2 convS 1x3x3 128->128", 8, 4, 14, 14, 128, 128, (1, 3, 3)),
    ("stage2 conv3 1x1x1 128->512", 8, 4, 14, 14, 128, 512, (1, 1, 1)),
    ("stage1 convS 1x3x3 64->64", 8, 8, 28, 28, 64, 64, (1, 3, 3)),
]


def main():
    st = torch.cuda.current_stream().cuda_stream
    dbg = torch.zeros(1024 * 16 * 2, device=dev, dtype=torch.int64)
    for name, N, D, H, W, cin, cout, k in CASES:
        x = torch.randn(N, D, H, W, cin, device=dev).to(torch.bfloat16)
        w = torch.randn(*k, cin, cout, device=dev) * 0.02
        b = torch.zeros(cout, device=dev)
        d = A.make_conv_desc(A.BF16, N, D, H, W, [cin], cout, k, (1, 1, 1), False, True, False, A.IMPL_TC)
        y = torch.empty(N, D, H, W, cout, device=dev, dtype=torch.bfloat16)
        rows = A.lib.sap3d_conv_stats_rows(C.byref(d))
        stats = torch.zeros(rows, 2, cout, device=dev)
        wf = torch.zeros(A.lib.sap3d_conv_packed_elems(C.byref(d), 0), device=dev, dtype=torch.bfloat16)
        A.check(A.lib.sap3d_conv_pack_weights(C.byref(d), A.ptr(w), A.ptr(wf), None, st), "pack")

        def launch(s=st):
            A.check(A.lib.sap3d_conv_fwd(C.byref(d), A.ptr(x), None, A.ptr(w), A.ptr(wf), A.ptr(b), A.ptr(y), A.ptr(stats), s), "conv")

        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        # (1) dependent chain of 50 launches replayed from a CUDA graph
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream()
        with torch.cuda.graph(g, stream=cap):
            for _ in range(50):
                launch(torch.cuda.current_stream().cuda_stream)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        chain_us = e0.elapsed_time(e1) * 1e3 / 200
        # (2) one launch with the phase probe
        dbg.zero_()
        A.check(A.lib.sap3d_debug_conv_timing(A.ptr(dbg)), "dbg on")
        launch()
        torch.cuda.synchronize()
        A.check(A.lib.sap3d_debug_conv_timing(None), "dbg off")
        t = dbg.view(-1, 16, 2).cpu()
        live = t[:, 0, 1] != 0
        t = t[live]
        ncta = t.shape[0]
        if ncta == 0:
            print(f"\n== {name}: chain of 50 in a graph: {chain_us:.2f} us per launch (persistent kernel: no phase probe)")
            continue
        clk, glb = t[:, :, 0].double(), t[:, :, 1].double()
        g0 = glb[:, 0].min()
        span = (glb[:, 8].max() - g0) / 1e3
        skew = (glb[:, 0].max() - g0) / 1e3
        # clock rate from the longest-lived CTA
        dt_clk = (clk[:, 8] - clk[:, 0])
        dt_glb = (glb[:, 8] - glb[:, 0])
        ghz = float((dt_clk / dt_glb.clamp(min=1)).median())
        print(f"\n== {name}: M={N * D * H * W} K={k[0] * k[1] * k[2] * cin} N={cout}; {ncta} CTAs; chain of 50 in a graph: {chain_us:.2f} us per launch; "
              f"probe launch: first entry -> last exit {span:.2f} us, entry skew {skew:.2f} us, SM clock ~{ghz:.2f} GHz")
        order = [0, 1, 2, 3, 9, 4, 5, 10, 11, 12, 6, 13, 14, 15, 7, 8]
        for i in order:
            v = clk[:, i] - clk[:, 0]
            ok = clk[:, i] != 0
            if ok.sum() == 0:
                continue
            us = v[ok] / (ghz * 1e3)
            print(f"   {PH[i]:32s} median {float(us.median()):6.2f} us   min {float(us.min()):6.2f}   max {float(us.max()):6.2f}   (since CTA entry, {int(ok.sum())} CTAs)")
    print("done")


if __name__ == "__main__":
    main()
