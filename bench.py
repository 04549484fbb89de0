#!/usr/bin/env python
"""bench.py — headline benchmark of the P3D saliency hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (ours; torchrun launches N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU restatement of the reference path)

Metric (BASELINE.json): clips/sec on synthetic 16x112x112 clips, bf16, training step
(P3D + saliency decoder, smooth-L1, Adam) at batch 8 per GPU (configs[1]); data-parallel over N GPUs
shards clips (weak scaling) with a bf16 gradient all-reduce between backward and the optimizer.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "clips/sec (16x112x112, bf16) train step"
UNIT = "clips/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--graph", default=None)
    ap.add_argument("--batch", type=int, default=8, help="clips per GPU")
    ap.add_argument("--size", type=int, default=112)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA-graph replay (debug)")
    return ap.parse_args()


def default_graph():
    import sap3d_tensorflow_b200 as sp

    return "p3d_unetplusplus_ds" if hasattr(sp.network, "attention") else "p3d_unetplusplus_nonsa"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [a.strip() for a in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        # keep the samples taken under load (upper half) for the median
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------
def cpu_reference_step(graph: str, size: int, batch: int, steps: int, warmup: int):
    """times the oracle's training step (CPU restatement of the reference graph; TensorFlow is not
    installable here) on all host cores; returns clips/s and description"""
    import torch
    from oracle import p3d_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = O.synthetic_clip(batch, 16, size, seed=0)
    y = O.synthetic_target(batch, 16, size, seed=1)
    vs = O.VarStore(seed=0)
    with torch.no_grad():
        O.forward(graph, x, vs, True)
    adam = {}
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(graph, x, y, vs, adam, i + 1)
        ts.append(time.perf_counter() - t0)
    ts = ts[warmup:]
    sec = sum(ts) / len(ts)
    return batch / sec, sec, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    graph = args.graph or default_graph()
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    val, sec, cores = cpu_reference_step(graph, args.size, 1, steps, warm)
    sample = f"{steps} training steps of {graph} on 1 clip 16x{args.size}x{args.size} (fp32, torch-CPU oracle, {cores} threads)"
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{graph} training step (fwd + smooth-L1 + bwd + Adam), batch 8 clips 16x{args.size}x{args.size} per GPU",
                   "note": "reference = CPU restatement of the reference TF graph (TensorFlow unavailable in this image)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def time_dominant_kernel(batch: int, size: int, iters: int = 20):
    """the FLOP-dominant launch: the x_1_2 / x_1_3 decoder conv (3x3x3, 128+128 -> 128 at 8 x size/2 x size/2),
    timed alone with CUDA events on the launching stream, L2 flushed between launches"""
    import ctypes as C

    import torch
    from sap3d_tensorflow_b200 import _abi as A

    dev = torch.device("cuda")
    N, D, H, W = batch, 8, size // 2, size // 2
    xs = [torch.randn(N, D, H, W, 128, device=dev).to(torch.bfloat16) for _ in range(2)]
    w = torch.randn(3, 3, 3, 256, 128, device=dev) * 0.02
    b = torch.zeros(128, device=dev)
    d = A.make_conv_desc(A.BF16, N, D, H, W, [128, 128], 128, (3, 3, 3), (1, 1, 1), False, True, False, A.IMPL_TC)
    y = torch.empty(N, D, H, W, 128, device=dev, dtype=torch.bfloat16)
    rows = A.lib.sap3d_conv_stats_rows(C.byref(d))
    stats = torch.zeros(rows, 2, 128, device=dev)
    wf = torch.zeros(A.lib.sap3d_conv_packed_elems(C.byref(d), 0), device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    A.check(A.lib.sap3d_conv_pack_weights(C.byref(d), A.ptr(w), A.ptr(wf), None, st), "pack")
    flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)

    def launch():
        A.check(A.lib.sap3d_conv_fwd(C.byref(d), A.ptr(xs[0]), A.ptr(xs[1]), A.ptr(w), A.ptr(wf), A.ptr(b), A.ptr(y), A.ptr(stats), st), "conv")

    for _ in range(3):
        launch()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / iters
    flops = 2.0 * N * D * H * W * 27 * 256 * 128
    return ms, flops


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import sap3d_tensorflow_b200 as sp
    from sap3d_tensorflow_b200 import parallel

    graph = args.graph or default_graph()
    B, size = args.batch, args.size
    dev = torch.device("cuda", local_rank)
    xin = sp.placeholder([B, 16, size, size, 3], dtype="bf16", training_graph=True, device=f"cuda:{local_rank}")
    head = getattr(sp.p3d, graph)(xin, 0.5, B, True)
    sess = sp.Session(head)
    if world > 1:
        parallel.attach_data_parallel(sess)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    x_host = ((torch.randint(0, 256, (B, 16, size, size, 3), generator=g).float() - torch.tensor([90.0, 102.0, 98.0])) / 255.0).pin_memory()
    y_host = (torch.randint(0, 256, (B, 16, size, size), generator=g).float() / 255.0).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    use_graph = not args.eager

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") ---------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        sess.train_step(x_dev, y_dev, graph=use_graph)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        sess.train_step(x_dev, y_dev, graph=use_graph)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(sess.eng.loss_buf.item())

    # ---- end-to-end timing: pinned host inputs -> H2D -> step -> D2H loss -------------------------
    for _ in range(2):
        sess.train_step(x_host, y_host, graph=use_graph)
        float(sess.eng.loss_buf.item())
    barrier()
    e0.record()
    for _ in range(args.steps):
        sess.train_step(x_host, y_host, graph=use_graph)
        float(sess.eng.loss_buf.item())
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps

    # ---- inference throughput (same engine, forward only, device resident) -----------------------
    for _ in range(3):
        sess.run(x_dev, graph=use_graph)
    barrier()
    e0.record()
    for _ in range(args.steps):
        sess.run(x_dev, graph=use_graph)
    e1.record()
    barrier()
    ms_inf = e0.elapsed_time(e1) / args.steps

    t = torch.tensor([ms, ms_e2e, ms_inf], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_inf = [float(v) for v in t.tolist()]

    if rank == 0:
        burst, sustained, hbm, src = peaks()
        kms, kflops = time_dominant_kernel(B, size)
        achieved = kflops / (kms * 1e-3) / 1e12
        eng = sess.eng
        launches = (eng.launches_fwd + eng.launches_bwd + 3 + 2 * len(eng.convs)) * args.steps
        out = {
            "metric": METRIC, "value": world * B / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": f"{graph} training step (fwd + smooth-L1 + bwd + Adam), batch {B} clips 16x{size}x{size} per GPU",
                "global_batch": world * B, "parallelism": f"dp{world}", "weights": "random init (TF initialisers)",
                "cuda_graph": use_graph,
                "l2": "per-step working set (activations + gradients, several GB) is far larger than the 126 MB L2",
                "infer_clips_per_s": world * B / (ms_inf * 1e-3), "infer_ms_per_step": ms_inf, "loss": loss_val,
            },
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 4),
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                         "traffic": None, "kernel": "conv_tc_kernel<128,4> (x_1_2/x_1_3 3x3x3 conv, 128+128->128)",
                         "ms_per_launch": kms, "flops_per_launch": kflops, "peak_source": f"{src} bf16 burst (kernel timed alone)"},
        }
        if not args.no_cpu_baseline:
            val, sec, cores = cpu_reference_step(graph, size, 1, 2, 1)
            out["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"2 training steps of {graph} on 1 clip 16x{size}x{size} (fp32 torch-CPU oracle)"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
