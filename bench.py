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
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU and step (default: 8 for the training workloads, 32 for eval)")
    ap.add_argument("--size", type=int, default=112)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="train", choices=["train", "gn160", "eval"],
                    help="train: BASELINE configs[1] (default, the headline line); gn160: configs[2] (GN+CBAM, batch 16, 160x160, "
                         "fwd+bwd+Adam); eval: configs[4] (1024 clips sharded over the ranks, forward + CC/SIM/NSS/KLdiv)")
    ap.add_argument("--clips", type=int, default=1024, help="evaluation set size of --workload eval")
    ap.add_argument("--full-res", action="store_true",
                    help="--workload eval: test.py convention (upsample the last frame to 1080x960, CC/SIM/AUC_Judd/AUC_Borji/NSS)")
    ap.add_argument("--eager", action="store_true", help="no CUDA-graph replay (debug)")
    ap.add_argument("--bn-statistics", default="clip", choices=["clip", "batch"],
                    help="--workload eval: the backbone's batch-statistics BatchNorm normalises every clip on its own (what gen_pred.py's "
                         "one-window-per-sess.run gives; default) or the whole batch of clips together")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 32 if args.workload == "eval" else 8
    return args


DEFAULT_GRAPH = "p3d_unetplusplus_ds"     # p3d.py:340 (what gen_pred.py:46 builds); named literally: the reference arm must not
                                          # import the product package


def default_graph():
    return DEFAULT_GRAPH


def workload_config(graph: str, batch: int, size: int, world: int):
    """the `config` object both arms print (identical for the same command line)"""
    return {
        "workload": f"{graph} training step (fwd + smooth-L1 + bwd + Adam), batch {batch} clips 16x{size}x{size} per GPU",
        "global_batch": world * batch, "parallelism": f"dp{world}", "weights": "random init (TF initialisers)",
        "l2": "per-step working set (activations + gradients, several GB) is far larger than the 126 MB L2",
    }


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


REF_SAMPLE_BATCH = 2   # clips per reference-arm step: a bounded sample of the batch-8 step (a full batch is ~17 s per step on
                       # 16 host cores; K + W = 25 steps of it would take 7 minutes)


def run_reference(args):
    """reference arm: the CPU restatement of the reference's TF graph (TensorFlow cannot be installed in this image, so
    `oracle/` stands in for the reference's own CPU path; kind = "port"), all host threads, EXACTLY --steps timed steps after
    --warmup untimed ones.  Each step is a training step on REF_SAMPLE_BATCH clips of the configured workload; the value is
    clips / second.  Imports nothing from the product package."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gn = args.workload == "gn160"
    graph = args.graph or ("inference_p3d" if gn else default_graph())
    B, size = (16, 160) if gn and args.batch == 8 and args.size == 112 else (args.batch, args.size)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    sb = min(REF_SAMPLE_BATCH, B)
    val, sec, cores = cpu_reference_step(graph, size, sb, steps, warm)
    sample = (f"each step = one training step of {graph} on {sb} clip(s) 16x{size}x{size} (a {sb}/{B} sample of the batch-{B} step; "
              f"fp32, torch-CPU restatement of the reference graph, {cores} threads); {steps} timed steps after {warm} warm-up steps")
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": sec * 1e3 * B / sb, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(graph, B, size, max(1, args.gpus)),
        "note": "reference = CPU restatement of the reference TF graph (TensorFlow unavailable in this image); ms_per_step is "
                "scaled from the sample to the full batch",
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if gn:
        out["metric"] = "clips/sec (16x160x160, bf16) GN+CBAM train step"
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
DOMINANT = {
    # workload -> the FLOP-dominant launch of its step: (description, D, H/size, W/size divisor, input segments, cout)
    "train": ("x_1_2 / x_1_3 decoder conv (3x3x3, 128+128 -> 128 at B x 8 x size/2 x size/2; p3d.py:388-390)", 8, 2, [128, 128], 128, 2),
    "eval": ("x_1_2 / x_1_3 decoder conv (3x3x3, 128+128 -> 128 at B x 8 x size/2 x size/2; p3d.py:388-390)", 8, 2, [128, 128], 128, 2),
    "gn160": ("conv_concat (3x3x3, 1536+256 -> 1024 at B x 4 x size/4 x size/4; gn/p3d_gn.py:245)", 4, 4, [1536, 256], 1024, 1),
}


def time_dominant_kernel(workload: str, batch: int, size: int, iters: int = 20):
    """the FLOP-dominant launch of the workload's step, timed alone with CUDA events on the launching stream, L2 flushed
    between launches.  Returns (ms per launch, FLOPs per launch, kernel name as CUPTI reports it, launches of it per step)."""
    import ctypes as C

    import torch
    from torch.profiler import ProfilerActivity, profile
    from sap3d_tensorflow_b200 import _abi as A

    desc, D, div, cins, cout, per_step = DOMINANT[workload]
    dev = torch.device("cuda")
    N, H, W = batch, size // div, size // div
    xs = [torch.randn(N, D, H, W, c, device=dev).to(torch.bfloat16) for c in cins]
    w = torch.randn(3, 3, 3, sum(cins), cout, device=dev) * 0.02
    b = torch.zeros(cout, device=dev)
    d = A.make_conv_desc(A.BF16, N, D, H, W, cins, cout, (3, 3, 3), (1, 1, 1), False, True, False, A.IMPL_TC)
    y = torch.empty(N, D, H, W, cout, device=dev, dtype=torch.bfloat16)
    rows = A.lib.sap3d_conv_stats_rows(C.byref(d))
    stats = torch.zeros(rows, 2, cout, device=dev)
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
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        launch()
        torch.cuda.synchronize()
    names = [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "conv_tc" in e.name]
    flops = 2.0 * N * D * H * W * 27 * sum(cins) * cout
    return ms, flops, (names[0] if names else None), per_step, desc


def in_graph_kernel_ms(step_fn, kernel_name: str, topk: int):
    """device durations of the `topk` longest launches of `kernel_name` INSIDE one replay of the step's CUDA graphs (CUPTI
    activity records): the same launches the stand-alone timing isolates, warm and with their real neighbours."""
    import torch
    from torch.profiler import ProfilerActivity, profile

    if not kernel_name:
        return None
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step_fn()
        torch.cuda.synchronize()
    ds = sorted((e.time_range.end - e.time_range.start for e in prof.events()
                 if e.device_type == torch.autograd.DeviceType.CUDA and e.name == kernel_name), reverse=True)
    if len(ds) < topk:
        return None
    return sum(ds[:topk]) / topk / 1e3


def count_launches(fn):
    """kernels launched on the device by one call of fn() (CUPTI activity records; memcpy / memset excluded)"""
    import torch
    from torch.profiler import ProfilerActivity, profile

    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    n = 0
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and not e.name.startswith(("Memcpy", "Memset", "memcpy", "memset")):
            n += 1
    return n


def run_eval(args):
    """BASELINE configs[4]: gen_pred.py-style batched inference + CC / SIM / NSS / KLdiv over --clips synthetic clips,
    contiguous clip ranges per rank, no communication during the forward passes, one final all-reduce of the per-metric
    (sum, count) pairs (test.py:164-183 scores the last frame of every clip)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import sap3d_tensorflow_b200 as sp
    from sap3d_tensorflow_b200 import metrics, parallel

    graph = args.graph or default_graph()
    B, size = args.batch, args.size
    dev = torch.device("cuda", local_rank)
    xin = sp.placeholder([B, 16, size, size, 3], dtype="bf16", training_graph=False, device=f"cuda:{local_rank}",
                         per_sample_statistics=(args.bn_statistics == "clip"))
    head = getattr(sp.p3d, graph)(xin, 0.0, B, False)
    sess = sp.Session(head)
    lo, hi = parallel.shard_clips(args.clips, rank, world)
    nb = (hi - lo + B - 1) // B
    g = torch.Generator(device="cpu").manual_seed(99 + rank)
    # a small pool of distinct pinned host batches, cycled (1024 distinct clips would be 3.2 GB of host memory per rank)
    pool = 4
    xs = [((torch.randint(0, 256, (B, 16, size, size, 3), generator=g).float() - torch.tensor([90.0, 102.0, 98.0])) / 255.0).pin_memory()
          for _ in range(pool)]
    gh, gw = (1080, 960) if args.full_res else (size, size)
    dens = [torch.rand(B, gh, gw, generator=g).pin_memory() for _ in range(pool)]
    fixs = [(torch.rand(B, gh, gw, generator=g) < (4e-5 if args.full_res else 0.01)).float().pin_memory() for _ in range(pool)]
    nmet = 5 if args.full_res else 4
    score = metrics.evaluate_clips_test_time if args.full_res else metrics.evaluate_clips
    for f in fixs:
        f[:, 0, 0] = 1.0   # at least one fixation per map

    def one_pass(host: bool):
        sums = torch.zeros(nmet, device=dev, dtype=torch.float64)
        cnts = torch.zeros(nmet, device=dev, dtype=torch.float64)
        if host:
            sess.prefetch(xs[0])     # as in the training arm: the NEXT batch's H2D copy runs on the copy stream under the current batch
        for i in range(nb):
            j = i % pool
            d = dens[j].to(dev, non_blocking=True) if host else dens_dev[j]
            f = fixs[j].to(dev, non_blocking=True) if host else fixs_dev[j]
            if host:
                pred = sess.run(None, graph=True)            # consumes the staged batch
                if i + 1 < nb:
                    sess.prefetch(xs[(i + 1) % pool])
            else:
                pred = sess.run(xs_dev[j], graph=True)
            r = score(pred, d, f)
            sums += r["sum"]
            cnts += r["count"]
        return parallel.reduce_metric_sums(sums, cnts)

    xs_dev = [x.to(dev) for x in xs]
    dens_dev = [d.to(dev) for d in dens]
    fixs_dev = [f.to(dev) for f in fixs]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    one_pass(False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, args.steps // 5)
    barrier()
    e0.record()
    for _ in range(reps):
        means = one_pass(False)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    e0.record()
    for _ in range(reps):
        means = one_pass(True)
        means_host = means.cpu()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / reps
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = [float(v) for v in t.tolist()]
    if rank == 0:
        launches = count_launches(lambda: (sess.run(xs_dev[0], graph=True), score(sess.head.output, dens_dev[0], fixs_dev[0])))
        out = {
            "metric": "clips/sec (16x112x112, bf16) inference + " + ("CC/SIM/AUC_Judd/AUC_Borji/NSS at 1080x960" if args.full_res else "CC/SIM/NSS/KLdiv"), "value": args.clips / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": reps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{graph} inference (training=False) + saliency metrics over {args.clips} clips 16x{size}x{size}, "
                                   f"batch {B} per iteration ({'per-clip' if args.bn_statistics == 'clip' else 'per-batch'} BatchNorm statistics), "
                                   f"clips sharded contiguously over {world} rank(s)",
                       "parallelism": f"dp{world}", "cuda_graph": True, "metric_means": [float(v) for v in means_host.tolist()],
                       "metric_names": ["CC", "SIM", "AUC_Judd", "AUC_Borji", "NSS"] if args.full_res else ["CC", "SIM", "NSS", "KLdiv"],
                       "l2": "every batch's activations (> 1 GB) exceed the 126 MB L2"},
            "e2e": {"value": args.clips / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(nb * (xs[0].numel() + 2 * dens[0].numel()) * 4),
                    "d2h_bytes_per_step": 32, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches * nb * reps), "clocks": clocks,
        }
        burst, sustained, hbm, src = peaks()
        kms, kflops, kname, per_step, kdesc = time_dominant_kernel("eval", B, size)
        achieved = kflops / (kms * 1e-3) / 1e12
        out["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst, "traffic": None,
                           "kernel": f"{kname}: {kdesc}, B={B}", "ms_per_launch": kms, "flops_per_launch": kflops,
                           "peak_source": f"{src} bf16 burst (kernel timed alone, L2 flushed)"}
        ig = in_graph_kernel_ms(lambda: sess.run(xs_dev[0], graph=True), kname, per_step)
        if ig is not None:
            out["roofline"]["in_graph"] = {"ms_per_launch": ig, "achieved": kflops / (ig * 1e-3) / 1e12, "peak": sustained,
                                           "frac": kflops / (ig * 1e-3) / 1e12 / sustained}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


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

    gn = args.workload == "gn160"
    graph = args.graph or ("inference_p3d" if gn else default_graph())
    B, size = (16, 160) if gn and args.batch == 8 and args.size == 112 else (args.batch, args.size)
    dev = torch.device("cuda", local_rank)
    xin = sp.placeholder([B, 16, size, size, 3], dtype="bf16", training_graph=True, device=f"cuda:{local_rank}")
    if gn:
        from sap3d_tensorflow_b200.gn import p3d_gn

        head = getattr(p3d_gn, graph)(xin, 0.5, B, True)
    else:
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
    # data-parallel correctness: after the timed steps every replica must hold bit-identical weights (same initial broadcast,
    # same all-reduced gradients, same Adam).  Checksum = exact integer sum of the fp32 bit patterns; min == max over the ranks.
    replicas_identical = None
    if world > 1:
        bits = sess.eng.flat_w[:sess.eng.n_train].view(torch.int32).to(torch.int64)
        cs = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=dev) % 8191 + 1)).sum()])
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        replicas_identical = bool(torch.equal(lo, hi))

    # ---- end-to-end timing: pinned host inputs -> H2D -> step -> D2H loss -------------------------
    # Every step's inputs come from pinned host memory and its loss is read back to the host.  As in the reference's
    # tensorpack pipeline (PrefetchDataZMQ, train.py:120-135) the NEXT batch is staged while the current step runs:
    # Session.prefetch() issues its H2D copy on a copy stream, train_step(None, None) consumes it.
    sess.prefetch(x_host, y_host)
    for _ in range(2):
        sess.train_step(None, None, graph=use_graph)
        sess.prefetch(x_host, y_host)
        float(sess.eng.loss_buf.item())
    barrier()
    e0.record()
    for _ in range(args.steps):
        sess.train_step(None, None, graph=use_graph)
        sess.prefetch(x_host, y_host)          # H2D of the next step's 25.7 MB overlaps this step's kernels
        float(sess.eng.loss_buf.item())        # D2H read of this step's loss
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
    # kernels per step, counted by CUPTI on EVERY rank (the step contains the gradient all-reduce: all ranks must take it)
    launches_per_step = count_launches(lambda: sess.train_step(x_dev, y_dev, graph=use_graph))
    barrier()

    if rank == 0:
        burst, sustained, hbm, src = peaks()
        wl = "gn160" if gn else "train"
        kms, kflops, kname, per_step, kdesc = time_dominant_kernel(wl, B, size)
        achieved = kflops / (kms * 1e-3) / 1e12
        launches = launches_per_step * args.steps
        traffic = None
        tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")   # dram bytes per launch from the ncu --set full capture
        if os.path.exists(tp) and not gn and B == 8 and size == 112:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        config = workload_config(graph, B, size, world)
        config["cuda_graph"] = use_graph
        out = {
            "metric": METRIC, "value": world * B / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": config,
            "extra": {"infer_clips_per_s": world * B / (ms_inf * 1e-3), "infer_ms_per_step": ms_inf, "loss": loss_val,
                      "replicas_identical": replicas_identical},
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 4),
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                         "traffic": traffic, "kernel": f"{kname}: {kdesc}, B={B}",
                         "ms_per_launch": kms, "flops_per_launch": kflops, "peak_source": f"{src} bf16 burst (kernel timed alone, L2 flushed)"},
        }
        ig = in_graph_kernel_ms(lambda: sess.train_step(x_dev, y_dev, graph=use_graph), kname, per_step) if (use_graph and world == 1) else None
        if ig is not None:
            out["roofline"]["in_graph"] = {"ms_per_launch": ig, "achieved": kflops / (ig * 1e-3) / 1e12, "peak": sustained,
                                           "frac": kflops / (ig * 1e-3) / 1e12 / sustained,
                                           "what": f"mean of the {per_step} longest launches of this kernel inside one CUDA-graph replay of the "
                                                   f"step (CUPTI), against the {src} sustained bf16 peak"}
        if gn:
            out["metric"] = "clips/sec (16x160x160, bf16) GN+CBAM train step"
        if not args.no_cpu_baseline and not gn:
            val, sec, cores = cpu_reference_step(graph, size, 1, 2, 1)
            out["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"2 training steps of {graph} on 1 clip 16x{size}x{size} after 1 warm-up step (fp32 torch-CPU "
                                             f"restatement of the reference graph, {cores} threads; a 1/{B} sample of the batch-{B} step)"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "eval":
        run_eval(a)
    else:
        run_ours(a)
