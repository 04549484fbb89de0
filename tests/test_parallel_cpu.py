"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: bucket partition, clip sharding, the
sum-semantics of the gradient exchange (the loss is a SUM over elements, so shard gradients add up to the
global-batch gradient with no averaging) and the NaN-filtered metric reduction of the sharded evaluation."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_buckets_cover_exactly_once():
    from sap3d_tensorflow_b200.parallel import make_buckets

    for n, per in [(10, 3), (85_000_000, 16 * 1024 * 1024), (5, 100), (64, 64)]:
        b = make_buckets(n, per)
        assert b[0][1] == n and b[-1][0] == 0
        assert all(b[i][0] == b[i + 1][1] for i in range(len(b) - 1))
        assert all(0 < hi - lo <= per for lo, hi in b)


def test_split_buckets_cover_tail_then_head():
    """overlapped exchange: the tail [split, n) (stage 3 + decoder gradients, finished first) and the head [0, split)
    are bucketed separately and together cover the flat gradient buffer exactly once"""
    from sap3d_tensorflow_b200.parallel import split_buckets

    for n, per, split in [(100, 16, 37), (85_000_000, 16 * 1024 * 1024, 10_700_000), (50, 8, 0), (50, 8, 50)]:
        tail, head = split_buckets(n, per, split)
        cover = sorted(tail + head)
        assert cover[0][0] == 0 and cover[-1][1] == n if n else True
        assert all(cover[i][1] == cover[i + 1][0] for i in range(len(cover) - 1))
        assert all(lo >= split for lo, _ in tail) and all(hi <= split for _, hi in head)


def test_segment_buckets_follow_the_backward_segments():
    """multi-segment exchange (Engine.dp_segments: the gradient ranges in the order backward completes them, tail first): every
    segment is bucketed on its own, tail bucket first, and together they cover the buffer exactly once"""
    from sap3d_tensorflow_b200.parallel import segment_buckets

    n = 84_922_240
    segments = [(3_000_128, n), (262_976, 3_000_128), (0, 262_976)]       # the _ds graph's three segments (r02 trace)
    for per in (16 * 1024 * 1024, n):                                      # 32 MB buckets; one call per segment
        sb = segment_buckets(segments, per)
        assert len(sb) == len(segments)
        for (lo, hi), buckets in zip(segments, sb):
            assert buckets[0][1] == hi and buckets[-1][0] == lo            # tail first, nothing outside the segment
            assert all(buckets[i][0] == buckets[i + 1][1] for i in range(len(buckets) - 1))
            assert all(0 < b - a <= per for a, b in buckets)
        cover = sorted(b for bs in sb for b in bs)
        assert cover[0][0] == 0 and cover[-1][1] == n and all(cover[i][1] == cover[i + 1][0] for i in range(len(cover) - 1))
    assert [len(b) for b in segment_buckets(segments, n)] == [1, 1, 1]


def test_bench_defaults_per_workload(monkeypatch):
    """bench.py: 8 clips per GPU for the training workloads (configs[1]), 32 clips per iteration and per-clip BatchNorm statistics
    for the evaluation workload (configs[4], gen_pred.py semantics); explicit flags win"""
    import importlib.util
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for argv, want in [([], ("train", 8, "clip")), (["--workload", "eval"], ("eval", 32, "clip")),
                       (["--workload", "eval", "--batch", "8", "--bn-statistics", "batch"], ("eval", 8, "batch")),
                       (["--workload", "gn160"], ("gn160", 8, "clip")), (["--batch", "32"], ("train", 32, "clip"))]:
        monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
        a = bench.parse()
        assert (a.workload, a.batch, a.bn_statistics) == want, (argv, a)


def test_clip_sharding_partitions_the_evaluation_set():
    from sap3d_tensorflow_b200.parallel import shard_clips

    for n, world in [(1024, 8), (10, 4), (3, 8)]:
        got = [shard_clips(n, r, world) for r in range(world)]
        covered = [i for lo, hi in got for i in range(lo, hi)]
        assert covered == list(range(n))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import tf_semantics as tfs
    from sap3d_tensorflow_b200.parallel import make_buckets, reduce_metric_sums

    # (1) gradient exchange semantics on a tiny model: sum-loss => grad(global batch) == sum of shard grads
    torch.manual_seed(0)
    w = torch.randn(6, 4, dtype=torch.float64, requires_grad=True)
    x = torch.randn(4, 6, dtype=torch.float64)      # global batch of 4 "clips"
    y = torch.rand(4, 4, dtype=torch.float64)
    tfs.smooth_l1_loss(torch.sigmoid(x @ w), y).backward()
    g_global = w.grad.clone()
    w.grad = None
    lo, hi = rank * 2, rank * 2 + 2
    tfs.smooth_l1_loss(torch.sigmoid(x[lo:hi] @ w), y[lo:hi]).backward()
    flat = w.grad.reshape(-1).clone().float()
    for a, b in make_buckets(flat.numel(), 7):      # bucketed all-reduce, tail first
        dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM)
    ok1 = torch.allclose(flat.double().reshape(6, 4), g_global, atol=1e-6)
    # (2) sharded evaluation: NaN-filtered means
    vals = torch.tensor([[0.5, float("nan")], [0.7, 0.2]]) if rank == 0 else torch.tensor([[0.9, 0.4], [float("nan"), 0.6]])
    okm = ~torch.isnan(vals)
    means = reduce_metric_sums(torch.where(okm, vals, torch.zeros_like(vals)).sum(0), okm.sum(0).float())
    ok2 = torch.allclose(means, torch.tensor([0.7, 0.4]), atol=1e-6)
    # (3) synchronised BatchNorm algebra (engine._NormActOp with parallel.SyncBatchNorm): statistics rows and the backward's
    # per-channel sums are SUMMED over the replicas and divided by the global count; d(gamma)/d(beta) stay local sums that
    # the gradient exchange adds up.  Checked against autograd over the whole batch.
    from sap3d_tensorflow_b200.parallel import SyncBatchNorm
    sb = SyncBatchNorm()
    torch.manual_seed(1)
    xb = torch.randn(4, 5, 3, dtype=torch.float64, requires_grad=True)     # global batch 4, 5 positions, 3 channels
    gamma = torch.rand(3, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(4, 5, 3, dtype=torch.float64)
    mean, var = xb.mean((0, 1)), xb.var((0, 1), unbiased=False)
    ((xb - mean) / torch.sqrt(var + 1e-3) * gamma * dy).sum().backward()
    xs, dys = xb.detach()[lo:hi], dy[lo:hi]
    stats = torch.stack([xs.sum((0, 1)), (xs * xs).sum((0, 1))])           # this replica's [2][C] row
    sb.all_reduce(stats)
    cnt = xs.shape[0] * xs.shape[1] * sb.world
    m = stats[0] / cnt
    v = stats[1] / cnt - m * m
    rstd = 1.0 / torch.sqrt(v + 1e-3)
    xhat = (xs - m) * rstd
    g = dys * gamma.detach()
    local = torch.stack([g.sum((0, 1)), (g * xhat).sum((0, 1))])
    dgamma_local = (dys * xhat).sum((0, 1))
    coef = local / cnt                                                     # phase 1 of sap3d_affine_act_bwd_sync
    sb.all_reduce(coef)
    dx = rstd * (g - coef[0] - xhat * coef[1])                             # phase 2
    dist.all_reduce(dgamma_local)                                          # what the gradient exchange does
    ok3 = (torch.allclose(m, mean.detach()) and torch.allclose(v, var.detach()) and torch.allclose(dx, xb.grad[lo:hi], atol=1e-10)
           and torch.allclose(dgamma_local, gamma.grad, atol=1e-10) and sb.calls == 2)
    q.put((rank, bool(ok1), bool(ok2 and ok3)))
    dist.destroy_process_group()


def test_two_rank_gloo_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(a and b for _, a, b in res), res
