"""Data-parallel training over the GPUs of one box: one process per GPU (torch.distributed / NCCL
over NVLink 5 + NVSwitch for the plumbing), clips sharded by rank, weights replicated.

The reference is single-GPU (train.py:73); DP is new work required by BASELINE.json configs[3].
The only exchange step is the gradient sum: the loss is a plain SUM over all elements
(utils/network.py:60), so the global-batch gradient is the sum of the shard gradients — no averaging.
Gradients travel as bf16 buckets (half the NVLink bytes); BatchNorm statistics stay per replica by default.
`enable_sync_batch_norm` makes them span the replicas (the reference's single-device batch): an option for
parity runs -- two small latency-bound collectives per norm layer, eager mode only; throughput is reported without it.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _abi as A


def make_buckets(n: int, per: int):
    """[lo, hi) element ranges covering [0, n), built from the END of the flat gradient buffer: parameters are laid
    out in forward order, so backward completes them tail-first and the tail buckets can be reduced first."""
    buckets = []
    hi = n
    while hi > 0:
        lo = max(0, hi - per)
        buckets.append((lo, hi))
        hi = lo
    return buckets


def shard_clips(n_clips: int, rank: int, world: int):
    """contiguous clip range of `rank` (inference / evaluation sharding: no communication during forward)"""
    per = (n_clips + world - 1) // world
    lo = min(n_clips, rank * per)
    return lo, min(n_clips, lo + per)


def reduce_metric_sums(sums: torch.Tensor, counts: torch.Tensor):
    """end-of-evaluation exchange: all-reduce per-metric (sum over non-NaN clips, count) pairs, return the means
    (NaN filtering mirrors test.py:177-181)"""
    if dist.is_initialized() and dist.get_world_size() > 1:
        packed = torch.stack([sums, counts]).contiguous()
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
        sums, counts = packed[0], packed[1]
    return sums / counts


def segment_buckets(segments, per: int):
    """per gradient segment [lo, hi) (in the order backward completes them): its buckets, tail first"""
    return [[(a + lo, b + lo) for a, b in make_buckets(hi - lo, per)] for lo, hi in segments]


def split_buckets(n: int, per: int, split: int):
    """buckets of the tail [split, n) (exchanged while backward is still running) and of the head [0, split)"""
    tail, head = segment_buckets([(split, n), (0, split)], per)
    return tail, head


class GradientExchange:
    """sum-all-reduce of the flat gradient buffer as bf16 on a communication stream.  The reduced values stay in `buf` and
    Adam reads them there (Engine.grad_source): no pass widens them back into the fp32 buffer.
    Default (measured best at 2 and 8 GPUs, profiles/r02_c14_dp_timeline_8gpu.txt): ONE all-reduce call over the whole
    buffer after backward -- 0.59 ms for 170 MB on 8 GPUs; 32 MB buckets cost 1.01 ms for the same bytes.
    Overlapped form (overlap=True / SAP3D_DP_OVERLAP=1; the builders' mark_dp_split cuts backward into segments, one CUDA graph
    each): start(eng, k) after segment k of the backward has finished -- its share of the gradient buffer is cast and its
    all-reduce is queued while the next segment runs --, finish(eng) after the last one.  On this hardware the backward segment
    that runs beside the all-reduce slows down by the all-reduce's own duration, so the overlap buys nothing (18.70 vs 18.49 ms
    per step at 8 GPUs) and it is not the default."""

    def __init__(self, eng, bucket_mb: int = 0, overlap: bool = False):
        self.eng = eng
        n = eng.n_train
        self.n = n
        self.buf = torch.zeros(n, device=eng.device, dtype=torch.bfloat16)
        per = bucket_mb * 1024 * 1024 // 2 if bucket_mb > 0 else n      # 0 = one all-reduce call per segment / buffer
        self.buckets = make_buckets(n, per)
        self.overlap = bool(overlap) and len(eng.dp_segments) > 1
        self.segments = [(g_lo, g_hi) for (_, _, g_lo, g_hi) in eng.dp_segments] if self.overlap else []
        self.seg_buckets = segment_buckets(self.segments, per)
        self.comm_stream = torch.cuda.Stream(device=eng.device)
        self._works = []
        eng.grad_source = self.buf

    def _reduce(self, eng, k):
        lo, hi = self.segments[k]
        cur = torch.cuda.current_stream(eng.device)
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            # the narrowing pass runs on the communication stream too: the main stream goes straight on to the next backward
            # segment, which writes a disjoint part of flat_g
            A.check(A.lib.sap3d_cast(A.F32, A.ptr(eng.flat_g[lo:hi]), A.ptr(self.buf[lo:hi]), hi - lo, self.comm_stream.cuda_stream), "grad cast")
            self._works += [dist.all_reduce(self.buf[a:b], op=dist.ReduceOp.SUM, async_op=True) for a, b in self.seg_buckets[k]]

    def start(self, eng, k: int = 0):
        if k == 0:
            self._works = []
        self._reduce(eng, k)

    def finish(self, eng):
        cur = torch.cuda.current_stream(eng.device)
        self._reduce(eng, len(self.segments) - 1)
        with torch.cuda.stream(self.comm_stream):
            for w in self._works:
                w.wait()
        cur.wait_stream(self.comm_stream)

    def __call__(self, eng):
        cur = torch.cuda.current_stream(eng.device)
        st = cur.cuda_stream
        A.check(A.lib.sap3d_cast(A.F32, A.ptr(eng.flat_g), A.ptr(self.buf), self.n, st), "grad cast")
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            works = [dist.all_reduce(self.buf[lo:hi], op=dist.ReduceOp.SUM, async_op=True) for lo, hi in self.buckets]
            for w in works:
                w.wait()
        cur.wait_stream(self.comm_stream)

    def gradient(self) -> torch.Tensor:
        """the exchanged gradient (fp32 copy of the bf16 buckets), for inspection"""
        return self.buf.float()


class SyncBatchNorm:
    """sum-all-reduce of the small per-layer statistics buffers (forward: the conv epilogue's [rows][2][C] partial sums;
    backward: the 4*C per-channel sums of the BN gradient) on the current stream, in fp32"""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise A.Sap3dError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.calls = 0

    def all_reduce(self, t: torch.Tensor):
        self.calls += 1
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)


def enable_sync_batch_norm(sess, group=None) -> SyncBatchNorm:
    """every batch-statistics BatchNorm of the session's graph normalises over the GLOBAL batch (all replicas must run the same
    graph with the same per-replica batch).  Eager execution only: Session.train_step(graph=True) refuses it."""
    sb = SyncBatchNorm(group)
    sess.eng.sync_bn = sb
    sess.graph_train = None
    sess.graph_fwd = None
    return sb


class ExactGradientExchange:
    """fp32 sum-all-reduce of the whole flat gradient buffer in one call: the parity-test form of GradientExchange"""

    def __init__(self, group=None):
        self.group = group

    def __call__(self, eng):
        dist.all_reduce(eng.flat_g, op=dist.ReduceOp.SUM, group=self.group)


def attach_data_parallel(sess, bucket_mb: Optional[int] = None, sync_bn: bool = False, exact: bool = False, overlap: Optional[bool] = None):
    """installs the gradient all-reduce between backward and Adam; broadcasts rank 0's variables.
    sync_bn: BatchNorm statistics over the global batch (parity option).  exact: fp32 un-bucketed exchange.
    bucket_mb: 0 = one all-reduce call (default; SAP3D_DP_BUCKET_MB overrides); overlap: run the exchange of finished backward
    segments beside the rest of backward (default off; SAP3D_DP_OVERLAP=1)."""
    if not dist.is_initialized():
        raise A.Sap3dError("torch.distributed is not initialised")
    eng = sess.eng
    if bucket_mb is None:
        bucket_mb = int(os.environ.get("SAP3D_DP_BUCKET_MB", "0"))
    if overlap is None:
        overlap = os.environ.get("SAP3D_DP_OVERLAP", "0") == "1"
    dist.broadcast(eng.flat_w, src=0)
    eng.pack_weights()
    # every replica draws its own dropout masks: fold the rank into the seed before the step is captured (the mask hash
    # depends only on (seed, op index, step, element index), so equal seeds would correlate the noise across the global batch)
    eng.dropout_seed = eng.dropout_seed * 8191 + dist.get_rank() + 1
    sess.graph_train = None
    sess.graph_fwd = None
    eng.grad_source = None
    ex = ExactGradientExchange() if exact else GradientExchange(eng, bucket_mb, overlap)
    sess.grad_hook = ex
    if sync_bn:
        enable_sync_batch_norm(sess)
    return ex
