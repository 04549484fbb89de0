"""Data-parallel training over the GPUs of one box: one process per GPU (torch.distributed / NCCL
over NVLink 5 + NVSwitch for the plumbing), clips sharded by rank, weights replicated.

The reference is single-GPU (train.py:73); DP is new work required by BASELINE.json configs[3].
The only exchange step is the gradient sum: the loss is a plain SUM over all elements
(utils/network.py:60), so the global-batch gradient is the sum of the shard gradients — no averaging.
Gradients travel as bf16 buckets (half the NVLink bytes); BatchNorm statistics stay per replica.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _abi as A


class GradientExchange:
    def __init__(self, eng, bucket_mb: int = 32):
        self.eng = eng
        n = eng.n_train
        self.n = n
        self.buf = torch.zeros(n, device=eng.device, dtype=torch.bfloat16)
        per = bucket_mb * 1024 * 1024 // 2
        # buckets from the END of the flat gradient buffer: parameters are laid out in forward order, so
        # backward completes them tail-first
        self.buckets = []
        hi = n
        while hi > 0:
            lo = max(0, hi - per)
            self.buckets.append((lo, hi))
            hi = lo
        self.comm_stream = torch.cuda.Stream(device=eng.device)

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
        A.check(A.lib.sap3d_cast(A.BF16, A.ptr(self.buf), A.ptr(eng.flat_g), self.n, st), "grad uncast")


def attach_data_parallel(sess, bucket_mb: int = 32) -> GradientExchange:
    """installs the gradient all-reduce between backward and Adam; broadcasts rank 0's variables"""
    if not dist.is_initialized():
        raise A.Sap3dError("torch.distributed is not initialised")
    eng = sess.eng
    dist.broadcast(eng.flat_w, src=0)
    eng.pack_weights()
    ex = GradientExchange(eng, bucket_mb)
    sess.grad_hook = ex
    return ex
