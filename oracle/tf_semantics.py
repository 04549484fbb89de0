"""TEST INFRASTRUCTURE — CPU restatement of the TensorFlow-1.x op semantics the reference graph relies on.

Nothing in the product path (sap3d_tensorflow_b200/) may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.

PARITY UNPINNED for these ops: the reference (A-Nasiri-M/sap3d_tensorflow) ships no tests, golden
vectors or seeds, and TensorFlow is not installable in this image.  The semantics
below are the published TF-1.x definitions; each function cites the reference call-site that uses
it.  They are cross-checked against an independent float64 NumPy direct-loop implementation in
oracle/np_direct.py (tests/test_oracle_semantics.py).

All tensors are NDHWC torch tensors (any float dtype; the oracle runs fp32 or fp64 on CPU).
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import torch
import torch.nn.functional as F


def same_pad(i: int, k: int, s: int) -> Tuple[int, int, int]:
    """TF 'SAME' geometry: returns (out, pad_before, pad_after); the extra padding goes at the END."""
    o = -(-i // s)
    total = max((o - 1) * s + k - i, 0)
    return o, total // 2, total - total // 2


def _to_ncdhw(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 4, 1, 2, 3)


def _to_ndhwc(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 2, 3, 4, 1).contiguous()


def conv3d_same(x: torch.Tensor, w: torch.Tensor, strides: Sequence[int] = (1, 1, 1), bias: torch.Tensor | None = None) -> torch.Tensor:
    """tf.nn.conv3d(x, w, strides=[1,sd,sh,sw,1], padding='SAME') (+ tf.nn.bias_add).

    x: [N,D,H,W,Cin]; w: DHWIO [kd,kh,kw,Cin,Cout].  Reference call-sites: p3d.py:19,24,86,112,125,343;
    tf.layers.conv3d(..., 'same') at utils/network.py:101,164-178,188,262.
    """
    kd, kh, kw = w.shape[:3]
    pads = []
    for i, k, s in zip(x.shape[1:4], (kd, kh, kw), strides):
        _, pb, pa = same_pad(i, k, s)
        pads.append((pb, pa))
    xn = _to_ncdhw(x)
    xn = F.pad(xn, (pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]))
    wt = w.permute(4, 3, 0, 1, 2)  # OIDHW
    y = F.conv3d(xn, wt, bias=bias, stride=tuple(strides))
    return _to_ndhwc(y)


def conv3d_transpose_same(x: torch.Tensor, w: torch.Tensor, strides: Sequence[int], bias: torch.Tensor | None = None) -> torch.Tensor:
    """tf.layers.conv3d_transpose(x, Cout, k, strides, 'same') — output size I*s.

    w: [kd,kh,kw,Cout,Cin] (the Keras/tf.layers transposed-conv kernel layout).  Defined as the
    input-gradient of a SAME conv from an I*s tensor: y[p] = sum_i x[i] w[k], p = i*s + k - pb with
    pb = max(k-s,0)//2, cropped to [0, I*s) (positions a k<s kernel never reaches hold only the bias).
    Reference call-sites: utils/network.py:107; p3d.py:200-217,238-275,333,393; gn/p3d_gn.py:20,234-257.
    """
    kd, kh, kw = w.shape[:3]
    xn = _to_ncdhw(x)
    wt = w.permute(4, 3, 0, 1, 2)  # [Cin, Cout, kd, kh, kw] = torch conv_transpose layout
    y = F.conv_transpose3d(xn, wt, bias=None, stride=tuple(strides))
    # full size is (I-1)*s + k; crop/pad to I*s starting at pb
    outs = []
    sl = [slice(None), slice(None)]
    for dim, (i, k, s) in enumerate(zip(x.shape[1:4], (kd, kh, kw), strides)):
        pb = max(k - s, 0) // 2
        want = i * s
        have = (i - 1) * s + k
        if have < pb + want:
            pad = [0, 0, 0, 0, 0, 0]
            pad[2 * (2 - dim) + 1] = pb + want - have
            y = F.pad(y, pad)
        sl.append(slice(pb, pb + want))
    y = y[tuple(sl)]
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1, 1)
    return _to_ndhwc(y)


def max_pool3d_same(x: torch.Tensor, ksize: Sequence[int], strides: Sequence[int]) -> torch.Tensor:
    """tf.nn.max_pool3d(x, [1,kd,kh,kw,1], [1,sd,sh,sw,1], 'SAME'); padding never wins (-inf).

    Reference call-sites: p3d.py:347-348,354,360,366 (k(2,1,1)/s(2,1,1) and k(2,3,3)/s(2,2,2)).
    """
    pads = []
    for i, k, s in zip(x.shape[1:4], ksize, strides):
        _, pb, pa = same_pad(i, k, s)
        pads.append((pb, pa))
    xn = _to_ncdhw(x)
    xn = F.pad(xn, (pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]), value=float("-inf"))
    y = F.max_pool3d(xn, kernel_size=tuple(ksize), stride=tuple(strides))
    return _to_ndhwc(y)


def max_pool3d_valid(x: torch.Tensor, size: int) -> torch.Tensor:
    """tf.layers.max_pooling3d(x, size, size) (padding defaults to 'valid') — utils/network.py:6-7."""
    if size == 1:
        return x
    return _to_ndhwc(F.max_pool3d(_to_ncdhw(x), kernel_size=size, stride=size))


BN_EPS = 1e-3        # tf.layers.batch_normalization default epsilon
BN_MOMENTUM = 0.99   # default momentum


def batch_norm(x, gamma, beta, moving_mean, moving_var, training: bool):
    """tf.layers.batch_normalization(x, training=training): axis -1, eps 1e-3, momentum 0.99.

    training: biased batch variance over (N,D,H,W); returns (y, new_moving_mean, new_moving_var)
    (5-D input => non-fused TF path => moving variance updated with the biased variance).
    Reference call-sites: p3d.py:58-127,344; utils/network.py:91.
    """
    if training:
        mean = x.mean(dim=(0, 1, 2, 3))
        var = x.var(dim=(0, 1, 2, 3), unbiased=False)
        y = (x - mean) * torch.rsqrt(var + BN_EPS) * gamma + beta
        new_mm = moving_mean * BN_MOMENTUM + mean.detach() * (1 - BN_MOMENTUM)
        new_mv = moving_var * BN_MOMENTUM + var.detach() * (1 - BN_MOMENTUM)
        return y, new_mm, new_mv
    y = (x - moving_mean) * torch.rsqrt(moving_var + BN_EPS) * gamma + beta
    return y, moving_mean, moving_var


def group_norm(x, gamma, beta, G: int = 32, eps: float = 1e-5):
    """utils/network.py:65-87 == gn/p3d_gn.py:24-46: G = min(G, C), moments over (C/G, D, H, W) per
    sample (biased variance), per-channel gamma/beta."""
    n, d, h, w, c = x.shape
    g = min(G, c)
    xg = x.reshape(n, d * h * w, g, c // g)
    mean = xg.mean(dim=(1, 3), keepdim=True)
    var = xg.var(dim=(1, 3), keepdim=True, unbiased=False)
    y = ((xg - mean) / torch.sqrt(var + eps)).reshape(n, d, h, w, c)
    return y * gamma + beta


def smooth_l1_loss(pred, target, sigma: float = 1.0):
    """utils/network.py:49-62 with inside/outside weights 1 (train.py:159): sum over ALL elements
    (tf.reduce_mean of the reduce_sum scalar is the identity)."""
    s2 = sigma ** 2
    d = pred - target
    a = d.abs()
    sign = (a < 1.0 / s2).to(d.dtype)
    loss = d * d * (s2 / 2.0) * sign + (a - 0.5 / s2) * (1.0 - sign)
    return loss.sum()


def adam_step_tf(p, g, m, v, t: int, lr: float = 1e-4, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """tf.train.AdamOptimizer (train.py:168): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    p = p - lr_t * m / (v.sqrt() + eps)
    return p, m, v
