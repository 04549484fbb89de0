"""TEST INFRASTRUCTURE — CPU oracle (torch, fp32/fp64) of the reference's P3D saliency graphs.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this; the product path never does.

PARITY PINNING.  The reference has no tests / golden vectors and TensorFlow cannot be installed in this image.
  * PINNED to the reference's own code: the graph WIRING of every builder below (layer order, kernels, strides, scopes,
    variable names, shapes and creation order, Python-2 integer division, `training` never reaching make_block) and the
    composite ops the reference spells out in primitive tf ops (GroupNorm, CBAM, attention, smooth_l1_loss) --
    tests/test_reference_wiring_cpu.py executes /root/reference/p3d.py, utils/network.py and gn/p3d_gn.py unmodified over a
    TF-1.x API emulation (tests/golden/tf1_emulation.py) and compares outputs, variable sets and creation order with this
    module, live in the build container and against tests/golden/reference_graphs_golden.npz everywhere.
  * UNPINNED: what TensorFlow's own primitive kernels compute (tf.nn.conv3d, tf.layers.*, tf.nn.max_pool3d ... -- an
    un-vendored dependency); that is oracle/tf_semantics.py, the published TF-1.x definitions, cross-checked only against
    independent fp64 loops (oracle/np_direct.py).

Graphs restated (reference file:line):
  p3d.p3d_unetplusplus_ds      p3d.py:340-399   (primary; what gen_pred.py:46 builds)
  p3d.p3d_unetplusplus_nonsa   p3d.py:401-459
  p3d.p3d_unet                 p3d.py:169-221
  p3d.p3d_concat               p3d.py:224-276
  gn.inference_p3d             gn/p3d_gn.py:214-258 (GroupNorm + CBAM backbone)
  gn.inference_p3d_concat      gn/p3d_gn.py:279-324
  gn.inference_p3d_decoder_block gn/p3d_gn.py:489-539
Variables are keyed by the names TensorFlow would give them (get_variable names of p3d.py, auto-numbered
tf.layers names), so checkpoints / parity dumps line up with the reference.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import tf_semantics as tfs

BLOCK_EXPANSION = 4  # p3d.py:8


class VarStore:
    """Creates variables on first use (TF get_variable semantics) and mimics TF-1.x name uniquification
    for unnamed tf.layers (conv3d, conv3d_1, ... per enclosing variable scope)."""

    def __init__(self, seed: int = 0, dtype=torch.float32, params: Optional[Dict[str, torch.Tensor]] = None,
                 gamma_res=(0.1, 0.3), sa_gamma=(0.3, 0.7)):
        """gamma_res: range of the LAST norm's gamma of every residual branch; sa_gamma: range of the attention gates.  The
        defaults give a network that amplifies perturbations ~200x from stem to output (47 batch-statistics blocks with random
        filters); (0.02, 0.06) / (0.05, 0.15) give a well-conditioned one (residual branches as small as in a
        zero-init-residual / trained ResNet), used by the strict bf16 tolerance tests."""
        self.gamma_res, self.sa_gamma = gamma_res, sa_gamma
        self.rng = np.random.RandomState(seed)
        self.dtype = dtype
        self.params: "OrderedDict[str, torch.Tensor]" = OrderedDict() if params is None else params
        self.frozen = params is not None
        self.counters: Dict[str, int] = {}
        self.trainable: Dict[str, bool] = {}
        self.prefix = ""  # enclosing tf.variable_scope of a whole builder (gn/p3d_gn.py:490)

    def reset_names(self):
        self.counters = {}
        self.prefix = ""

    def unique(self, scope: str, base: str) -> str:
        key = scope + "/" + base
        n = self.counters.get(key, 0)
        self.counters[key] = n + 1
        name = base if n == 0 else f"{base}_{n}"
        return (scope + "/" + name) if scope else name

    def get(self, name: str, shape: Sequence[int], kind: str, trainable: bool = True) -> torch.Tensor:
        name = self.prefix + name
        if name in self.params:
            p = self.params[name]
            assert tuple(p.shape) == tuple(shape), (name, tuple(p.shape), tuple(shape))
            return p
        assert not self.frozen, f"missing variable {name}"
        shape = tuple(int(s) for s in shape)
        if kind == "glorot":  # xavier_initializer()/glorot_uniform with receptive-field fans
            rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            v = self.rng.uniform(-lim, lim, size=shape)
        elif kind == "glorot_t":  # transposed-conv kernel [k..., Cout, Cin]
            rf = int(np.prod(shape[:-2]))
            lim = math.sqrt(6.0 / ((shape[-1] + shape[-2]) * rf))
            v = self.rng.uniform(-lim, lim, size=shape)
        elif kind == "bias":  # synthetic non-zero biases so the bias path is exercised
            v = self.rng.normal(0, 0.05, size=shape)
        elif kind == "gamma":
            v = self.rng.uniform(0.5, 1.5, size=shape)
        elif kind == "gamma_res":  # last BN of a residual branch: small, as in trained / zero-init-residual ResNets,
            v = self.rng.uniform(self.gamma_res[0], self.gamma_res[1], size=shape)  # 
        elif kind == "beta":
            v = self.rng.normal(0, 0.1, size=shape)
        elif kind == "mean":
            v = self.rng.normal(0, 0.1, size=shape)
        elif kind == "var":
            v = self.rng.uniform(0.5, 1.5, size=shape)
        elif kind == "sa_gamma":  # reference init is 0 (network.py:191) which would hide the branch
            v = self.rng.uniform(self.sa_gamma[0], self.sa_gamma[1], size=shape)
        else:
            raise ValueError(kind)
        t = torch.tensor(v, dtype=self.dtype)
        self.params[name] = t
        self.trainable[name] = trainable
        return t


def _ste_round_bf16(t: torch.Tensor) -> torch.Tensor:
    """value rounded to bf16 (round-to-nearest-even, what a bf16 store does), gradient passed straight through"""
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


class Ctx:
    """bf16=True emulates the bf16-STORAGE arithmetic of the CUDA path inside this fp32 oracle: every tensor the CUDA path
    keeps in HBM as bf16 (conv outputs, normalised activations, attention probabilities, tensor-core weight operands) is
    rounded to bf16 at the point where it is stored; accumulation, statistics, scale/shift and the loss stay fp32.  This is
    NOT a different definition of the reference graph -- it is the same graph (same functions below) with storage rounding,
    used by the tests to (a) compare the bf16 CUDA path against an equal-rounding reference and (b) measure how far ANY
    bf16-storage implementation is from the fp32 graph (the distance fp32-oracle <-> bf16-oracle)."""

    def __init__(self, vs: VarStore, training: bool, dropout: float = 0.0, taps: Optional[dict] = None,
                 backbone_training: bool = True, dropout_mask_fn: Optional[Callable] = None, bf16: bool = False):
        self.vs = vs
        self.bf16 = bf16
        self.training = training
        self.dropout = dropout
        self.taps = taps
        self.backbone_training = backbone_training  # make_block is never given `training` (p3d.py:140,350)
        self.new_moving: Dict[str, torch.Tensor] = {}
        self.dropout_mask_fn = dropout_mask_fn

    def tap(self, name, x):
        if self.taps is not None:
            self.taps[name] = x
        return x

    def q(self, t):
        """storage point of an activation / operand"""
        return _ste_round_bf16(t) if self.bf16 else t


# ------------------------------------------------------------------------------------------------
# layer helpers
# ------------------------------------------------------------------------------------------------
def bn(ctx: Ctx, x, training: bool, name: Optional[str] = None, scope: str = "", gamma_kind: str = "gamma"):
    """tf.layers.batch_normalization (auto-named batch_normalization[_N] unless `name`)."""
    nm = name if name is not None else ctx.vs.unique(scope, "batch_normalization")
    c = x.shape[-1]
    g = ctx.vs.get(nm + "/gamma", [c], gamma_kind)
    b = ctx.vs.get(nm + "/beta", [c], "beta")
    mm = ctx.vs.get(nm + "/moving_mean", [c], "mean", trainable=False)
    mv = ctx.vs.get(nm + "/moving_variance", [c], "var", trainable=False)
    if ctx.bf16 and training:
        # CUDA path: (sum, sum of squares) come from the conv's fp32 accumulators, the affine pass reads the bf16-stored raw
        mean = x.mean(dim=(0, 1, 2, 3))
        var = x.var(dim=(0, 1, 2, 3), unbiased=False)
        y = (ctx.q(x) - mean) * torch.rsqrt(var + tfs.BN_EPS) * g + b
        nmm = mm * tfs.BN_MOMENTUM + mean.detach() * (1 - tfs.BN_MOMENTUM)
        nmv = mv * tfs.BN_MOMENTUM + var.detach() * (1 - tfs.BN_MOMENTUM)
    else:
        # (moving-statistics BN of an inference graph is folded into the conv epilogue: applied to the fp32 accumulators)
        y, nmm, nmv = tfs.batch_norm(x, g, b, mm, mv, training)
    if training:
        ctx.new_moving[nm + "/moving_mean"] = nmm
        ctx.new_moving[nm + "/moving_variance"] = nmv
    return y


def gn_layer(ctx: Ctx, x, scope: str = "", gamma_kind: str = "gamma"):
    """GroupNorm (network.py:65-87): tf.Variable gamma/beta inside variable_scope('group_norm') —
    tf.Variable names are uniquified per graph: group_norm/gamma, group_norm_1/gamma, ..."""
    nm = ctx.vs.unique(scope, "group_norm")
    c = x.shape[-1]
    g = ctx.vs.get(nm + "/gamma", [c], gamma_kind)
    b = ctx.vs.get(nm + "/beta", [c], "beta")
    return tfs.group_norm(ctx.q(x), g, b)   # CUDA path: group statistics are taken from the stored raw tensor


def norm(ctx, x, training, mode, scope=""):
    return bn(ctx, x, training, scope=scope) if mode == "bn" else gn_layer(ctx, x, scope)


def conv_w(ctx, name, kshape):  # get_conv_weight (p3d.py:10-16)
    return ctx.vs.get(name, kshape, "glorot" if len(kshape) > 1 else "bias")


def convS(ctx, name, x, cin, cout):  # p3d.py:18-22
    return tfs.conv3d_same(x, ctx.q(conv_w(ctx, name, [1, 3, 3, cin, cout])), (1, 1, 1), conv_w(ctx, name + "_bias", [cout]))


def convT(ctx, name, x, cin, cout):  # p3d.py:23-27
    return tfs.conv3d_same(x, ctx.q(conv_w(ctx, name, [3, 1, 1, cin, cout])), (1, 1, 1), conv_w(ctx, name + "_bias", [cout]))


def layers_conv3d(ctx, x, cout, k, s, name=None, scope="", use_bias=True):
    """tf.layers.conv3d(x, cout, k, s, 'same'[, name]) — kernel/bias variables '<name>/kernel', '<name>/bias'."""
    k = (k, k, k) if isinstance(k, int) else tuple(k)
    s = (s, s, s) if isinstance(s, int) else tuple(s)
    nm = (scope + "/" + name if scope else name) if name is not None else ctx.vs.unique(scope, "conv3d")
    w = ctx.vs.get(nm + "/kernel", [*k, x.shape[-1], cout], "glorot")
    b = ctx.vs.get(nm + "/bias", [cout], "bias") if use_bias else None
    return tfs.conv3d_same(x, ctx.q(w), s, b)


def layers_deconv3d(ctx, x, cout, k, s, name=None, scope=""):
    """tf.layers.conv3d_transpose(x, cout, k, s, 'same'[, name]); kernel [k..., cout, cin]."""
    k = (k, k, k) if isinstance(k, int) else tuple(k)
    s = (s, s, s) if isinstance(s, int) else tuple(s)
    nm = (scope + "/" + name if scope else name) if name is not None else ctx.vs.unique(scope, "conv3d_transpose")
    w = ctx.vs.get(nm + "/kernel", [*k, cout, x.shape[-1]], "glorot_t")
    b = ctx.vs.get(nm + "/bias", [cout], "bias")
    return tfs.conv3d_transpose_same(x, ctx.q(w), s, b)


def net_conv3d(ctx, x, cout, k, s, training, name, mode="bn"):  # network.py:100-104
    return ctx.q(torch.relu(norm(ctx, layers_conv3d(ctx, x, cout, k, s, name), training, mode)))


def net_deconv3d(ctx, x, cout, k, s, training, name, mode="bn"):  # network.py:106-110
    return ctx.q(torch.relu(norm(ctx, layers_deconv3d(ctx, x, cout, k, s, name), training, mode)))


def attention(ctx, x, name, training, mode="bn", subsample=False, sub_size=2):
    """network.py:157-193 (Python-2 integer division at :182,187,188)."""
    n, d, h, w, ch = x.shape
    inter = max(1, ch // 8)
    f = ctx.q(layers_conv3d(ctx, x, inter, 1, 1, scope=name))
    g = ctx.q(layers_conv3d(ctx, x, inter, 1, 1, scope=name))
    hh = ctx.q(layers_conv3d(ctx, x, ch, 1, 1, scope=name))
    if subsample:
        f = tfs.max_pool3d_valid(f, sub_size)
        g = tfs.max_pool3d_valid(g, sub_size // 2)
        hh = tfs.max_pool3d_valid(hh, sub_size)
    gq = g.reshape(n, -1, inter)
    fk = f.reshape(n, -1, inter)
    hv = hh.reshape(n, -1, ch)
    s = torch.matmul(gq, fk.transpose(1, 2))
    beta = ctx.q(torch.softmax(s, dim=-1))      # probabilities are a bf16 tensor-core operand
    o = ctx.q(torch.matmul(beta, hv))
    o = o.reshape(n, d * 2 // sub_size, h * 2 // sub_size, w * 2 // sub_size, ch)
    o = layers_conv3d(ctx, o, ch, 1, sub_size // 2)
    o = ctx.q(torch.relu(norm(ctx, o, training, mode)))
    gamma = ctx.vs.get("gamma" + name, [1], "sa_gamma")
    return ctx.q(o * gamma + x)


def cbam_block(ctx, x, name, ratio=8):
    """network.py:198-274: channel attention (shared MLP on mean & max over D,H,W) then spatial
    attention (7x7x7 conv over [mean_c, max_c], no bias)."""
    n, d, h, w, c = x.shape
    sc = name + "/ch_at"
    w0 = ctx.vs.get(sc + "/mlp_0/kernel", [c, c // ratio], "glorot")
    b0 = ctx.vs.get(sc + "/mlp_0/bias", [c // ratio], "bias")
    w1 = ctx.vs.get(sc + "/mlp_1/kernel", [c // ratio, c], "glorot")
    b1 = ctx.vs.get(sc + "/mlp_1/bias", [c], "bias")
    avg = x.mean(dim=(1, 2, 3))
    mx = x.amax(dim=(1, 2, 3))

    def mlp(v):
        return torch.relu(v @ w0 + b0) @ w1 + b1

    scale = torch.sigmoid(mlp(avg) + mlp(mx)).view(n, 1, 1, 1, c)
    x = x * scale
    sp = name + "/sp_at"
    wk = ctx.vs.get(sp + "/conv3d/kernel", [7, 7, 7, 2, 1], "glorot")
    cat = torch.cat([x.mean(dim=4, keepdim=True), x.amax(dim=4, keepdim=True)], dim=4)
    att = torch.sigmoid(tfs.conv3d_same(cat, wk, (1, 1, 1)))
    return x * att


def dropout(ctx, x, name):
    """tf.layers.dropout(x, rate, training): inverted scaling; the mask comes from dropout_mask_fn so
    the CUDA path and the oracle can share it (TF's Philox stream is not reproducible anyway)."""
    if not ctx.training or ctx.dropout <= 0.0:
        return x
    keep = ctx.dropout_mask_fn(name, x.shape).to(x.dtype)
    return ctx.q(x * keep / (1.0 - ctx.dropout))


# ------------------------------------------------------------------------------------------------
# backbone (p3d.py:30-166; gn/p3d_gn.py:74-209)
# ------------------------------------------------------------------------------------------------
def bottleneck(ctx: Ctx, x, inplanes, planes, idx, first_of_stage, mode):
    """Bottleneck.infer for n_s < depth_3d (the 2-D branches are dead code: 47 blocks == depth_3d)."""
    tr = ctx.backbone_training
    stride_hw = 2 if (first_of_stage and idx != 0) else 1  # p3d.py:45-49
    nrm = (lambda t, gk="gamma": bn(ctx, t, tr, gamma_kind=gk)) if mode == "bn" else (lambda t, gk="gamma": gn_layer(ctx, t, gamma_kind=gk))
    q = ctx.q
    residual = x
    out = tfs.conv3d_same(x, q(conv_w(ctx, f"conv3_{idx}_1", [1, 1, 1, inplanes, planes])), (1, stride_hw, stride_hw))
    out = q(torch.relu(nrm(out)))
    st = "ABC"[idx % 3]
    nm = f"ST{st}_{idx}_2"
    if st == "A":  # serial S -> T
        out = q(torch.relu(nrm(convS(ctx, nm + "_S", out, planes, planes))))
        out = q(torch.relu(nrm(convT(ctx, nm + "_T", out, planes, planes))))
    elif st == "B":  # parallel S + T  (CUDA path: both norms + ReLUs + the sum are one pass, stored once)
        s_br = torch.relu(nrm(convS(ctx, nm + "_S", out, planes, planes)))
        t_br = torch.relu(nrm(convT(ctx, nm + "_T", out, planes, planes)))
        out = q(t_br + s_br)
    else:  # C: S then S + T(S)
        s_br = q(torch.relu(nrm(convS(ctx, nm + "_S", out, planes, planes))))
        t_br = torch.relu(nrm(convT(ctx, nm + "_T", s_br, planes, planes)))
        out = q(s_br + t_br)
    out = nrm(tfs.conv3d_same(out, q(conv_w(ctx, f"conv3_{idx}_3", [1, 1, 1, planes, planes * BLOCK_EXPANSION]))), "gamma_res")
    if first_of_stage:  # downsample=['3d', stride_p] (p3d.py:149-155,124-127)
        residual = tfs.conv3d_same(residual, q(conv_w(ctx, f"dw3d_{idx}", [1, 1, 1, inplanes, planes * BLOCK_EXPANSION])),
                                   (1, stride_hw, stride_hw))
        residual = nrm(residual)
        if mode == "gn":
            residual = q(residual)   # GN graphs store the normalised shortcut (it feeds the CBAM kernels); BN graphs fuse it
    if mode == "gn":
        residual = cbam_block(ctx, residual, f"cbam_{idx}")  # gn/p3d_gn.py:175
    return ctx.tap(f"b{idx}", q(torch.relu(out + residual)))


def make_block(ctx, x, planes, num, inplanes, cnt, mode):
    x = bottleneck(ctx, x, inplanes, planes, cnt, True, mode)
    for i in range(1, num):
        x = bottleneck(ctx, x, planes * BLOCK_EXPANSION, planes, cnt + i, False, mode)
    return x, cnt + num


TPOOL = ((2, 1, 1), (2, 1, 1))


def backbone(ctx: Ctx, x, mode="bn", stem_training=None):
    """stem + 3 stages; returns dict of the tensors the decoders consume."""
    w = ctx.q(conv_w(ctx, "firstconv1", [1, 7, 7, 3, 64]))
    c1 = ctx.tap("firstconv1", tfs.conv3d_same(ctx.q(x), w, (1, 2, 2)))
    if mode == "bn":
        c1 = bn(ctx, c1, ctx.training if stem_training is None else stem_training)  # p3d.py:344 follows `training`
    else:
        c1 = gn_layer(ctx, c1)
    c1 = ctx.tap("stem", ctx.q(torch.relu(c1)))
    t = {}
    t["x_1_0"] = ctx.tap("x_1_0", tfs.max_pool3d_same(c1, *TPOOL))
    pool1 = ctx.tap("pool1", tfs.max_pool3d_same(c1, (2, 3, 3), (2, 2, 2)))
    res1, cnt = make_block(ctx, pool1, 64, 3, 64, 0, mode)
    t["x_2_0"] = ctx.tap("x_2_0", tfs.max_pool3d_same(res1, *TPOOL))
    res2, cnt = make_block(ctx, t["x_2_0"], 128, 8, 256, cnt, mode)
    t["x_3_0"] = ctx.tap("x_3_0", tfs.max_pool3d_same(res2, *TPOOL))
    res3, cnt = make_block(ctx, t["x_3_0"], 256, 36, 512, cnt, mode)
    t["x_4_0"] = ctx.tap("x_4_0", tfs.max_pool3d_same(res3, *TPOOL))
    return t


# ------------------------------------------------------------------------------------------------
# decoders
# ------------------------------------------------------------------------------------------------
def _unetpp(ctx, x, sa: bool):
    tr = ctx.training
    t = backbone(ctx, x)
    x10, x20, x30, x40 = t["x_1_0"], t["x_2_0"], t["x_3_0"], t["x_4_0"]
    cat = lambda a, b: torch.cat([a, b], dim=-1)  # noqa: E731
    if sa:
        x40 = ctx.tap("x_4_0_sa", attention(ctx, x40, "x_4_0_sa", tr))
    up40 = ctx.tap("upx_4_0", net_deconv3d(ctx, x40, 512, (1, 3, 3), 2, tr, "upx_4_0"))
    x31 = ctx.tap("x_3_1", net_conv3d(ctx, cat(x30, up40), 512, (2, 3, 3), 1, tr, "x_3_1"))
    if sa:
        x31 = ctx.tap("x_3_1_sa", attention(ctx, x31, "x_3_1_sa", tr))
    up30 = ctx.tap("upx_3_0", net_deconv3d(ctx, x30, 256, (2, 3, 3), 2, tr, "upx_3_0"))
    x21 = ctx.tap("x_2_1", net_conv3d(ctx, cat(x20, up30), 256, 3, 1, tr, "x_2_1"))
    up31 = ctx.tap("upx_3_1", net_deconv3d(ctx, x31, 256, (2, 3, 3), 2, tr, "upx_3_1"))
    x22 = ctx.tap("x_2_2", net_conv3d(ctx, cat(x21, up31), 256, 3, 1, tr, "x_2_2"))
    if sa:
        x22 = ctx.tap("x_2_2_sa", attention(ctx, x22, "x_2_2_sa", tr))
    up20 = ctx.tap("upx_2_0", net_deconv3d(ctx, x20, 128, 3, 2, tr, "upx_2_0"))
    x11 = ctx.tap("x_1_1", net_conv3d(ctx, cat(x10, up20), 128, 3, 1, tr, "x_1_1"))
    up21 = ctx.tap("upx_2_1", net_deconv3d(ctx, x21, 128, 3, 2, tr, "upx_2_1"))
    x12 = ctx.tap("x_1_2", net_conv3d(ctx, cat(x11, up21), 128, 3, 1, tr, "x_1_2"))
    up22 = ctx.tap("upx_2_2", net_deconv3d(ctx, x22, 128, 3, 2, tr, "upx_2_2"))
    x13 = ctx.tap("x_1_3", net_conv3d(ctx, cat(x12, up22), 128, 3, 1, tr, "x_1_3"))
    if sa:
        x13 = ctx.tap("x_1_3_sa", attention(ctx, x13, "x_1_3_sa", tr, subsample=True))
    x13 = dropout(ctx, x13, "x_1_3_drop")
    logits = ctx.tap("x_0_1", layers_deconv3d(ctx, x13, 1, 3, 2, "x_0_1"))
    return ctx.tap("pred", torch.sigmoid(logits))


def p3d_unetplusplus_ds(ctx, x):  # p3d.py:340-399
    return _unetpp(ctx, x, True)


def p3d_unetplusplus_nonsa(ctx, x):  # p3d.py:401-459
    return _unetpp(ctx, x, False)


def p3d_unet(ctx, x):  # p3d.py:169-221
    tr = ctx.training
    t = backbone(ctx, x)
    cat = lambda a, b: torch.cat([a, b], dim=-1)  # noqa: E731
    d1 = ctx.q(torch.relu(bn(ctx, layers_deconv3d(ctx, t["x_4_0"], 512, (1, 3, 3), 2), tr, name="deconv1_bn")))
    d2 = ctx.q(torch.relu(bn(ctx, layers_deconv3d(ctx, cat(d1, t["x_3_0"]), 256, (2, 3, 3), 2), tr, name="deconv2_bn")))
    d3 = ctx.q(torch.relu(bn(ctx, layers_deconv3d(ctx, cat(d2, t["x_2_0"]), 128, 3, 2), tr, name="deconv3_bn")))
    d3 = ctx.tap("deconv3", dropout(ctx, d3, "deconv3_drop"))  # deconv3_concat is computed and ignored (:213-214)
    c = ctx.q(layers_conv3d(ctx, d3, 32, 1, 1))
    logits = ctx.tap("x_0_1", layers_deconv3d(ctx, c, 1, 3, 2))
    return ctx.tap("pred", torch.sigmoid(logits))


def p3d_concat(ctx, x):  # p3d.py:224-276 (returns logits: no sigmoid at :275-276)
    tr = ctx.training
    w = ctx.q(conv_w(ctx, "firstconv1", [1, 7, 7, 3, 64]))
    c1 = ctx.q(torch.relu(bn(ctx, tfs.conv3d_same(ctx.q(x), w, (1, 2, 2)), tr)))
    pool1 = tfs.max_pool3d_same(c1, (2, 3, 3), (2, 2, 2))
    res1, cnt = make_block(ctx, pool1, 64, 3, 64, 0, "bn")
    pool2 = tfs.max_pool3d_same(res1, *TPOOL)
    dp2 = ctx.q(torch.relu(bn(ctx, layers_deconv3d(ctx, pool2, 128, 3, 1, "deconv_pool2"), tr, name="deconv_pool2_bn")))
    res2, cnt = make_block(ctx, pool2, 128, 8, 256, cnt, "bn")
    pool3 = tfs.max_pool3d_same(res2, *TPOOL)
    dp3 = ctx.q(torch.relu(bn(ctx, layers_deconv3d(ctx, pool3, 256, 3, 2, "deconv_pool3"), tr, name="deconv_pool3_bn")))
    res3, cnt = make_block(ctx, pool3, 256, 36, 512, cnt, "bn")
    pool4 = tfs.max_pool3d_same(res3, *TPOOL)
    dp4 = ctx.q(torch.relu(bn(ctx, layers_deconv3d(ctx, pool4, 512, 3, 4, "deconv_pool4"), tr, name="deconv_pool4_bn")))
    cc = torch.cat([dp2, dp3, dp4], dim=-1)
    cc = ctx.q(torch.relu(bn(ctx, layers_conv3d(ctx, cc, 512, 3, 1, "conv_concat"), tr, name="conv_concat_bn")))
    dr = ctx.q(torch.relu(bn(ctx, layers_deconv3d(ctx, cc, 128, 3, 2, "deconv_revise"), tr, name="deconv1_revise_bn")))
    dr = dropout(ctx, dr, "deconv_revise_drop")
    return ctx.tap("pred", layers_deconv3d(ctx, dr, 1, 3, 2, "predict_revise"))


def gn_inference_p3d(ctx, x, pool4_filters=1024):  # gn/p3d_gn.py:214-258 (and :279-324 with 512)
    w = ctx.q(conv_w(ctx, "firstconv1", [1, 7, 7, 3, 64]))
    c1 = ctx.q(torch.relu(gn_layer(ctx, tfs.conv3d_same(ctx.q(x), w, (1, 2, 2)))))
    pool1 = tfs.max_pool3d_same(c1, (2, 3, 3), (2, 2, 2))
    res1, cnt = make_block(ctx, pool1, 64, 3, 64, 0, "gn")
    pool2 = ctx.tap("pool2", tfs.max_pool3d_same(res1, *TPOOL))
    res2, cnt = make_block(ctx, pool2, 128, 8, 256, cnt, "gn")
    pool3 = ctx.tap("pool3", tfs.max_pool3d_same(res2, *TPOOL))
    dp3 = ctx.q(torch.relu(gn_layer(ctx, layers_deconv3d(ctx, pool3, 512, 3, 2, "deconv_pool3"))))
    res3, cnt = make_block(ctx, pool3, 256, 36, 512, cnt, "gn")
    pool4 = ctx.tap("pool4", tfs.max_pool3d_same(res3, *TPOOL))
    dp4 = ctx.q(torch.relu(gn_layer(ctx, layers_deconv3d(ctx, pool4, pool4_filters, 3, 4, "deconv_pool4"))))
    cc = torch.cat([dp3, dp4, pool2], dim=-1)
    cc = ctx.tap("conv_concat", ctx.q(torch.relu(gn_layer(ctx, layers_conv3d(ctx, cc, 1024, 3, 1, "conv_concat")))))
    dr = ctx.q(torch.relu(gn_layer(ctx, layers_deconv3d(ctx, cc, 256, 3, 2, "deconv_revise"))))
    dr = dropout(ctx, dr, "deconv_revise_drop")
    return ctx.tap("pred", layers_deconv3d(ctx, dr, 1, 3, 2, "predict_revise"))


def gn_inference_p3d_concat(ctx, x):
    return gn_inference_p3d(ctx, x, pool4_filters=512)


def gn_inference_p3d_decoder_block(ctx, x):  # gn/p3d_gn.py:489-539 (inside tf.variable_scope('P3D'))
    ctx.vs.prefix = "P3D/"
    gnrelu = lambda t: ctx.q(torch.relu(gn_layer(ctx, t)))  # noqa: E731
    w = ctx.q(conv_w(ctx, "firstconv1", [1, 7, 7, 3, 64]))
    c1 = gnrelu(tfs.conv3d_same(ctx.q(x), w, (1, 2, 2)))
    pool1 = tfs.max_pool3d_same(c1, (2, 3, 3), (2, 2, 2))
    res1, cnt = make_block(ctx, pool1, 64, 3, 64, 0, "gn")
    pool2 = ctx.tap("pool2", tfs.max_pool3d_same(res1, *TPOOL))
    dp2 = gnrelu(layers_deconv3d(ctx, pool2, 128, 3, 1, "deconv_pool2"))
    res2, cnt = make_block(ctx, pool2, 128, 8, 256, cnt, "gn")
    pool3 = ctx.tap("pool3", tfs.max_pool3d_same(res2, *TPOOL))
    dp3 = gnrelu(layers_deconv3d(ctx, pool3, 256, (2, 3, 3), 2, "deconv_pool3"))
    res3, cnt = make_block(ctx, pool3, 256, 36, 512, cnt, "gn")
    pool4 = ctx.tap("pool4", tfs.max_pool3d_same(res3, *TPOOL))
    dp4 = gnrelu(layers_deconv3d(ctx, pool4, 512, (1, 3, 3), 4, "deconv_pool4"))
    cc = ctx.tap("conv_concat", gnrelu(layers_conv3d(ctx, torch.cat([dp2, dp3, dp4], dim=-1), 1024, 3, 1, "conv_concat")))
    d = gnrelu(layers_conv3d(ctx, cc, 256, 3, 1, "decoder1_conv1"))
    d = gnrelu(layers_deconv3d(ctx, d, 256, 3, 2, "decoder1_deconv"))
    d = gnrelu(layers_conv3d(ctx, d, 128, 3, 1, "decoder1_conv2"))
    d = gnrelu(layers_conv3d(ctx, d, 32, 3, 1, "decoder2_conv1"))
    d = gnrelu(layers_deconv3d(ctx, d, 32, 3, 2, "decoder2_deconv"))
    d = ctx.tap("decoder2_conv2", gnrelu(layers_conv3d(ctx, d, 16, 3, 1, "decoder2_conv2")))
    d = dropout(ctx, d, "final_drop")
    return ctx.tap("pred", layers_conv3d(ctx, d, 1, 3, 1, "results"))


GRAPHS = {
    "p3d_unetplusplus_ds": p3d_unetplusplus_ds,
    "p3d_unetplusplus_nonsa": p3d_unetplusplus_nonsa,
    "p3d_unet": p3d_unet,
    "p3d_concat": p3d_concat,
    "inference_p3d": gn_inference_p3d,
    "inference_p3d_concat": gn_inference_p3d_concat,
    "inference_p3d_decoder_block": gn_inference_p3d_decoder_block,
}


# ------------------------------------------------------------------------------------------------
# drivers
# ------------------------------------------------------------------------------------------------
def synthetic_clip(batch: int, frames: int = 16, size: int = 112, seed: int = 0, dtype=torch.float32):
    """mapf of dataflow.py:194-209: (uint8 RGB - [90,102,98]) / 255 (BGR mean [98,102,90] reversed)."""
    rng = np.random.RandomState(seed)
    u = rng.randint(0, 256, size=(batch, frames, size, size, 3)).astype(np.float32)
    x = (u - np.array([90.0, 102.0, 98.0], dtype=np.float32)) / 255.0
    return torch.tensor(x, dtype=dtype)


def synthetic_target(batch: int, frames: int = 16, size: int = 112, seed: int = 1, dtype=torch.float32):
    rng = np.random.RandomState(seed)
    return torch.tensor(rng.randint(0, 256, size=(batch, frames, size, size)).astype(np.float32) / 255.0, dtype=dtype)


def forward(graph: str, x: torch.Tensor, vs: VarStore, training: bool, dropout_rate: float = 0.0,
            taps: Optional[dict] = None, dropout_mask_fn=None, bf16: bool = False) -> torch.Tensor:
    vs.reset_names()
    ctx = Ctx(vs, training, dropout_rate, taps, dropout_mask_fn=dropout_mask_fn, bf16=bf16)
    out = GRAPHS[graph](ctx, x)
    forward.last_ctx = ctx
    return out


def train_step(graph: str, x, y, vs: VarStore, adam_state: dict, step: int, lr: float = 1e-4, dropout_rate: float = 0.0,
               dropout_mask_fn=None, bf16: bool = False, taps: Optional[dict] = None):
    """One iteration of train.py:156-172,217: smooth-L1 (sum) loss, Adam(lr), BN moving-average updates.
    Returns (loss, grads dict).  Updates vs.params and adam_state in place."""
    names = [n for n in vs.params if vs.trainable.get(n, True)]
    leaves = {}
    for n in names:
        leaves[n] = vs.params[n].detach().clone().requires_grad_(True)
        vs.params[n] = leaves[n]
    was_frozen = vs.frozen
    vs.frozen = True
    pred = forward(graph, x, vs, True, dropout_rate, taps=taps, dropout_mask_fn=dropout_mask_fn, bf16=bf16)
    train_step.last_pred = pred.detach()
    ctx = forward.last_ctx
    loss = tfs.smooth_l1_loss(pred.reshape(y.shape), y, 1.0)
    loss.backward()
    grads = {}
    with torch.no_grad():
        for n in names:
            g = leaves[n].grad if leaves[n].grad is not None else torch.zeros_like(leaves[n])
            grads[n] = g
            m = adam_state.setdefault("m/" + n, torch.zeros_like(g))
            v = adam_state.setdefault("v/" + n, torch.zeros_like(g))
            p, m2, v2 = tfs.adam_step_tf(leaves[n].detach(), g, m, v, step, lr)
            vs.params[n] = p
            adam_state["m/" + n] = m2
            adam_state["v/" + n] = v2
        for n, v in ctx.new_moving.items():
            vs.params[n] = v.detach()
    vs.frozen = was_frozen
    return float(loss.detach()), grads
