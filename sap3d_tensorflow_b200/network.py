"""Layer helpers with the public names of the reference's utils/network.py, lowered onto the engine's
fused primitives (conv + statistics epilogue -> finalize -> fused affine/ReLU/residual pass).

Reference surface mirrored (utils/network.py): pool3d :6, smooth_l1_loss :49, GroupNorm :65,
normalize :89, concat :97, conv3d :100, transpose_conv3d :106, attention :157, cbam_block :198.
Handles are engine tensors (`engine.T`); `concat` returns a lazy channel concatenation that the
convolution kernels consume as K-segments (it is never materialised).
"""
from __future__ import annotations

from typing import List, Sequence, Union

from . import _abi as A
from .engine import ConvOut, Engine, T

DEFAULT_PADDING = "SAME"


class Cat:
    """lazy tf.concat(xs, axis=-1)"""

    def __init__(self, parts: Sequence[T]):
        self.parts: List[T] = list(parts)
        self.eng: Engine = parts[0].eng

    @property
    def shape(self):
        s = list(self.parts[0].shape)
        s[-1] = sum(p.C for p in self.parts)
        return tuple(s)


Handle = Union[T, Cat]


def _parts(x: Handle) -> List[T]:
    return x.parts if isinstance(x, Cat) else [x]


def _k3(k):
    return (k, k, k) if isinstance(k, int) else tuple(k)


def concat(x: Sequence[T]) -> Cat:  # utils/network.py:97
    return Cat(x)


def pool3d(value: T, sub_size: int) -> T:  # utils/network.py:6 (tf.layers.max_pooling3d, 'valid')
    if sub_size == 1:
        return value
    return value.eng.maxpool(value, _k3(sub_size), _k3(sub_size), same=False, name="pool3d")


# ---- normalisation -------------------------------------------------------------------------------
def _bn_state(eng: Engine, C: int, name=None, scope=""):
    """variables of one tf.layers.batch_normalization call (auto-named unless `name`)"""
    nm = name if name is not None else eng.names.unique(scope, "batch_normalization")
    return eng.norm_state(C, eng.param(nm + "/gamma", [C], "ones"), eng.param(nm + "/beta", [C], "zeros"),
                          eng.param(nm + "/moving_mean", [C], "mean", trainable=False),
                          eng.param(nm + "/moving_variance", [C], "var", trainable=False))


def bn_relu(co: ConvOut, training: bool, name=None, relu=True, tap: str = "") -> T:
    eng = co.raw.eng
    ns = _bn_state(eng, co.raw.C, name)
    return eng.norm_act(co, ns, training, relu, name=tap or (name or "bn"))


def _as_conv_out(x) -> ConvOut:
    """normalisation helpers take the raw output of a conv (ConvOut) or any activation tensor"""
    return x if isinstance(x, ConvOut) else ConvOut(x, None, 0)


def gn_state(t: T):
    """variables + statistics buffers of one GroupNorm call: tf.Variable gamma / beta inside variable_scope('group_norm'),
    uniquified per graph as group_norm, group_norm_1, ... (utils/network.py:69,78-79)"""
    from .engine_gn import GNState
    eng = t.eng
    nm = eng.names.unique("", "group_norm")
    N, Cc = t.shape[0], t.C
    return GNState(eng, N, Cc, t.positions // N, eng.param(nm + "/gamma", [Cc], "ones"), eng.param(nm + "/beta", [Cc], "zeros"))


def GroupNorm(x, G=32, esp=1e-5, relu=False, name="") -> T:  # utils/network.py:65-87 (== gn/p3d_gn.py:24-46)
    """per-sample statistics over (C/G, D, H, W) with G = min(32, C), biased variance, per-channel gamma / beta"""
    from .engine_gn import GN_EPS, GN_GROUPS, GNActOp
    if G != GN_GROUPS or abs(esp - GN_EPS) > 1e-12:
        raise A.Sap3dError("GroupNorm: the kernels implement the reference's only configuration (G = 32, esp = 1e-5)")
    co = _as_conv_out(x)
    return GNActOp(co.raw.eng, co, gn_state(co.raw), relu, name=name or "group_norm").y


def normalize(x, training, mode="bn"):  # utils/network.py:89-94 (without the ReLU)
    if mode == "bn":
        co = _as_conv_out(x)
        if co.stats is None:
            raise A.Sap3dError("normalize(mode='bn'): batch statistics come from the producing conv's epilogue; pass the conv output")
        return bn_relu(co, training, relu=False)
    if mode == "gn":
        return GroupNorm(x)
    raise ValueError(mode)


# ---- conv / deconv + norm + relu -----------------------------------------------------------------
def layers_conv3d(x: Handle, channel: int, kernel, strides, name=None, scope="", use_bias=True, want_stats=True,
                  bias_grad=True) -> ConvOut:
    """tf.layers.conv3d(x, channel, kernel, strides, 'same', name=name): '<name>/kernel' DHWIO, '<name>/bias'"""
    parts = _parts(x)
    eng = parts[0].eng
    k, s = _k3(kernel), _k3(strides)
    nm = ((scope + "/" + name) if scope else name) if name is not None else eng.names.unique(scope, "conv3d")
    cin = sum(p.C for p in parts)
    w = eng.param(nm + "/kernel", [*k, cin, channel], "glorot")
    b = eng.param(nm + "/bias", [channel], "zeros") if use_bias else None
    return eng.conv(parts, channel, k, s, w, b, transposed=False, want_stats=want_stats, name=nm, bias_grad=bias_grad)


def layers_conv3d_transpose(x: Handle, channel: int, kernel, strides, name=None, scope="", want_stats=True,
                            bias_grad=True) -> ConvOut:
    """tf.layers.conv3d_transpose(x, channel, kernel, strides, 'same', name=name): kernel [k..., Cout, Cin]"""
    parts = _parts(x)
    eng = parts[0].eng
    k, s = _k3(kernel), _k3(strides)
    nm = ((scope + "/" + name) if scope else name) if name is not None else eng.names.unique(scope, "conv3d_transpose")
    cin = sum(p.C for p in parts)
    w = eng.param(nm + "/kernel", [*k, channel, cin], "glorot_t")
    b = eng.param(nm + "/bias", [channel], "zeros")
    return eng.conv(parts, channel, k, s, w, b, transposed=True, want_stats=want_stats, name=nm, bias_grad=bias_grad)


def conv3d(x: Handle, channel, kernel, strides, training, name, mode="bn") -> T:  # utils/network.py:100
    co = layers_conv3d(x, channel, kernel, strides, name, bias_grad=not training)  # bias in front of batch-stat BN
    return co.raw.eng.tap(name, bn_relu(co, training, tap=name))


def transpose_conv3d(x: Handle, channel, kernel, strides, training, name, mode="bn") -> T:  # utils/network.py:106
    co = layers_conv3d_transpose(x, channel, kernel, strides, name, bias_grad=not training)
    return co.raw.eng.tap(name, bn_relu(co, training, tap=name))


def attention(x: T, name, training, mode="bn", subsample=False, sub_size=2) -> T:
    """utils/network.py:157-193 (SAGAN-style self-attention; Python-2 integer division at :182,187,188).
    f, g: 1x1x1 conv -> max(1, C/8) channels, h: 1x1x1 conv -> C (inside variable_scope(name));
    optional max-pool of f/h (sub_size) and g (sub_size/2); o = softmax(g f^T) h -> 1x1x1 conv
    (stride sub_size/2) -> BN -> ReLU; result = o * gamma + x with the scalar variable 'gamma'+name."""
    eng = x.eng
    ch = x.C
    inter = max(1, ch // 8)
    def proj(cout):
        """f / g projection (tf.layers.conv3d 1x1x1, utils/network.py:164-173).  In the bf16 path 16- and 32-channel
        projections are executed with 64 output channels (zero-padded filters) so they stay on the tensor cores;
        the extra channels are exactly zero and do not change g.f^T."""
        nm = eng.names.unique(name, "conv3d")
        w = eng.param(nm + "/kernel", [1, 1, 1, ch, cout], "glorot")
        b = eng.param(nm + "/bias", [cout], "zeros")
        if eng.dt == A.BF16 and cout % 64 != 0:
            return eng.conv_padded_cout([x], (cout + 63) // 64 * 64, (1, 1, 1), (1, 1, 1), w, b, name=nm).raw
        return eng.conv([x], cout, (1, 1, 1), (1, 1, 1), w, b, want_stats=False, name=nm).raw

    f = proj(inter)
    g = proj(inter)
    h = layers_conv3d(x, ch, 1, 1, scope=name, want_stats=False).raw
    if subsample:
        f = pool3d(f, sub_size)
        g = pool3d(g, sub_size // 2)
        h = pool3d(h, sub_size)
    o = eng.attn_core(g, f, h, name=name)
    oc = layers_conv3d(o, ch, 1, sub_size // 2, bias_grad=not training)
    o = bn_relu(oc, training, tap=name + "/o")
    gamma = eng.param("gamma" + name, [1], "sa_gamma")
    return eng.gate(o, x, gamma, name=name)


# ---- CBAM (utils/network.py:198-274) ------------------------------------------------------------------
def _cbam_params(eng: Engine, name: str, Cc: int, ratio: int, channel: bool, spatial: bool):
    if ratio != 8:
        raise A.Sap3dError("cbam_block: the kernels implement the reference's only configuration (ratio = 8)")
    w0 = b0 = w1 = b1 = w_sp = None
    if channel:   # channel_attention: tf.layers.dense mlp_0 / mlp_1 inside variable_scope(name + '/ch_at' ...)
        w0 = eng.param(name + "/mlp_0/kernel", [Cc, Cc // ratio], "vscale")
        b0 = eng.param(name + "/mlp_0/bias", [Cc // ratio], "zeros")
        w1 = eng.param(name + "/mlp_1/kernel", [Cc // ratio, Cc], "vscale")
        b1 = eng.param(name + "/mlp_1/bias", [Cc], "zeros")
    return w0, b0, w1, b1, w_sp


def channel_attention(input_feature: T, name, ratio=8) -> T:  # utils/network.py:208-249
    """mean & max over D,H,W -> shared MLP C -> C/ratio (ReLU) -> C -> sum -> sigmoid -> scale"""
    from .engine_gn import CbamOp
    eng = input_feature.eng
    w0, b0, w1, b1, _ = _cbam_params(eng, name, input_feature.C, ratio, True, False)
    return CbamOp(eng, input_feature, "channel", w0, b0, w1, b1, None, name=name).y


def spatial_attention(input_feature: T, name) -> T:  # utils/network.py:251-274
    """mean & max over C -> 7x7x7 conv (2 -> 1, no bias) -> sigmoid -> scale"""
    from .engine_gn import CbamOp
    eng = input_feature.eng
    w_sp = eng.param(name + "/conv3d/kernel", [7, 7, 7, 2, 1], "vscale")
    return CbamOp(eng, input_feature, "spatial", w_sp=w_sp, name=name).y


def cbam_block(input_feature: T, name, ratio=8) -> T:  # utils/network.py:198-206
    """channel attention then spatial attention on any feature map, as ONE fused op (variables '<name>/ch_at/mlp_{0,1}/...',
    '<name>/sp_at/conv3d/kernel').  The GN backbone uses the block-tail form that also folds the residual add + ReLU
    (engine_gn.CbamBlockTailOp); this is the stand-alone call of the reference surface."""
    from .engine_gn import CbamOp
    eng = input_feature.eng
    w0, b0, w1, b1, _ = _cbam_params(eng, name + "/ch_at", input_feature.C, ratio, True, False)
    w_sp = eng.param(name + "/sp_at/conv3d/kernel", [7, 7, 7, 2, 1], "vscale")
    return CbamOp(eng, input_feature, "both", w0, b0, w1, b1, w_sp, name=name).y


# ---- loss --------------------------------------------------------------------------------------------------
class Loss:
    """result of smooth_l1_loss: what `Session` trains on (the reference hands the loss tensor to AdamOptimizer.minimize,
    train.py:166-168).  `head` is the network output op it was built from."""

    def __init__(self, head, sigma, inside, outside):
        self.head, self.eng = head, head.eng
        self.sigma, self.inside, self.outside = sigma, inside, outside

    @property
    def output(self):
        return self.head.output


def smooth_l1_loss(bbox_pred, bbox_targets, bbox_inside_weights, bbox_outside_weights, sigma=3.0, dim=[0]):  # noqa: B006
    """utils/network.py:49-62 (call sites train.py:159, gn/train_p3d_gn_dataset.py:186, both with weights 1, sigma 1):
    in = inside * (pred - target); per element |in| < 1/sigma^2 ? in^2 sigma^2 / 2 : |in| - 0.5/sigma^2; times outside;
    reduce_sum over everything (the reduce_mean of that scalar is the identity; `dim` is unused in the reference too).

    bbox_pred is the handle a graph builder returned (the tf.reshape the drivers wrap around it is a view: the loss is over
    all elements either way); bbox_targets is fed per step through Session.train_step(x, y) and may be None here.  The
    weights are the reference's scalars.  Records the loss configuration on the output op, whose backward launch is the fused
    loss + gradient kernel (sap3d_loss_smooth_l1_ex)."""
    head = bbox_pred.head if isinstance(bbox_pred, Loss) else bbox_pred
    if not hasattr(head, "loss_sigma"):
        raise A.Sap3dError("smooth_l1_loss: the first argument must be the output handle of a graph builder")
    if not head.eng.training_graph:
        raise A.Sap3dError("smooth_l1_loss needs a training graph (placeholder(..., training_graph=True))")
    for wgt in (bbox_inside_weights, bbox_outside_weights):
        if not isinstance(wgt, (int, float)):
            raise A.Sap3dError("smooth_l1_loss: inside / outside weights are scalars (the reference passes 1, 1)")
    head.loss_sigma, head.loss_w_in, head.loss_w_out = float(sigma), float(bbox_inside_weights), float(bbox_outside_weights)
    return Loss(head, float(sigma), float(bbox_inside_weights), float(bbox_outside_weights))
