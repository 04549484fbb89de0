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


def normalize(x: ConvOut, training, mode="bn"):  # utils/network.py:89 (without the ReLU)
    if mode != "bn":
        raise NotImplementedError("GroupNorm graphs are built by gn/p3d_gn.py")
    return bn_relu(x, training, relu=False)


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


def smooth_l1_loss(*_args, **_kw):  # utils/network.py:49 — fused into the head's backward (engine._HeadOp)
    raise NotImplementedError("the smooth-L1 loss is fused into Session.train_step (sap3d_loss_smooth_l1)")
