"""P3D-199 (minus layer4) backbone + saliency decoders on the B200 engine.

Public surface = the reference's p3d.py: get_conv_weight :10, convS :18, convT :23, Bottleneck :30,
make_block :139, p3d_unet :169, p3d_concat :224, p3d_unetplusplus_ds :340, p3d_unetplusplus_nonsa :401
(same argument lists; `_X` is a `placeholder`, the result is an engine handle consumed by `Session`).
Variable names are the reference's (firstconv1, conv3_{id}_{1,3}, ST{A,B,C}_{id}_2_{S,T}[_bias],
dw3d_{id}, auto-numbered batch_normalization_N, named decoder layers) so checkpoints line up.

The wiring is table-driven; every conv -> BN -> ReLU (-> add) chain is lowered to
  tcgen05 implicit-GEMM (statistics in the epilogue) -> bn_finalize -> one fused apply pass.
"""
from __future__ import annotations

from typing import Optional

from . import network as nw
from .engine import ConvOut, Engine, T
from .engine_gn import ConcatOp

CROP_SIZE = 112
NUM_FRAMES_PER_CLIP = 16
RGB_CHANNEL = 3
BLOCK_EXPANSION = 4

# (planes, number of bottlenecks, inplanes, spatial stride of the first bottleneck)  p3d.py:350-363
STAGES = ((64, 3, 64, 1), (128, 8, 256, 2), (256, 36, 512, 2))
TEMPORAL_POOL = ((2, 1, 1), (2, 1, 1))


def get_conv_weight(eng: Engine, name, kshape, wd=0.001):
    """p3d.py:10-16.  Xavier-uniform variable; the weight-decay collection is never added to the loss
    in the reference (train.py:161-162), so `wd` is accepted and ignored."""
    return eng.param(name, kshape, "glorot" if len(kshape) > 1 else "xavier1d")


def convS(name, l_input: T, in_channels, out_channels, bias_grad=False) -> ConvOut:  # p3d.py:18-22: 1x3x3 + bias
    """bias_grad=False: under BatchNorm the bias gradient is exactly zero (engine.conv); the GroupNorm variant passes True"""
    eng = l_input.eng
    return eng.conv([l_input], out_channels, (1, 3, 3), (1, 1, 1), get_conv_weight(eng, name, [1, 3, 3, in_channels, out_channels]),
                    get_conv_weight(eng, name + "_bias", [out_channels], 0), name=name, bias_grad=bias_grad)


def convT(name, l_input: T, in_channels, out_channels, bias_grad=False) -> ConvOut:  # p3d.py:23-27: 3x1x1 + bias
    eng = l_input.eng
    return eng.conv([l_input], out_channels, (3, 1, 1), (1, 1, 1), get_conv_weight(eng, name, [3, 1, 1, in_channels, out_channels]),
                    get_conv_weight(eng, name + "_bias", [out_channels], 0), name=name, bias_grad=bias_grad)


class Bottleneck:
    """p3d.py:30-136 for the 3-D branch (n_s < depth_3d; the 2-D branches are unreachable with 47 blocks)."""

    def __init__(self, l_input: T, inplanes, planes, stride=1, downsample="", training=True, n_s=0, depth_3d=47):
        if n_s >= depth_3d:
            raise NotImplementedError("2-D bottlenecks (n_s >= depth_3d) are dead code in the reference graphs")
        self.x, self.inplanes, self.planes = l_input, inplanes, planes
        self.id = n_s
        self.training = training
        self.first = downsample != ""
        self.hw_stride = 2 if (self.first and n_s != 0) else 1  # p3d.py:45-49
        self.ST = "ABC"[n_s % 3]

    def infer(self) -> T:
        x, eng, tr, pl, i = self.x, self.x.eng, self.training, self.planes, self.id
        s = (1, self.hw_stride, self.hw_stride)
        one = (1, 1, 1)
        c1 = eng.conv([x], pl, one, s, get_conv_weight(eng, f"conv3_{i}_1", [1, 1, 1, self.inplanes, pl]), name=f"conv3_{i}_1")
        o = nw.bn_relu(c1, tr, tap=f"b{i}/1")
        nm = f"ST{self.ST}_{i}_2"
        if self.ST == "A":      # S -> T in series
            o = nw.bn_relu(convS(nm + "_S", o, pl, pl), tr, tap=nm + "_S")
            o = nw.bn_relu(convT(nm + "_T", o, pl, pl), tr, tap=nm + "_T")
        elif self.ST == "B":    # relu(bn(T(o))) + relu(bn(S(o))) in one fused pass
            s_raw = convS(nm + "_S", o, pl, pl)
            s_raw.op.aux = True     # S and T are independent: S runs on the aux stream, the fused norm op joins them
            ns_s = nw._bn_state(eng, pl)
            t_raw = convT(nm + "_T", o, pl, pl)
            ns_t = nw._bn_state(eng, pl)
            o = eng.norm_act(t_raw, ns_t, tr, True, b=s_raw, n2=ns_s, train2=tr, relu2=True, name=nm)
        else:                   # C: s = relu(bn(S(o))); s + relu(bn(T(s)))
            s_act = nw.bn_relu(convS(nm + "_S", o, pl, pl), tr, tap=nm + "_S")
            t_raw = convT(nm + "_T", s_act, pl, pl)
            o = eng.norm_act(t_raw, nw._bn_state(eng, pl), tr, True, b=s_act, name=nm)
        c3 = eng.conv([o], pl * BLOCK_EXPANSION, one, one,
                      get_conv_weight(eng, f"conv3_{i}_3", [1, 1, 1, pl, pl * BLOCK_EXPANSION]), name=f"conv3_{i}_3")
        ns3 = nw._bn_state(eng, pl * BLOCK_EXPANSION)
        if self.first:          # projection shortcut dw3d_{id} (p3d.py:124-127)
            r = eng.conv([x], pl * BLOCK_EXPANSION, one, s,
                         get_conv_weight(eng, f"dw3d_{i}", [1, 1, 1, self.inplanes, pl * BLOCK_EXPANSION]), name=f"dw3d_{i}")
            nsr = nw._bn_state(eng, pl * BLOCK_EXPANSION)
            # (not flagged aux: created after conv3 to keep TF's variable numbering, it would only hop to the aux stream behind c3
            #  and be joined immediately -- three launches per step, no overlap to win)
            y = eng.norm_act(c3, ns3, tr, False, b=r, n2=nsr, train2=tr, relu2=False, relu_out=True, name=f"b{i}")
        else:
            y = eng.norm_act(c3, ns3, tr, False, b=x, relu_out=True, name=f"b{i}")
        return eng.tap(f"b{i}", y)


class make_block:
    """p3d.py:139-166.  NB: no caller passes `training`, so backbone BatchNorm always uses batch statistics."""

    def __init__(self, _X: T, planes, num, inplanes, cnt, training=True, depth_3d=47, stride=1):
        self.input, self.planes, self.num, self.inplanes, self.cnt = _X, planes, num, inplanes, cnt
        self.training, self.depth_3d, self.stride = training, depth_3d, stride

    def infer(self) -> T:
        x = self.input
        for j in range(self.num):
            x = Bottleneck(x, self.inplanes if j == 0 else BLOCK_EXPANSION * self.planes, self.planes, self.stride,
                           downsample="3d" if j == 0 else "", training=self.training, n_s=self.cnt, depth_3d=self.depth_3d).infer()
            self.cnt += 1
        self.inplanes = BLOCK_EXPANSION * self.planes
        return x


def _stem(_X: T, training: bool) -> T:
    """conv 1x7x7 s(1,2,2) 3->64 (no bias) -> BN(training) -> ReLU   (p3d.py:343-345)"""
    eng = _X.eng
    c = eng.conv([_X], 64, (1, 7, 7), (1, 2, 2), get_conv_weight(eng, "firstconv1", [1, 7, 7, 3, 64]), name="firstconv1")
    eng.tap("firstconv1", c.raw)
    return eng.tap("stem", nw.bn_relu(c, training, tap="stem"))


class _ActivationOutput:
    """network output that is an activation tensor (what Session.run returns for the stem-only graph)"""

    def __init__(self, t: T):
        self.eng, self.t = t.eng, t

    @property
    def output(self):
        return self.t.buf


def p3d_stem(_X, _dropout=0.0, batch_size=2, training=False):
    """the frame-local head of every p3d.py graph on its own: conv 1x7x7 s(1,2,2) -> BN(training) -> ReLU (p3d.py:343-345).
    With training=False (moving statistics) its output for one frame does not depend on the other frames of the clip, which is
    what lets the sliding-window inference of gen_pred.py:88-135 compute it ONCE per frame (video.predict_video_cached).
    _X: [F, D, H, W, 3] frames (D = 1 for single frames); returns a handle whose Session.run gives [F, D, H/2, W/2, 64]."""
    if training:
        raise NotImplementedError("p3d_stem is the inference-time frame cache (batch statistics would couple the frames)")
    return _ActivationOutput(_stem(_X, False))


def _backbone(_X: T, training: bool, skip_1_0: bool = True, forks=None):
    """forks: {feature name: n} -> the returned dict holds a LIST of n cross-branch aliases (Engine.fork) for that feature instead
    of the tensor, one per consumer branch; recorded right behind the feature's producer so that, in the backward pass, the main
    chain waits for the consumer branches exactly where it is about to need the feature's gradient"""
    eng = _X.eng
    forks = forks or {}

    def out(name, x):
        n = forks.get(name, 0)
        return [eng.fork(x, f"{name}/fork{i}") for i in range(n)] if n else x
    if _X.C == 64:
        # the input already IS the stem activation (frame cache of the sliding-window inference): keep TF's variable numbering
        # by burning the names the stem's layers would have taken (firstconv1 is named explicitly; its BN is the first
        # auto-numbered batch_normalization)
        if training:
            raise NotImplementedError("stem activations as graph input are an inference-only path")
        eng.names.unique("", "batch_normalization")
        stem = eng.tap("stem", _X)
    else:
        stem = _stem(_X, training)
    t = {}
    if skip_1_0:
        t["x_1_0"] = out("x_1_0", eng.tap("x_1_0", eng.maxpool(stem, *TEMPORAL_POOL, name="x_1_0")))
    x = eng.tap("pool1", eng.maxpool(stem, (2, 3, 3), (2, 2, 2), name="pool1"))
    cnt = 0
    for si, (planes, num, inplanes, stride) in enumerate(STAGES):
        if si >= 1:                # data-parallel overlap: backward is cut before every stage but the first
            eng.mark_dp_split()
        blk = make_block(x, planes, num, inplanes, cnt, stride=stride)
        res = blk.infer()
        cnt = blk.cnt
        x = eng.tap(f"x_{si + 2}_0", eng.maxpool(res, *TEMPORAL_POOL, name=f"x_{si + 2}_0"))
        t[f"x_{si + 2}_0"] = out(f"x_{si + 2}_0", x)
    return t


def _unetpp(_X: T, _dropout: float, training: bool, sa: bool):
    """The decoder runs on two branches beside the backbone (Engine.branch; SAP3D_BRANCHES=0 = one stream, same ops, same order):
      branch 2 (early): upx_3_0, x_2_1, upx_2_0, x_1_1 -- they need only x_1_0 .. x_3_0, so their forward runs beside the 36-block
                        stage of the backbone and their backward beside the rest of the decoder's, off the critical path;
      branch 1 (late):  everything that needs x_4_0 (x_4_0_sa ... x_1_3, head).
    The creation order of the layers -- TF's variable numbering -- is the reference's."""
    eng = _X.eng
    br = eng.branches_enabled
    t = _backbone(_X, training, forks={"x_1_0": 1, "x_2_0": 1, "x_3_0": 2, "x_4_0": 1} if br else None)
    if br:
        (x10,), (x20,), (x30, x30e), (x40,) = t["x_1_0"], t["x_2_0"], t["x_3_0"], t["x_4_0"]
    else:
        x10, x20, x30, x40 = t["x_1_0"], t["x_2_0"], t["x_3_0"], t["x_4_0"]
        x30e = x30
    att = (lambda h, name, **kw: eng.tap(name, nw.attention(h, name, training=training, **kw))) if sa else (lambda h, name, **kw: h)
    with eng.branch(1):
        eng.consume_forked([x40, x30])
        x40 = att(x40, "x_4_0_sa")
        up40 = nw.transpose_conv3d(x40, 512, [1, 3, 3], [2, 2, 2], training, "upx_4_0")
        x31 = nw.conv3d(nw.concat([x30, up40]), 512, [2, 3, 3], [1, 1, 1], training, "x_3_1")
        x31 = att(x31, "x_3_1_sa")
    with eng.branch(2):
        eng.consume_forked([x30e, x20, x10])
        up30 = nw.transpose_conv3d(x30e, 256, [2, 3, 3], [2, 2, 2], training, "upx_3_0")
        x21 = nw.conv3d(nw.concat([x20, up30]), 256, [3, 3, 3], [1, 1, 1], training, "x_2_1")
        x21l = eng.fork(x21, "x_2_1/fork")          # for the late branch (upx_2_1, x_2_2)
    with eng.branch(1):
        up31 = nw.transpose_conv3d(x31, 256, [2, 3, 3], [2, 2, 2], training, "upx_3_1")
        eng.consume_forked([x21l])
        x22 = nw.conv3d(nw.concat([x21l, up31]), 256, [3, 3, 3], [1, 1, 1], training, "x_2_2")
        x22 = att(x22, "x_2_2_sa")
    with eng.branch(2):
        up20 = nw.transpose_conv3d(x20, 128, [3, 3, 3], [2, 2, 2], training, "upx_2_0")
        x11 = nw.conv3d(nw.concat([x10, up20]), 128, [3, 3, 3], [1, 1, 1], training, "x_1_1")
        x11l = eng.fork(x11, "x_1_1/fork")
    with eng.branch(1):
        up21 = nw.transpose_conv3d(x21l, 128, [3, 3, 3], [2, 2, 2], training, "upx_2_1")
        eng.consume_forked([x11l])
        x12 = nw.conv3d(nw.concat([x11l, up21]), 128, [3, 3, 3], [1, 1, 1], training, "x_1_2")
        up22 = nw.transpose_conv3d(x22, 128, [3, 3, 3], [2, 2, 2], training, "upx_2_2")
        x13 = nw.conv3d(nw.concat([x12, up22]), 128, [3, 3, 3], [1, 1, 1], training, "x_1_3")
        x13 = att(x13, "x_1_3_sa", subsample=True)
        if training:
            x13 = eng.dropout(x13, _dropout, name="x_1_3_drop")
        w = eng.param("x_0_1/kernel", [3, 3, 3, 1, x13.C], "glorot_t")
        b = eng.param("x_0_1/bias", [1], "zeros")
        return eng.head(x13, w, b, (3, 3, 3), 2, sigmoid=True, name="x_0_1")


def p3d_unetplusplus_ds(_X, _dropout, batch_size=2, training=True, SA=False):
    """p3d.py:340-399 — UNet++ decoder with self-attention at x_4_0, x_3_1, x_2_2, x_1_3 (what gen_pred.py:46 builds)."""
    return _unetpp(_X, _dropout, training, True)


def p3d_unetplusplus_nonsa(_X, _dropout, batch_size=2, training=True, SA=False):
    """p3d.py:401-459 — the same decoder without the attention blocks."""
    return _unetpp(_X, _dropout, training, False)


def p3d_unetplusplus(_X, _dropout, batch_size=2, training=True, SA=False):
    """p3d.py:280-338 is not buildable in the reference (shape error at p3d.py:334)."""
    raise NotImplementedError("p3d_unetplusplus has a shape mismatch at p3d.py:334; use p3d_unetplusplus_ds")


def p3d_unet(_X, _dropout, batch_size=2, training=True):
    """p3d.py:169-221"""
    eng = _X.eng
    t = _backbone(_X, training, skip_1_0=False)
    d1 = nw.bn_relu(nw.layers_conv3d_transpose(t["x_4_0"], 512, [1, 3, 3], [2, 2, 2]), training, name="deconv1_bn")
    d2 = nw.bn_relu(nw.layers_conv3d_transpose(nw.concat([d1, t["x_3_0"]]), 256, [2, 3, 3], [2, 2, 2]), training, name="deconv2_bn")
    d3 = nw.bn_relu(nw.layers_conv3d_transpose(nw.concat([d2, t["x_2_0"]]), 128, 3, [2, 2, 2]), training, name="deconv3_bn")
    if training:
        d3 = eng.dropout(d3, _dropout, name="deconv3_drop")
    eng.tap("deconv3", d3)
    c = nw.layers_conv3d(d3, 32, 1, 1, want_stats=False)
    w = eng.param(eng.names.unique("", "conv3d_transpose") + "/kernel", [3, 3, 3, 1, 32], "glorot_t")
    b = eng.param(w.name.rsplit("/", 1)[0] + "/bias", [1], "zeros")
    return eng.head(c.raw, w, b, (3, 3, 3), 2, sigmoid=True, name="results")


def p3d_concat(_X, _dropout, batch_size=2, training=True):
    """p3d.py:224-276: each stage's pooled output is upsampled to 4x28x28 (deconv 3^3 with stride 1 / 2 / 4), the three
    maps are concatenated (128+256+512) -> conv_concat -> deconv_revise -> predict_revise.  Returns LOGITS (no sigmoid
    at p3d.py:275-276)."""
    eng = _X.eng
    x = eng.tap("pool1", eng.maxpool(_stem(_X, training), (2, 3, 3), (2, 2, 2), name="pool1"))
    side = ((1, 128), (2, 256), (4, 512))   # (stride, filters) of deconv_pool{2,3,4}
    ups, cnt = [], 0
    for si, (planes, num, inplanes, stride) in enumerate(STAGES):
        if si >= 1:                # data-parallel overlap: backward is cut before every stage but the first
            eng.mark_dp_split()
        blk = make_block(x, planes, num, inplanes, cnt, stride=stride)
        res = blk.infer()
        cnt = blk.cnt
        x = eng.tap(f"pool{si + 2}", eng.maxpool(res, *TEMPORAL_POOL, name=f"pool{si + 2}"))
        s, f = side[si]
        nm = f"deconv_pool{si + 2}"
        co = nw.layers_conv3d_transpose(x, f, 3, [s, s, s], nm, bias_grad=not training)
        ups.append(nw.bn_relu(co, training, name=nm + "_bn", tap=nm))
    cat = ConcatOp(eng, ups[0], ups[1], name="concatenator_01").y
    cc = nw.bn_relu(nw.layers_conv3d(nw.concat([cat, ups[2]]), 512, 3, 1, "conv_concat", bias_grad=not training), training,
                    name="conv_concat_bn", tap="conv_concat")
    eng.tap("conv_concat", cc)
    dr = nw.bn_relu(nw.layers_conv3d_transpose(cc, 128, 3, 2, "deconv_revise", bias_grad=not training), training,
                    name="deconv1_revise_bn", tap="deconv_revise")
    if training:
        dr = eng.dropout(dr, _dropout, name="deconv_revise_drop")
    w = eng.param("predict_revise/kernel", [3, 3, 3, 1, 128], "glorot_t")
    b = eng.param("predict_revise/bias", [1], "zeros")
    return eng.head(dr, w, b, (3, 3, 3), 2, sigmoid=False, name="predict_revise")
