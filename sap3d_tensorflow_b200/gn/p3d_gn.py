"""GroupNorm + CBAM variant of the P3D saliency model on the B200 engine — public surface of the reference's
gn/p3d_gn.py: GroupNorm :24, GNReLU :49, conv3d_layers :14, deconv3d_layers :19, Bottleneck :74, make_block :182,
inference_p3d :214, inference_p3d_concat :279, inference_p3d_decoder_block :489.  Every block ends with cbam_block on the residual
(gn/p3d_gn.py:175).  Variables: group_norm[_N]/{gamma,beta}, cbam_{id}/ch_at/mlp_{0,1}/{kernel,bias},
cbam_{id}/sp_at/conv3d/kernel, conv names as in p3d.py.

The *_sa_* builders of the reference call attention() with a stale signature (gn/p3d_gn.py:340 vs
utils/network.py:157) and cannot be built; they are not mirrored."""
from __future__ import annotations

from .. import network as nw
from ..engine import ConvOut, T
from ..engine_gn import CbamBlockTailOp, ConcatOp, GNActOp, GNState
from ..p3d import BLOCK_EXPANSION, STAGES, TEMPORAL_POOL, get_conv_weight
from ..p3d import convS as _convS_bn, convT as _convT_bn


def convS(name, l_input, in_channels, out_channels):   # gn/p3d_gn.py:62-65 (a GroupNorm does not cancel the bias gradient)
    return _convS_bn(name, l_input, in_channels, out_channels, bias_grad=True)


def convT(name, l_input, in_channels, out_channels):   # gn/p3d_gn.py:67-70
    return _convT_bn(name, l_input, in_channels, out_channels, bias_grad=True)


_gn_state = nw.gn_state


def GroupNorm(x: ConvOut, G=32, esp=1e-5, relu=False, name="") -> T:   # gn/p3d_gn.py:24-46
    return GNActOp(x.raw.eng, x, _gn_state(x.raw), relu, name=name).y


def GNReLU(x: ConvOut, name=None) -> T:                                 # gn/p3d_gn.py:49-51
    return GroupNorm(x, relu=True, name=name or "gnrelu")


def conv3d_layers(x, filters, kernel, strides, name) -> T:              # gn/p3d_gn.py:14-17
    return GNReLU(nw.layers_conv3d(x, filters, kernel, strides, name, want_stats=False), name)


def deconv3d_layers(x, filters, kernel, strides, name) -> T:            # gn/p3d_gn.py:19-22
    return GNReLU(nw.layers_conv3d_transpose(x, filters, kernel, strides, name, want_stats=False), name)


class Bottleneck:
    """gn/p3d_gn.py:74-179 (3-D branch)"""

    def __init__(self, l_input: T, inplanes, planes, stride=1, downsample="", training=True, n_s=0, depth_3d=47):
        if n_s >= depth_3d:
            raise NotImplementedError("2-D bottlenecks are dead code in the reference graphs")
        self.x, self.inplanes, self.planes, self.id = l_input, inplanes, planes, n_s
        self.first = downsample != ""
        self.hw_stride = 2 if (self.first and n_s != 0) else 1
        self.ST = "ABC"[n_s % 3]

    def infer(self) -> T:
        x, eng, pl, i = self.x, self.x.eng, self.planes, self.id
        s = (1, self.hw_stride, self.hw_stride)
        one = (1, 1, 1)
        conv = lambda t, cout, st, nm: eng.conv([t], cout, one, st, get_conv_weight(eng, nm, [1, 1, 1, t.C, cout]), name=nm, want_stats=False)  # noqa: E731
        o = GroupNorm(conv(x, pl, s, f"conv3_{i}_1"), relu=True, name=f"b{i}/1")
        nm = f"ST{self.ST}_{i}_2"
        if self.ST == "A":
            o = GroupNorm(convS(nm + "_S", o, pl, pl), relu=True, name=nm + "_S")
            o = GroupNorm(convT(nm + "_T", o, pl, pl), relu=True, name=nm + "_T")
        elif self.ST == "B":
            s_raw = convS(nm + "_S", o, pl, pl)
            gs = _gn_state(s_raw.raw)
            t_raw = convT(nm + "_T", o, pl, pl)
            gt = _gn_state(t_raw.raw)
            o = GNActOp(eng, t_raw, gt, True, b=s_raw, g2=gs, relu2=True, name=nm).y
        else:
            s_act = GroupNorm(convS(nm + "_S", o, pl, pl), relu=True, name=nm + "_S")
            t_raw = convT(nm + "_T", s_act, pl, pl)
            o = GNActOp(eng, t_raw, _gn_state(t_raw.raw), True, b=s_act, name=nm).y
        c3 = conv(o, pl * BLOCK_EXPANSION, one, f"conv3_{i}_3")
        g3 = _gn_state(c3.raw)
        residual = x
        if self.first:
            residual = GroupNorm(conv(x, pl * BLOCK_EXPANSION, s, f"dw3d_{i}"), relu=False, name=f"dw3d_{i}")
        Cc = pl * BLOCK_EXPANSION
        sc = f"cbam_{i}"
        tail = CbamBlockTailOp(eng, c3, g3, residual,
                               eng.param(sc + "/ch_at/mlp_0/kernel", [Cc, Cc // 8], "vscale"), eng.param(sc + "/ch_at/mlp_0/bias", [Cc // 8], "zeros"),
                               eng.param(sc + "/ch_at/mlp_1/kernel", [Cc // 8, Cc], "vscale"), eng.param(sc + "/ch_at/mlp_1/bias", [Cc], "zeros"),
                               eng.param(sc + "/sp_at/conv3d/kernel", [7, 7, 7, 2, 1], "vscale"), name=f"b{i}")
        return eng.tap(f"b{i}", tail.y)


class make_block:
    """gn/p3d_gn.py:182-209"""

    def __init__(self, _X: T, planes, num, inplanes, cnt, training=True, depth_3d=47, stride=1):
        self.input, self.planes, self.num, self.inplanes, self.cnt, self.stride, self.depth_3d = _X, planes, num, inplanes, cnt, stride, depth_3d

    def infer(self) -> T:
        x = self.input
        for j in range(self.num):
            x = Bottleneck(x, self.inplanes if j == 0 else BLOCK_EXPANSION * self.planes, self.planes, self.stride,
                           downsample="3d" if j == 0 else "", n_s=self.cnt, depth_3d=self.depth_3d).infer()
            self.cnt += 1
        return x


def _inference(_X: T, _dropout, training, pool4_filters):
    eng = _X.eng
    c = eng.conv([_X], 64, (1, 7, 7), (1, 2, 2), get_conv_weight(eng, "firstconv1", [1, 7, 7, 3, 64]), name="firstconv1", want_stats=False)
    x = eng.maxpool(GroupNorm(c, relu=True, name="stem"), (2, 3, 3), (2, 2, 2), name="pool1")
    pools, cnt, dp3 = [], 0, None
    for si, (planes, num, inplanes, stride) in enumerate(STAGES):
        if si >= 1:                # data-parallel overlap: backward is cut before every stage but the first
            eng.mark_dp_split()
        blk = make_block(x, planes, num, inplanes, cnt, stride=stride)
        res = blk.infer()
        cnt = blk.cnt
        x = eng.tap(f"pool{si + 2}", eng.maxpool(res, *TEMPORAL_POOL, name=f"pool{si + 2}"))
        pools.append(x)
        if si == 1:  # the reference creates deconv_pool3 (and its GroupNorm variables) BEFORE stage 3 (gn/p3d_gn.py:234-237)
            dp3 = GNReLU(nw.layers_conv3d_transpose(x, 512, 3, [2, 2, 2], "deconv_pool3", want_stats=False), "deconv_pool3_gn")
    pool2, pool3, pool4 = pools  # [4,28,28,256], [2,14,14,512], [1,7,7,1024]
    dp4 = GNReLU(nw.layers_conv3d_transpose(pool4, pool4_filters, 3, [4, 4, 4], "deconv_pool4", want_stats=False), "deconv_pool4_gn")
    cat = ConcatOp(eng, dp3, dp4, name="concat_dp3_dp4").y
    cc = GNReLU(nw.layers_conv3d(nw.concat([cat, pool2]), 1024, 3, 1, "conv_concat", want_stats=False), "conv_concat")
    eng.tap("conv_concat", cc)
    dr = GNReLU(nw.layers_conv3d_transpose(cc, 256, 3, 2, "deconv_revise", want_stats=False), "deconv_revise")
    if training:
        dr = eng.dropout(dr, _dropout, name="deconv_revise_drop")
    w = eng.param("predict_revise/kernel", [3, 3, 3, 1, 256], "glorot_t")
    b = eng.param("predict_revise/bias", [1], "zeros")
    return eng.head(dr, w, b, (3, 3, 3), 2, sigmoid=False, name="predict_revise")  # returns logits (gn/p3d_gn.py:257-258)


def inference_p3d(_X, _dropout, batch_size=2, training=True):
    """gn/p3d_gn.py:214-258"""
    return _inference(_X, _dropout, training, 1024)


def inference_p3d_concat(_X, _dropout, batch_size=2, training=True):
    """gn/p3d_gn.py:279-324"""
    return _inference(_X, _dropout, training, 512)


def inference_p3d_decoder_block(_X, _dropout, batch_size=2, training=True):
    """gn/p3d_gn.py:489-539: three-scale concat -> conv_concat -> two decoder blocks -> 3x3x3 conv to 1 channel (logits).
    The whole builder sits inside tf.variable_scope('P3D'), so every variable name carries the 'P3D/' prefix."""
    eng = _X.eng
    eng.var_prefix = "P3D/"
    c = eng.conv([_X], 64, (1, 7, 7), (1, 2, 2), get_conv_weight(eng, "firstconv1", [1, 7, 7, 3, 64]), name="firstconv1", want_stats=False)
    x = eng.maxpool(GroupNorm(c, relu=True, name="stem"), (2, 3, 3), (2, 2, 2), name="pool1")
    # (kernel, stride, filters) of the deconv applied to each stage's pooled output, created right after that stage
    side = (((3, 3, 3), (1, 1, 1), 128), ((2, 3, 3), (2, 2, 2), 256), ((1, 3, 3), (4, 4, 4), 512))
    ups, cnt = [], 0
    for si, (planes, num, inplanes, stride) in enumerate(STAGES):
        if si >= 1:                # data-parallel overlap: backward is cut before every stage but the first
            eng.mark_dp_split()
        blk = make_block(x, planes, num, inplanes, cnt, stride=stride)
        res = blk.infer()
        cnt = blk.cnt
        x = eng.tap(f"pool{si + 2}", eng.maxpool(res, *TEMPORAL_POOL, name=f"pool{si + 2}"))
        k, s, f = side[si]
        ups.append(deconv3d_layers(x, f, k, s, f"deconv_pool{si + 2}"))
    cat = ConcatOp(eng, ups[0], ups[1], name="concat_dp2_dp3").y
    cc = eng.tap("conv_concat", conv3d_layers(nw.concat([cat, ups[2]]), 1024, 3, 1, "conv_concat"))
    d = conv3d_layers(cc, 256, 3, 1, "decoder1_conv1")
    d = deconv3d_layers(d, 256, 3, 2, "decoder1_deconv")
    d = conv3d_layers(d, 128, 3, 1, "decoder1_conv2")
    d = conv3d_layers(d, 32, 3, 1, "decoder2_conv1")
    d = deconv3d_layers(d, 32, 3, 2, "decoder2_deconv")
    d = eng.tap("decoder2_conv2", conv3d_layers(d, 16, 3, 1, "decoder2_conv2"))
    if training:
        d = eng.dropout(d, _dropout, name="final_drop")
    w = eng.param("results/kernel", [3, 3, 3, 16, 1], "glorot")
    b = eng.param("results/bias", [1], "zeros")
    logits = eng.conv([d], 1, (3, 3, 3), (1, 1, 1), w, b, want_stats=False, name="results", out_f32=True)
    return eng.logits_loss(logits, name="results")
