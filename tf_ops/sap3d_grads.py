"""Gradient registrations and thin Python wrappers for the custom ops of tf_ops/sap3d_tf_ops.cc (TensorFlow 1.x /
tf.compat.v1 graph mode).  This is what a maintainer of the reference imports from utils/network.py / p3d.py instead of
calling tf.nn.conv3d, tf.layers.batch_normalization, ... directly (INTEGRATION.md section A shows the edited call sites).

STATUS: TensorFlow is not installable in the build image, so this module has never been imported; CI parses it and checks
that every op it references is registered in sap3d_tf_ops.cc with the inputs / outputs / attributes used here
(tests/test_tf_shim_cpu.py).

tf.gradients (AdamOptimizer.minimize, train.py:166-168) walks these registrations exactly as it walks the stock ones.
"""
import os

import tensorflow as tf
from tensorflow.python.framework import ops

_HERE = os.path.dirname(os.path.abspath(__file__))
_sap3d = tf.load_op_library(os.path.join(_HERE, "libsap3d_tf_ops.so"))

CONV_ATTRS = ("input_dims", "cin", "cout", "ksize", "strides", "transposed", "has_bias", "out_f32", "storage")


def _conv_attrs(op):
    return {k: op.get_attr(k) for k in CONV_ATTRS}


# ---- convolution family ------------------------------------------------------------------------------------------------
def conv3d(x, filt, bias=None, strides=(1, 1, 1), transposed=False, cout=None, out_f32=False, storage="bf16"):
    """tf.nn.conv3d(+bias_add) / tf.layers.conv3d / tf.layers.conv3d_transpose with TF 'SAME' / 'same' geometry.
    x: one tensor or a list of two (a tf.concat on the channel axis that is never materialised).  Returns (y, stats):
    stats feeds batch_norm() below (the per-tile sum / sum of squares the conv epilogue produced)."""
    xs = list(x) if isinstance(x, (list, tuple)) else [x]
    n, d, h, w = [int(s) for s in xs[0].shape[:4]]
    ksize = [int(s) for s in filt.shape[:3]]
    cin = [int(t.shape[4]) for t in xs]
    cout = int(cout if cout is not None else (filt.shape[3] if transposed else filt.shape[4]))
    attrs = dict(input_dims=[n, d, h, w], cin=cin, cout=cout, ksize=ksize, strides=list(strides), transposed=transposed,
                 has_bias=bias is not None, out_f32=out_f32, storage=storage)
    packed, _ = _sap3d.sap3d_pack_filter(filt, **attrs)
    b = bias if bias is not None else tf.zeros([cout], tf.float32)
    y, stats, _ = _sap3d.sap3d_conv(xs[0], xs[-1], filt, packed, b, Tout=tf.float32 if (out_f32 or storage == "f32") else tf.bfloat16, **attrs)
    return y, stats


@ops.RegisterGradient("Sap3dConv")
def _conv_grad(op, dy, _dstats, _doperand):
    x0, x1, filt, _packed, _bias = op.inputs
    a = _conv_attrs(op)
    _, packed_dgrad = _sap3d.sap3d_pack_filter(filt, **a)
    dx0 = _sap3d.sap3d_conv_grad_input(dy, filt, packed_dgrad, seg=0, **a)
    dx1 = _sap3d.sap3d_conv_grad_input(dy, filt, packed_dgrad, seg=1, **a) if len(a["cin"]) > 1 else None
    dw, db = _sap3d.sap3d_conv_grad_filter(x0, x1, dy, op.outputs[2], **a)
    return dx0, dx1, dw, None, (db if a["has_bias"] else None)


ops.NotDifferentiable("Sap3dPackFilter")
ops.NotDifferentiable("Sap3dConvAffine")        # inference graphs only
ops.NotDifferentiable("Sap3dConvGradInput")
ops.NotDifferentiable("Sap3dConvGradFilter")


# ---- BatchNorm / GroupNorm + ReLU + residual add -----------------------------------------------------------------------
def batch_norm(y, stats, gamma, beta, moving_mean, moving_variance, training, relu=True, residual=None, relu_out=False):
    """tf.layers.batch_normalization(y, training=training) [+ tf.nn.relu] [+ residual, relu] (p3d.py:56-81,133-134).
    The moving-average updates are added to tf.GraphKeys.UPDATE_OPS like the stock layer's (train.py:170-172)."""
    count = 1.0
    for s in y.shape[:4]:
        count *= int(s)
    scale, shift, mean, rstd, new_mm, new_mv = _sap3d.sap3d_bn_finalize(stats, gamma, beta, moving_mean, moving_variance, count=count,
                                                                         training=training)
    if training:
        tf.add_to_collection(tf.GraphKeys.UPDATE_OPS, tf.assign(moving_mean, new_mm))
        tf.add_to_collection(tf.GraphKeys.UPDATE_OPS, tf.assign(moving_variance, new_mv))
    b = residual if residual is not None else y
    out = _sap3d.sap3d_norm_apply(y, scale, shift, b, scale, shift, relu1=relu, relu2=False, relu_out=relu_out,
                                  has_b=residual is not None, norm_b=False)
    # the gradient op needs mean / rstd and the affine parameters: keep them reachable from the forward op
    out.op._sap3d_bn = (mean, rstd, gamma, beta, training)
    return out


@ops.RegisterGradient("Sap3dNormApply")
def _norm_apply_grad(op, dy):
    a, s1, t1, b, s2, t2 = op.inputs
    mean, rstd, _gamma, _beta, training = op._sap3d_bn
    da, db, dgamma1, dbeta1, _dg2, _db2 = _sap3d.sap3d_norm_apply_grad(
        dy, a, s1, t1, mean, rstd, b, s2, t2, mean, rstd, relu1=op.get_attr("relu1"), relu2=op.get_attr("relu2"),
        relu_out=op.get_attr("relu_out"), has_b=op.get_attr("has_b"), norm_b=op.get_attr("norm_b"), batch_stats1=training,
        batch_stats2=training)
    # scale = gamma * rstd and shift = beta - mean * scale are functions of (gamma, beta) through Sap3dBnFinalize: with the
    # batch statistics already differentiated inside the fused backward, d scale = dgamma / rstd and d shift = dbeta
    return da, dgamma1 / rstd, dbeta1, (db if op.get_attr("has_b") else None), None, None


@ops.RegisterGradient("Sap3dBnFinalize")
def _bn_finalize_grad(op, dscale, dshift, *_unused):
    _stats, gamma, _beta, _mm, _mv = op.inputs
    rstd = op.outputs[3]
    mean = op.outputs[2]
    # (statistics carry no gradient here: Sap3dNormApplyGrad already contains the full BatchNorm backward)
    return None, dscale * rstd - dshift * mean * rstd, dshift, None, None


def group_norm(x, gamma, beta, relu=False, residual=None, relu_out=False):
    """GroupNorm (utils/network.py:65-87) [+ ReLU] [+ residual, relu]"""
    scale, shift, mean, rstd = _sap3d.sap3d_group_norm_stats(x, gamma, beta)
    pps = 1
    for s in x.shape[1:4]:
        pps *= int(s)
    b = residual if residual is not None else x
    out = _sap3d.sap3d_norm_apply(x, scale, shift, b, scale, shift, relu1=relu, relu2=False, relu_out=relu_out, has_b=residual is not None,
                                  norm_b=False, positions_per_sample=pps)
    out.op._sap3d_gn = (mean, rstd, gamma, beta)
    return out


def _group_norm_apply_grad(op, dy):
    a, s1, t1, b, s2, t2 = op.inputs
    mean, rstd, gamma, _beta = op._sap3d_gn
    da, db, dgamma, dbeta, _g2, _b2 = _sap3d.sap3d_group_norm_grad(dy, a, s1, t1, mean, rstd, gamma, b, s2, t2, mean, rstd, gamma,
                                                                   relu1=op.get_attr("relu1"), relu2=op.get_attr("relu2"),
                                                                   relu_out=op.get_attr("relu_out"), has_b=op.get_attr("has_b"),
                                                                   norm_b=op.get_attr("norm_b"))
    return da, dgamma, dbeta, (db if op.get_attr("has_b") else None)


ops.NotDifferentiable("Sap3dGroupNormStats")     # its gradient is inside Sap3dGroupNormGrad (dgamma / dbeta returned there)
ops.NotDifferentiable("Sap3dNormApplyGrad")
ops.NotDifferentiable("Sap3dGroupNormGrad")


# ---- CBAM block tail (gn/p3d_gn.py:175-177) ----------------------------------------------------------------------------
def cbam_block_tail(c3, gamma3, beta3, residual, w0, b0, w1, b1, w_sp):
    """relu(GroupNorm(c3) + cbam_block(residual))"""
    scale3, shift3, mean3, rstd3 = _sap3d.sap3d_group_norm_stats(c3, gamma3, beta3)
    y, cscale, sp, att, save = _sap3d.sap3d_cbam_tail(c3, scale3, shift3, residual, w0, b0, w1, b1, w_sp)
    y.op._sap3d_cbam = (mean3, rstd3, gamma3)
    return y


@ops.RegisterGradient("Sap3dCbamTail")
def _cbam_tail_grad(op, dy, *_unused):
    c3, scale3, _shift3, r, w0, _b0, w1, _b1, w_sp = op.inputs
    y, cscale, sp, att, save = op.outputs
    mean3, rstd3, gamma3 = op._sap3d_cbam
    dc3, dr, dgamma3, dbeta3, dw0, db0, dw1, db1, dw_sp = _sap3d.sap3d_cbam_tail_grad(dy, y, c3, scale3, mean3, rstd3, gamma3, r, w0, w1, w_sp,
                                                                                    cscale, sp, att, save)
    # d scale3 / d shift3 route the GroupNorm affine gradient back to gamma3 / beta3 through Sap3dGroupNormStats' outputs
    tf.add_to_collection("sap3d_param_grads", (gamma3, dgamma3))
    tf.add_to_collection("sap3d_param_grads", (op.inputs[2], dbeta3))
    return dc3, None, None, dr, dw0, db0, dw1, db1, dw_sp


ops.NotDifferentiable("Sap3dCbamTailGrad")


# ---- pooling, attention, gate, head, loss, dropout, concat -------------------------------------------------------------
def max_pool3d(x, ksize, strides, padding="SAME"):
    """tf.nn.max_pool3d(x, [1,kd,kh,kw,1], [1,sd,sh,sw,1], padding) (p3d.py:347-348,354,360,366)"""
    y, _ = _sap3d.sap3d_max_pool3d(x, ksize=list(ksize), strides=list(strides), same=(padding == "SAME"))
    return y


@ops.RegisterGradient("Sap3dMaxPool3d")
def _max_pool3d_grad(op, dy, _dargmax):
    return _sap3d.sap3d_max_pool3d_grad(op.inputs[0], dy, op.outputs[1], ksize=op.get_attr("ksize"), strides=op.get_attr("strides"),
                                        same=op.get_attr("same"))


ops.NotDifferentiable("Sap3dMaxPool3dGrad")


def attention_core(g, f, h):
    """softmax(g f^T) h of utils/network.py:184-186 on [B, N, d] tensors; bf16 with d_k == 64 takes the fused kernel"""
    if g.dtype == tf.bfloat16 and int(g.shape[2]) == 64 and int(h.shape[2]) in (128, 256):
        o, _ = _sap3d.sap3d_flash_attention(g, f, h)
        return o
    o, _ = _sap3d.sap3d_attention(g, f, h)
    return o


@ops.RegisterGradient("Sap3dFlashAttention")
def _flash_attention_grad(op, d_o, _dlse):
    q, k, v = op.inputs
    return _sap3d.sap3d_flash_attention_grad(q, k, v, op.outputs[0], d_o, op.outputs[1])


@ops.RegisterGradient("Sap3dAttention")
def _attention_grad(op, d_o, _dbeta):
    g, f, h = op.inputs
    return _sap3d.sap3d_attention_grad(g, f, h, op.outputs[1], d_o)


ops.NotDifferentiable("Sap3dFlashAttentionGrad")
ops.NotDifferentiable("Sap3dAttentionGrad")


@ops.RegisterGradient("Sap3dGate")
def _gate_grad(op, dy):
    d_o, dx, dgamma = _sap3d.sap3d_gate_grad(dy, op.inputs[0], op.inputs[2])
    return d_o, dx, dgamma


ops.NotDifferentiable("Sap3dGateGrad")


@ops.RegisterGradient("Sap3dHead")
def _head_grad(op, dlogits, dpred):
    x, filt, _bias = op.inputs
    pred = op.outputs[1]
    dl = dlogits if dpred is None else (dlogits if dlogits is not None else 0.0) + dpred * pred * (1.0 - pred)
    dx, dw = _sap3d.sap3d_head_grad(dl, x, filt, ksize=op.get_attr("ksize"), stride=op.get_attr("stride"))
    return dx, dw, tf.reshape(tf.reduce_sum(dl), [1])


ops.NotDifferentiable("Sap3dHeadGrad")


def smooth_l1_loss(logits, target, inside_weight=1.0, outside_weight=1.0, sigma=1.0, apply_sigmoid=True):
    """utils/network.py:49-62 on the head's logits (the sigmoid of p3d.py:397 folded in when apply_sigmoid)"""
    loss, _, _ = _sap3d.sap3d_smooth_l1_loss(logits, target, apply_sigmoid=apply_sigmoid, sigma=sigma, inside_weight=inside_weight,
                                             outside_weight=outside_weight)
    return loss


@ops.RegisterGradient("Sap3dSmoothL1Loss")
def _smooth_l1_loss_grad(op, dloss, *_unused):
    return tf.cast(dloss, tf.float32) * op.outputs[1], None


@ops.RegisterGradient("Sap3dDropout")
def _dropout_grad(op, dy):
    return _sap3d.sap3d_dropout(dy, op.inputs[1], rate=op.get_attr("rate"), seed=op.get_attr("seed")), None


@ops.RegisterGradient("Sap3dConcatChannels")
def _concat_channels_grad(op, dy):
    return _sap3d.sap3d_split_channels(dy, ca=int(op.inputs[0].shape[4]), cb=int(op.inputs[1].shape[4]))


ops.NotDifferentiable("Sap3dSplitChannels")
ops.NotDifferentiable("Sap3dAdam")
ops.NotDifferentiable("Sap3dSaliencyMetrics")
ops.NotDifferentiable("Sap3dResizeBilinear")
ops.NotDifferentiable("Sap3dSaliencyAuc")
ops.NotDifferentiable("Sap3dPreprocessFrames")


# the NormApply gradient depends on which statistics produced its scale / shift
_bn_grad = _norm_apply_grad


def _dispatch_norm_apply_grad(op, dy):
    return _group_norm_apply_grad(op, dy) + (None, None) if hasattr(op, "_sap3d_gn") else _bn_grad(op, dy)
