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


# ---- BatchNorm / GroupNorm + ReLU + residual add (one op per normalisation layer) ---------------------------------------
def batch_norm_act(conv_out, gamma, beta, moving_mean, moving_variance, training, relu=True, other=None, other_norm=None,
                   relu_other=False, relu_out=False):
    """tf.layers.batch_normalization(y, training=training) [+ tf.nn.relu] [+ other, relu] (p3d.py:56-81,133-134).
    conv_out = (y, stats) from conv3d().  other: a plain tensor (identity shortcut, ST_C) or a second (y, stats) pair with its
    own variables other_norm = (gamma, beta, moving_mean, moving_variance) (ST_B, projection shortcut).  The moving-average
    updates are added to tf.GraphKeys.UPDATE_OPS like the stock layer's (train.py:170-172)."""
    y, stats = conv_out
    has_b, norm_b = other is not None, other_norm is not None
    if norm_b:
        b, stats_b = other
        g2, b2, mm2, mv2 = other_norm
    else:
        b, stats_b = (other if has_b else y), stats
        g2, b2, mm2, mv2 = gamma, beta, moving_mean, moving_variance
    outs = _sap3d.sap3d_batch_norm_act(y, stats, gamma, beta, moving_mean, moving_variance, b, stats_b, g2, b2, mm2, mv2, training=training,
                                       relu1=relu, relu2=relu_other, relu_out=relu_out, has_b=has_b, norm_b=norm_b)
    if training:
        tf.add_to_collection(tf.GraphKeys.UPDATE_OPS, tf.assign(moving_mean, outs[5]))
        tf.add_to_collection(tf.GraphKeys.UPDATE_OPS, tf.assign(moving_variance, outs[6]))
        if norm_b:
            tf.add_to_collection(tf.GraphKeys.UPDATE_OPS, tf.assign(mm2, outs[11]))
            tf.add_to_collection(tf.GraphKeys.UPDATE_OPS, tf.assign(mv2, outs[12]))
    return outs[0]


@ops.RegisterGradient("Sap3dBatchNormAct")
def _batch_norm_act_grad(op, dy, *_unused):
    a, _sa, _g1, _b1, _mm1, _mv1, b, _sb, _g2, _b2, _mm2, _mv2 = op.inputs
    o = op.outputs
    has_b, norm_b = op.get_attr("has_b"), op.get_attr("norm_b")
    da, db, dgamma1, dbeta1, dgamma2, dbeta2 = _sap3d.sap3d_batch_norm_act_grad(
        dy, a, o[1], o[2], o[3], o[4], b, o[7], o[8], o[9], o[10], training=op.get_attr("training"), relu1=op.get_attr("relu1"),
        relu2=op.get_attr("relu2"), relu_out=op.get_attr("relu_out"), has_b=has_b, norm_b=norm_b)
    two = has_b and norm_b
    # the statistics inputs carry no gradient of their own: the fused backward already goes through the batch statistics
    return (da, None, dgamma1, dbeta1, None, None, db if has_b else None, None, dgamma2 if two else None, dbeta2 if two else None, None, None)


def group_norm_act(x, gamma, beta, relu=False, other=None, other_norm=None, relu_other=False, relu_out=False):
    """GroupNorm (utils/network.py:65-87) [+ ReLU] [+ other, relu]; other_norm = (gamma, beta) of a second GroupNorm"""
    has_b, norm_b = other is not None, other_norm is not None
    b = other if has_b else x
    g2, b2 = other_norm if norm_b else (gamma, beta)
    return _sap3d.sap3d_group_norm_act(x, gamma, beta, b, g2, b2, relu1=relu, relu2=relu_other, relu_out=relu_out, has_b=has_b, norm_b=norm_b)[0]


@ops.RegisterGradient("Sap3dGroupNormAct")
def _group_norm_act_grad(op, dy, *_unused):
    a, gamma1, _beta1, b, gamma2, _beta2 = op.inputs
    o = op.outputs
    has_b, norm_b = op.get_attr("has_b"), op.get_attr("norm_b")
    da, db, dgamma1, dbeta1, dgamma2, dbeta2 = _sap3d.sap3d_group_norm_act_grad(
        dy, a, o[1], o[2], o[3], o[4], gamma1, b, o[5], o[6], o[7], o[8], gamma2, relu1=op.get_attr("relu1"), relu2=op.get_attr("relu2"),
        relu_out=op.get_attr("relu_out"), has_b=has_b, norm_b=norm_b)
    two = has_b and norm_b
    return da, dgamma1, dbeta1, (db if has_b else None), (dgamma2 if two else None), (dbeta2 if two else None)


ops.NotDifferentiable("Sap3dBatchNormActGrad")
ops.NotDifferentiable("Sap3dGroupNormActGrad")
ops.NotDifferentiable("Sap3dClipBatchNormAct")  # per-clip statistics: the gen_pred.py inference path only


# ---- CBAM block tail (gn/p3d_gn.py:175-177) ----------------------------------------------------------------------------
def cbam_block_tail(c3, gamma3, beta3, residual, w0, b0, w1, b1, w_sp):
    """relu(GroupNorm(c3) + cbam_block(residual))"""
    return _sap3d.sap3d_cbam_tail(c3, gamma3, beta3, residual, w0, b0, w1, b1, w_sp)[0]


@ops.RegisterGradient("Sap3dCbamTail")
def _cbam_tail_grad(op, dy, *_unused):
    c3, gamma3, _beta3, r, w0, _b0, w1, _b1, w_sp = op.inputs
    y, scale3, mean3, rstd3, cscale, sp, att, save = op.outputs
    # one gradient per input, in input order: c3, gamma3, beta3, r, w0, b0, w1, b1, w_sp
    return _sap3d.sap3d_cbam_tail_grad(dy, y, c3, scale3, mean3, rstd3, gamma3, r, w0, w1, w_sp, cscale, sp, att, save)


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
