"""ctypes binding of libsap3d_b200.so (the C ABI declared in include/sap3d.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``python sap3d_build.py``.
There is no fallback: if the shared object is missing, import of this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsap3d_b200.so")

BF16, F32 = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2


class Sap3dError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA extension first (python sap3d_build.py). "
        "This framework has no CPU / PyTorch fallback path."
    )

lib = C.CDLL(LIB_PATH)


class ConvDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("impl", C.c_int32),
        ("N", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("nseg", C.c_int32), ("cin", C.c_int32 * 2), ("cout", C.c_int32),
        ("kd", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("sd", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("transposed", C.c_int32), ("has_bias", C.c_int32), ("out_f32", C.c_int32),
    ]


_vp = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f32 = C.c_float

lib.sap3d_last_error.restype = C.c_char_p
lib.sap3d_abi_version.restype = C.c_int
lib.sap3d_device_ok.restype = C.c_int


def _sig(name, argtypes, restype=C.c_int):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
    return fn


_P = C.POINTER
_sig("sap3d_debug_conv_timing", [_vp])
_sig("sap3d_debug_conv_halo_launches", [], C.c_longlong)
_sig("sap3d_debug_conv_swap_launches", [], C.c_longlong)
_sig("sap3d_conv_out_dims", [_P(ConvDesc), _P(C.c_int32)])
_sig("sap3d_conv_stats_rows", [_P(ConvDesc)])
_sig("sap3d_conv_packed_elems", [_P(ConvDesc), _i32], C.c_size_t)
_sig("sap3d_conv_pack_weights", [_P(ConvDesc), _vp, _vp, _vp, _vp])
_sig("sap3d_conv_fwd", [_P(ConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp])
_sig("sap3d_conv_fwd_on_tensor_cores", [_P(ConvDesc)])
_sig("sap3d_conv_fwd_operand_is_workspace", [_P(ConvDesc)])
class BnFuse(C.Structure):
    _fields_ = [("gamma", C.c_void_p), ("beta", C.c_void_p), ("moving_mean", C.c_void_p), ("moving_var", C.c_void_p),
                ("momentum", C.c_float), ("eps", C.c_float), ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p),
                ("rstd", C.c_void_p), ("relu1", C.c_int32), ("residual", C.c_void_p), ("relu_out", C.c_int32), ("y", C.c_void_p)]


_sig("sap3d_conv_fwd_bn_supported", [_P(ConvDesc)])
_sig("sap3d_conv_fwd_bn", [_P(ConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _P(BnFuse), _vp])
_sig("sap3d_conv_fwd_affine", [_P(ConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp])
_sig("sap3d_conv_dgrad", [_P(ConvDesc), _i32, _vp, _vp, _vp, _vp, _i32, _vp])
_sig("sap3d_conv_dgrad2_supported", [_P(ConvDesc)])
_sig("sap3d_conv_dgrad2", [_P(ConvDesc), _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp])
_sig("sap3d_conv_wgrad", [_P(ConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp])


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise Sap3dError(f"{what}: {lib.sap3d_last_error().decode()}")


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else t.data_ptr()


def make_conv_desc(dtype, N, D, H, W, cin, cout, kernel, strides, transposed=False, has_bias=False, out_f32=False,
                   impl=IMPL_AUTO) -> ConvDesc:
    d = ConvDesc()
    d.dtype, d.impl = dtype, impl
    d.N, d.D, d.H, d.W = N, D, H, W
    cin = list(cin)
    d.nseg = len(cin)
    d.cin[0] = cin[0]
    d.cin[1] = cin[1] if len(cin) > 1 else 0
    d.cout = cout
    d.kd, d.kh, d.kw = kernel
    d.sd, d.sh, d.sw = strides
    d.transposed = int(transposed)
    d.has_bias = int(has_bias)
    d.out_f32 = int(out_f32)
    return d


def conv_out_dims(d: ConvDesc):
    out = (C.c_int32 * 3)()
    check(lib.sap3d_conv_out_dims(C.byref(d), out), "conv_out_dims")
    return tuple(out)


_f64 = C.c_double
_u64 = C.c_uint64
_sig("sap3d_bn_finalize", [_vp, _i32, _i32, _f64, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _vp, _vp, _vp, _vp, _vp])
_sig("sap3d_gn_stats", [_i32, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp])
_sig("sap3d_affine_act", [_i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _i32, _i64, _vp])
_sig("sap3d_sample_norm_apply_supported", [_i32, _i64, _i32])
_sig("sap3d_sample_norm_apply", [_i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _i32, _i64, _i32, _f32, _vp])
_sig("sap3d_bn_apply_fused", [_i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32,
                              _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _i32, _f64, _f32,
                              _f32, _vp])
_sig("sap3d_affine_act_bwd_workspace", [_i32], C.c_size_t)
_sig("sap3d_affine_act_bwd", [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _i32,
                              _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp])
_sig("sap3d_affine_act_bwd_sync", [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _i32,
                                   _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _f64, _i32])
_I3 = C.c_int32 * 3
_sig("sap3d_maxpool3d_out_dims", [_i32, _i32, _i32, _P(C.c_int32), _P(C.c_int32), _i32, _P(C.c_int32)])
_sig("sap3d_maxpool3d_fwd", [_i32, _vp, _i32, _i32, _i32, _i32, _i32, _P(C.c_int32), _P(C.c_int32), _i32, _vp, _vp, _vp])
_sig("sap3d_maxpool3d_bwd", [_i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _P(C.c_int32), _P(C.c_int32), _i32, _vp, _vp, _i32, _vp])
_sig("sap3d_head_fwd", [_i32, _vp, _i32, _i32, _i32, _i32, _i32, _P(C.c_int32), _i32, _vp, _vp, _vp, _vp, _vp])
_sig("sap3d_head_bwd", [_i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _P(C.c_int32), _i32, _vp, _vp, _i32, _vp, _vp])
_sig("sap3d_head_tc_workspace", [_i32, _i32, _i32, _i32, _i32], C.c_size_t)
_sig("sap3d_head_tc_fwd", [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp])
_sig("sap3d_head_tc_bwd", [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp])
_sig("sap3d_loss_smooth_l1", [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp])
_sig("sap3d_loss_smooth_l1_ex", [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _vp])
_sig("sap3d_dropout", [_i32, _vp, _vp, _i64, _f32, _u64, _vp, _i32, _vp])
_sig("sap3d_gate_fwd", [_i32, _vp, _vp, _vp, _vp, _i64, _vp])
_sig("sap3d_gate_bwd", [_i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp])
_sig("sap3d_adam_step", [_vp, _vp, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _f32, _f32, _vp])
_sig("sap3d_adam_step_g", [_vp, _vp, _i32, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _f32, _f32, _vp])
_sig("sap3d_step_increment", [_vp, _vp])
_sig("sap3d_cast", [_i32, _vp, _vp, _i64, _vp])
_i32x = _i32
_sig("sap3d_attention_fwd", [_i32, _vp, _vp, _vp, _vp, _vp] + [_i32] * 10 + [_vp])
_sig("sap3d_attention_bwd", [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp] + [_i32] * 10 + [_vp])
_sig("sap3d_flash_attn_fwd", [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp])
_sig("sap3d_flash_attn_bwd_workspace", [_i32, _i32, _i32, _i32], C.c_size_t)
_sig("sap3d_flash_attn_bwd", [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp])
_sig("sap3d_gemm_nt", [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp])
_sig("sap3d_gemm_nt_batched", [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp])
_sig("sap3d_gemm_tn", [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp])
_sig("sap3d_softmax_rows", [_i32, _vp, _vp, _i64, _i32, _i32, _i32, _vp])
_sig("sap3d_softmax_bwd_rows", [_vp, _vp, _i64, _i32, _i32, _vp])
_sig("sap3d_transpose", [_i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _vp])
_sig("sap3d_pad_channels", [_i32, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp])


class PackEntry(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("taps", C.c_int32), ("rows", C.c_int32), ("rows_pad", C.c_int32),
                ("cols", C.c_int32), ("s_tap", C.c_int64), ("s_r", C.c_int64), ("s_c", C.c_int64), ("start", C.c_int64)]


_sig("sap3d_conv_pack_entries", [_P(ConvDesc), _vp, _vp, _vp, _P(PackEntry)])
_sig("sap3d_pack_multi", [_vp, _i32, _i64, _vp])
_sig("sap3d_sample_stats_rows", [_i64, _i32, _i32])
_sig("sap3d_sample_channel_partials", [_i32, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp])
_sig("sap3d_gn_finalize", [_vp, _i32, _i32, _i64, _i32, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp])
_sig("sap3d_cbam_fwd", [_i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp])
_sig("sap3d_split_channels", [_i32, _vp, _vp, _i32, _vp, _i32, _i64, _i32, _i32, _vp])
_sig("sap3d_gn_bwd_workspace", [_i32, _i64, _i32], C.c_size_t)
_sig("sap3d_gn_act_bwd", [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, _i32, _i32,
                          _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp])
_sig("sap3d_cbam_tail_bwd", [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp,
                             _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp])
_sig("sap3d_cbam_merge", [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp])
_sig("sap3d_concat_channels", [_i32, _vp, _vp, _vp, _i64, _i32, _i32, _vp])
_sig("sap3d_preprocess_frames", [_vp, _i32, _i32, _i32, _P(C.c_float), _i32, _vp, _i32, _i32, _vp])
_sig("sap3d_crc32c", [C.c_uint32, _vp, C.c_size_t], C.c_uint32)
_sig("sap3d_nan_sum_count", [_vp, _i32, _i32, _vp, _vp, _vp])
_sig("sap3d_resize_bilinear", [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _vp])
_sig("sap3d_saliency_auc_workspace", [_i32, _i32], C.c_size_t)
_sig("sap3d_saliency_auc", [_vp, _vp, _i32, _i64, _i32, _i32, _f64, _u64, _vp, _vp, _vp])
_sig("sap3d_saliency_metrics", [_vp, _vp, _vp, _i32, _i64, _i64, _i64, _i64, _vp, _vp])


def i3(v):
    return _I3(*[int(a) for a in v])
