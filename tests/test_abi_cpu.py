"""CPU tests of the drop-in boundary: the library builds, loads and exports every symbol declared in
include/sap3d.h; geometry helpers that need no device; compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sap3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sap3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib_built):
    lib = C.CDLL(lib_built)
    syms = declared_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_conv_geometry_matches_tf_same(lib_built):
    from sap3d_tensorflow_b200 import _abi as A

    d = A.make_conv_desc(A.BF16, 2, 16, 112, 112, [3], 64, (1, 7, 7), (1, 2, 2))
    assert A.conv_out_dims(d) == (16, 56, 56)
    d = A.make_conv_desc(A.BF16, 2, 2, 14, 14, [512, 512], 512, (2, 3, 3), (1, 1, 1))
    assert A.conv_out_dims(d) == (2, 14, 14)
    d = A.make_conv_desc(A.BF16, 2, 1, 7, 7, [1024], 512, (1, 3, 3), (2, 2, 2), transposed=True)
    assert A.conv_out_dims(d) == (2, 14, 14)
    d = A.make_conv_desc(A.BF16, 2, 4, 28, 28, [256], 128, (1, 1, 1), (1, 2, 2))
    assert A.conv_out_dims(d) == (4, 14, 14)
    # packed weight sizes: [cout_pad][taps*cin] and [cin_pad][taps*cout]
    d = A.make_conv_desc(A.BF16, 1, 8, 56, 56, [128, 128], 128, (3, 3, 3), (1, 1, 1))
    assert A.lib.sap3d_conv_packed_elems(C.byref(d), 0) == 128 * 27 * 256
    assert A.lib.sap3d_conv_packed_elems(C.byref(d), 1) == 256 * 27 * 128
    assert A.lib.sap3d_conv_stats_rows(C.byref(d)) == 8 * 56 * 56 // 128


def test_no_cpu_fallback(lib_built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sap3d_tensorflow_b200 import _abi as A
    import sap3d_tensorflow_b200 as sp

    assert A.lib.sap3d_device_ok() == 0
    rc = A.lib.sap3d_adam_step(None, None, None, None, 0, None, 1e-4, 0.9, 0.999, 1e-8, 1.0, None)
    assert rc != 0 and b"no CUDA device" in A.lib.sap3d_last_error()
    with pytest.raises(A.Sap3dError):
        sp.placeholder([1, 16, 32, 32, 3])


def test_tf_name_uniquifier():
    from sap3d_tensorflow_b200.engine import NameScope

    ns = NameScope()
    assert [ns.unique("", "conv3d") for _ in range(3)] == ["conv3d", "conv3d_1", "conv3d_2"]
    assert ns.unique("x_4_0_sa", "conv3d") == "x_4_0_sa/conv3d"
    assert ns.unique("x_4_0_sa", "conv3d") == "x_4_0_sa/conv3d_1"
    assert ns.unique("", "batch_normalization") == "batch_normalization"


def test_dropout_hash_is_deterministic_and_unbiased():
    from oracle.dropout_hash import keep_mask

    m1, m2 = keep_mask(7, 100000, 0.5), keep_mask(7, 100000, 0.5)
    assert (m1 == m2).all() and abs(m1.mean() - 0.5) < 0.01
    assert abs(keep_mask(8, 100000, 0.25).mean() - 0.75) < 0.01


def _prototypes():
    """{name: number of parameters} of every function declared in include/sap3d.h"""
    text = open(os.path.join(ROOT, "include", "sap3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"\b(sap3d_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def test_ctypes_bindings_match_the_header_prototypes(lib_built):
    """every function the Python host binds (_abi.py: argtypes) takes exactly as many arguments as include/sap3d.h declares:
    a binding that drifts from the header would otherwise only show up as a wrong value or a crash on the GPU"""
    from sap3d_tensorflow_b200 import _abi as A

    protos = _prototypes()
    assert len(protos) >= 60
    checked = 0
    for name, n in protos.items():
        fn = getattr(A.lib, name)
        if fn.argtypes is None:      # not bound with a signature (restype-only helpers)
            continue
        assert len(fn.argtypes) == n, (name, len(fn.argtypes), n)
        checked += 1
    assert checked >= 55, checked
