"""The TensorFlow custom-op boundary (tf_ops/): TensorFlow is not installable in the build image, so the shim is checked as far
as it can be without it:
  * tf_ops/sap3d_tf_ops.cc compiles (`g++ -fsyntax-only`) against stand-ins of the TF-1.15 headers it includes
    (tf_ops/tf_stub/) and against include/sap3d.h -- every C-ABI call is type-checked;
  * every op the gradient module (tf_ops/sap3d_grads.py) and INTEGRATION.md use is registered, with the number of inputs and the
    attribute names used there; every registered op is either differentiable through a registered gradient or declared
    NotDifferentiable;
  * every compute entry-point family of include/sap3d.h that replaces a reference call-site is reachable from some op."""
import ast
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CC = os.path.join(ROOT, "tf_ops", "sap3d_tf_ops.cc")
GRADS = os.path.join(ROOT, "tf_ops", "sap3d_grads.py")


def snake(name: str) -> str:
    """TensorFlow's op-name -> Python-function-name rule"""
    return re.sub(r"([a-z0-9])([A-Z])", r"\1_\2", name).lower()


def registered_ops():
    text = open(CC).read()
    text = re.sub(r"//[^\n]*", "", text)
    conv_attrs = re.search(r"#define SAP3D_CONV_ATTRS(.*?)\n\n", text, re.S).group(1)
    ops = {}
    for m in re.finditer(r'REGISTER_OP\("(\w+)"\)(.*?);', text, re.S):
        body = m.group(2).replace("SAP3D_CONV_ATTRS", conv_attrs)
        body = body.split(".SetShapeFn")[0]
        ins = re.findall(r'\.Input\("(\w+):', body)
        outs = re.findall(r'\.Output\("(\w+):', body)
        attrs = re.findall(r'\.Attr\("(\w+):', body)
        ops[m.group(1)] = (ins, outs, attrs)
    return ops


def test_shim_compiles_against_the_header_stand_ins():
    gxx = shutil.which("g++")
    assert gxx, "g++ not found"
    cuda_inc = "/usr/local/cuda/include"
    assert os.path.exists(os.path.join(cuda_inc, "cuda_runtime_api.h"))
    r = subprocess.run([gxx, "-std=c++14", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tf_ops", "tf_stub"), "-I",
                        os.path.join(ROOT, "include"), "-I", cuda_inc, CC], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_every_op_has_shape_function_and_gpu_kernel():
    text = open(CC).read()
    ops = registered_ops()
    assert len(ops) >= 31, sorted(ops)
    for name in ops:
        assert re.search(r'REGISTER_KERNEL_BUILDER\(Name\("%s"\)\.Device\(tf::DEVICE_GPU\)' % name, text), name
        block = text[text.index('REGISTER_OP("%s")' % name):]
        block = block[:block.index("REGISTER_KERNEL_BUILDER")]
        assert ".SetShapeFn(" in block, f"{name} has no shape function (downstream layers would lose their static shapes)"


def _sap3d_calls(tree):
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name) \
                and node.func.value.id == "_sap3d":
            yield node


def test_gradient_module_uses_only_registered_ops_with_their_signatures():
    ops = registered_ops()
    by_func = {snake(n): n for n in ops}
    tree = ast.parse(open(GRADS).read())
    used = set()
    for call in _sap3d_calls(tree):
        fn = call.func.attr
        assert fn in by_func, f"sap3d_grads.py calls _sap3d.{fn}, which no REGISTER_OP provides"
        ins, _, attrs = ops[by_func[fn]]
        used.add(by_func[fn])
        if not any(isinstance(a, ast.Starred) for a in call.args):
            assert len(call.args) == len(ins), (fn, len(call.args), ins)
        for kw in call.keywords:
            if kw.arg is not None:
                assert kw.arg in attrs, (fn, kw.arg, attrs)
    grads, nondiff = set(), set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in ("RegisterGradient", "NotDifferentiable"):
            name = node.args[0].value
            assert name in ops, name
            (grads if node.func.attr == "RegisterGradient" else nondiff).add(name)
    assert not (grads & nondiff)
    assert grads | nondiff == set(ops), sorted(set(ops) - grads - nondiff)
    # the forward ops of the training graph are differentiable
    for name in ("Sap3dConv", "Sap3dBatchNormAct", "Sap3dGroupNormAct", "Sap3dMaxPool3d", "Sap3dFlashAttention", "Sap3dAttention", "Sap3dGate", "Sap3dHead",
                 "Sap3dSmoothL1Loss", "Sap3dDropout", "Sap3dCbamTail", "Sap3dConcatChannels"):
        assert name in grads, name


def test_integration_doc_uses_only_registered_ops():
    ops = registered_ops()
    by_func = {snake(n) for n in ops}
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for fn in set(re.findall(r"_sap3d\.(\w+)", doc)):
        assert fn in by_func, f"INTEGRATION.md uses _sap3d.{fn}, which is not a registered op"
    helpers = {n.name for n in ast.parse(open(GRADS).read()).body if isinstance(n, ast.FunctionDef)}
    for fn in set(re.findall(r"sap3d_grads\.(\w+)\(", doc)):
        assert fn in helpers, f"INTEGRATION.md uses sap3d_grads.{fn}, which does not exist"


def test_reference_call_site_families_are_reachable_from_the_ops():
    text = open(CC).read()
    for entry in ("sap3d_conv_fwd", "sap3d_conv_fwd_affine", "sap3d_conv_dgrad", "sap3d_conv_wgrad", "sap3d_conv_pack_weights", "sap3d_bn_finalize",
                  "sap3d_affine_act", "sap3d_affine_act_bwd", "sap3d_gn_stats", "sap3d_gn_act_bwd", "sap3d_cbam_fwd", "sap3d_cbam_merge",
                  "sap3d_cbam_tail_bwd", "sap3d_maxpool3d_fwd", "sap3d_maxpool3d_bwd", "sap3d_flash_attn_fwd", "sap3d_flash_attn_bwd",
                  "sap3d_attention_fwd", "sap3d_attention_bwd", "sap3d_gate_fwd", "sap3d_gate_bwd", "sap3d_head_fwd", "sap3d_head_bwd",
                  "sap3d_loss_smooth_l1_ex", "sap3d_dropout", "sap3d_concat_channels", "sap3d_split_channels", "sap3d_adam_step",
                  "sap3d_saliency_metrics", "sap3d_resize_bilinear", "sap3d_saliency_auc", "sap3d_preprocess_frames", "sap3d_sample_norm_apply"):
        assert re.search(r"\b%s\(" % entry, text), entry
