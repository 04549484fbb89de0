"""TEST INFRASTRUCTURE -- a TensorFlow-1.x API emulation, just large enough to EXECUTE THE REFERENCE'S OWN graph-builder code
(/root/reference/p3d.py, utils/network.py, gn/p3d_gn.py) with no TensorFlow installed.

Why: TensorFlow cannot be installed in this image, so the oracle (oracle/p3d_oracle.py) is a restatement of the reference's
graphs.  The part of that restatement that IS citable in the reference -- the wiring: which layers exist, in which order,
with which strides / kernels / variable names / scopes, Python-2 integer division, dead branches, `training` never reaching
make_block -- is pinned here by running the reference's builder functions themselves, unmodified, against this module
registered as `tensorflow`.  What TensorFlow's kernels compute is not in the reference (un-vendored dependency); each tf.*
function below forwards to the SAME op-semantics restatement the oracle uses (oracle/tf_semantics.py), so a mismatch between
"reference code over this emulation" and "oracle" can only come from the wiring.

Variables: TF naming rules are re-implemented here independently of the oracle's -- tf.get_variable under tf.variable_scope,
tf.layers default names uniquified per enclosing variable scope (conv3d, conv3d_1, ...; batch_normalization_N), tf.Variable
under the NAME scope a re-entered variable_scope opens (group_norm, group_norm_1, ...) -- and every variable the reference
code creates must exist, with that name and shape, in the value store the oracle produced (and vice versa).
"""
from __future__ import annotations

import ast
import contextlib
import math
import re
import sys
import types
from typing import Dict, List

import torch

from oracle import tf_semantics as tfs


class Shape(tuple):
    def as_list(self):
        return list(self)


class TFTensor(torch.Tensor):
    """NDHWC tensor with the two TF-1.x shape accessors the reference uses"""

    def get_shape(self):
        return Shape(int(s) for s in self.shape)


def _t(x) -> TFTensor:
    if isinstance(x, TFTensor):
        return x
    return torch.as_tensor(x).as_subclass(TFTensor)


class Graph:
    """per-run state: variable store, scopes, creation log"""

    def __init__(self, values: Dict[str, torch.Tensor]):
        self.values = values
        self.created: List[str] = []
        self.var_scope: List[str] = []          # variable-scope path (affects tf.get_variable / tf.layers variable names)
        self.name_scope: List[str] = []         # name-scope path (affects tf.Variable names)
        self.layer_names: Dict[str, int] = {}   # default-name counters per enclosing variable scope
        self.name_scope_used: Dict[str, int] = {}
        self.reuse = False
        self.collections: Dict[str, list] = {}

    def variable(self, full_name: str, shape) -> TFTensor:
        if full_name not in self.values:
            raise KeyError(f"the reference code creates variable '{full_name}' {list(shape)}, which the oracle does not have")
        v = self.values[full_name]
        if tuple(v.shape) != tuple(int(s) for s in shape):
            raise ValueError(f"variable '{full_name}': reference shape {list(shape)} vs oracle {list(v.shape)}")
        if full_name not in self.created:
            self.created.append(full_name)
        return _t(v)

    def scoped(self, name: str) -> str:
        return "/".join(self.var_scope + [name])

    def unique_layer_name(self, base: str) -> str:
        key = "/".join(self.var_scope) + "|" + base
        n = self.layer_names.get(key, 0)
        self.layer_names[key] = n + 1
        return base if n == 0 else f"{base}_{n}"


G: Graph = None  # set by build_module()


# ---- scopes ----------------------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def variable_scope(name, default_name=None, reuse=None):
    # tf.variable_scope(name) pushes `name` on the variable-scope path (no uniquification) and opens a NAME scope that IS
    # uniquified (name, name_1, ...) -- tf.Variable names follow the name scope
    G.var_scope.append(name)
    key = "/".join(G.name_scope + [name])
    n = G.name_scope_used.get(key, 0)
    G.name_scope_used[key] = n + 1
    G.name_scope.append(name if n == 0 else f"{name}_{n}")
    old = G.reuse
    if reuse is not None:
        G.reuse = reuse
    try:
        yield
    finally:
        G.reuse = old
        G.var_scope.pop()
        G.name_scope.pop()


@contextlib.contextmanager
def name_scope(name):
    yield


@contextlib.contextmanager
def device(_):
    yield


# ---- variables -------------------------------------------------------------------------------------------------------------
def get_variable(name, shape=None, initializer=None, dtype=None, trainable=True, regularizer=None):
    return G.variable(G.scoped(name), shape)


def Variable(initial_value, dtype=None, name=None, trainable=True):
    return G.variable("/".join(G.name_scope + [name]), tuple(initial_value.shape))


def constant(value, shape=None, dtype=None):
    return _t(torch.full(tuple(shape), float(value)) if shape is not None else torch.tensor(value))


def constant_initializer(value=0.0):
    return ("constant", value)


def add_to_collection(name, value):
    G.collections.setdefault(name, []).append(value)


# ---- ops -------------------------------------------------------------------------------------------------------------------
def _same(padding):
    assert padding.lower() == "same", padding


def nn_conv3d(x, filt, strides, padding):
    _same(padding)
    assert strides[0] == 1 and strides[4] == 1
    return _t(tfs.conv3d_same(x, filt, tuple(strides[1:4])))


def nn_bias_add(x, b):
    return _t(x + b)


def nn_max_pool3d(x, ksize, strides, padding):
    _same(padding)
    return _t(tfs.max_pool3d_same(x, tuple(ksize[1:4]), tuple(strides[1:4])))


def _k3(k):
    return (k, k, k) if isinstance(k, int) else tuple(int(v) for v in k)


def layers_conv3d(inputs, filters, kernel_size, strides=(1, 1, 1), padding="valid", activation=None, use_bias=True, kernel_initializer=None,
                  kernel_regularizer=None, name=None, reuse=None):
    _same(padding)
    assert activation is None
    k, s = _k3(kernel_size), _k3(strides)
    lname = name if name is not None else G.unique_layer_name("conv3d")
    w = G.variable(G.scoped(lname + "/kernel"), (*k, inputs.shape[-1], filters))
    b = G.variable(G.scoped(lname + "/bias"), (filters,)) if use_bias else None
    return _t(tfs.conv3d_same(inputs, w, s, b))


def layers_conv3d_transpose(inputs, filters, kernel_size, strides=(1, 1, 1), padding="valid", activation=None, use_bias=True,
                            kernel_initializer=None, kernel_regularizer=None, name=None):
    _same(padding)
    k, s = _k3(kernel_size), _k3(strides)
    lname = name if name is not None else G.unique_layer_name("conv3d_transpose")
    w = G.variable(G.scoped(lname + "/kernel"), (*k, filters, inputs.shape[-1]))
    b = G.variable(G.scoped(lname + "/bias"), (filters,)) if use_bias else None
    return _t(tfs.conv3d_transpose_same(inputs, w, s, b))


def layers_max_pooling3d(inputs, pool_size, strides, padding="valid"):
    assert padding == "valid" and pool_size == strides
    return _t(tfs.max_pool3d_valid(inputs, int(pool_size)))


def layers_batch_normalization(inputs, training=False, name=None):
    lname = name if name is not None else G.unique_layer_name("batch_normalization")
    c = inputs.shape[-1]
    gamma = G.variable(G.scoped(lname + "/gamma"), (c,))
    beta = G.variable(G.scoped(lname + "/beta"), (c,))
    mm = G.variable(G.scoped(lname + "/moving_mean"), (c,))
    mv = G.variable(G.scoped(lname + "/moving_variance"), (c,))
    if not isinstance(training, bool):
        raise TypeError("the emulation needs a Python bool for `training`")
    return _t(tfs.batch_norm(inputs, gamma, beta, mm, mv, training)[0])


def layers_dropout(inputs, rate=0.5, training=False):
    if float(rate) != 0.0 and training:
        raise NotImplementedError("dropout rate > 0 in training mode is not deterministic; run the builders with rate 0")
    return inputs


def layers_dense(inputs, units, activation=None, kernel_initializer=None, bias_initializer=None, name=None, reuse=None):
    w = G.variable(G.scoped(name + "/kernel"), (inputs.shape[-1], units))
    b = G.variable(G.scoped(name + "/bias"), (units,))
    y = _t(torch.matmul(inputs, w) + b)
    return activation(y) if activation is not None else y


def _axes(axis):
    return tuple(axis) if isinstance(axis, (list, tuple)) else axis


def reduce_mean(x, axis=None, keepdims=False, keep_dims=False):
    x = _t(x)
    return x if (axis is None and x.dim() == 0) else _t(x.mean(dim=_axes(axis), keepdim=keepdims or keep_dims)) if axis is not None else _t(x.mean())


def reduce_max(x, axis=None, keepdims=False, keep_dims=False):
    return _t(x.amax(dim=_axes(axis), keepdim=keepdims or keep_dims))


def reduce_sum(x, axis=None, keepdims=False):
    return _t(x.sum()) if axis is None else _t(x.sum(dim=_axes(axis), keepdim=keepdims))


def moments(x, axes, keep_dims=False):
    return _t(x.mean(dim=tuple(axes), keepdim=keep_dims)), _t(x.var(dim=tuple(axes), unbiased=False, keepdim=keep_dims))


def matmul(a, b, transpose_b=False):
    return _t(torch.matmul(a, b.transpose(-1, -2) if transpose_b else b))


def reshape(x, shape, name=None):
    return _t(x.reshape([int(s) for s in shape]))


def shape(x):
    return [int(s) for s in x.shape]


def concat(values, axis, name=None):
    return _t(torch.cat(list(values), dim=axis))


def sigmoid(x, name=None):
    return _t(torch.sigmoid(x))


def build_module(values: Dict[str, torch.Tensor]) -> types.ModuleType:
    """a fresh `tensorflow` module object bound to a fresh graph state"""
    global G
    G = Graph(values)
    tf = types.ModuleType("tensorflow")
    tf.float32 = torch.float32
    tf.device, tf.variable_scope, tf.name_scope = device, variable_scope, name_scope
    tf.get_variable, tf.Variable, tf.constant, tf.constant_initializer, tf.add_to_collection = get_variable, Variable, constant, constant_initializer, add_to_collection
    tf.reduce_mean, tf.reduce_max, tf.reduce_sum = reduce_mean, reduce_max, reduce_sum
    tf.matmul, tf.reshape, tf.shape, tf.concat, tf.sigmoid = matmul, reshape, shape, concat, sigmoid
    tf.transpose = lambda x, perm: _t(x.permute(*perm))
    tf.sqrt = lambda x: _t(torch.sqrt(x))
    tf.abs = lambda x: _t(torch.abs(x))
    tf.pow = lambda x, y: _t(torch.pow(x, y))
    tf.less = lambda a, b: _t(torch.lt(a, b))
    tf.to_float = lambda x: _t(x.float())
    tf.stop_gradient = lambda x: _t(x.detach())
    tf.zeros_like = lambda x: _t(torch.zeros_like(x))
    tf.nn = types.SimpleNamespace(conv3d=nn_conv3d, bias_add=nn_bias_add, max_pool3d=nn_max_pool3d, relu=lambda x, name=None: _t(torch.relu(x)),
                                  softmax=lambda s, axis=-1: _t(torch.softmax(s, dim=axis)), moments=moments,
                                  l2_loss=lambda v: _t((v * v).sum() / 2))
    tf.layers = types.SimpleNamespace(conv3d=layers_conv3d, conv3d_transpose=layers_conv3d_transpose, max_pooling3d=layers_max_pooling3d,
                                      batch_normalization=layers_batch_normalization, dropout=layers_dropout, dense=layers_dense)
    tf.contrib = types.SimpleNamespace(layers=types.SimpleNamespace(xavier_initializer=lambda: ("xavier",),
                                                                    variance_scaling_initializer=lambda: ("variance_scaling",),
                                                                    l2_regularizer=lambda scale, scope=None: ("l2", scale)))
    return tf


# ---- loading the reference's Python-2 sources --------------------------------------------------------------------------------
def _py2_div(a, b):
    """Python-2 `/`: floor division for two ints (utils/network.py:182,187,188), true division otherwise"""
    if isinstance(a, int) and isinstance(b, int) and not isinstance(a, bool) and not isinstance(b, bool):
        return a // b
    return a / b


class _Py2Division(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(func=ast.Name(id="_py2_div", ctx=ast.Load()), args=[node.left, node.right], keywords=[]), node)
        return node


def load_reference_module(path: str, name: str, tf_module, extra_modules: Dict[str, types.ModuleType] = None) -> types.ModuleType:
    """executes a reference source file (read where it lies, never copied) as a module: Python-2 print statements become calls,
    `/` keeps its Python-2 meaning, `import tensorflow` resolves to the emulation"""
    src = open(path).read()
    src = re.sub(r"^(\s*)print (?!\()(.*)$", r"\1print(\2)", src, flags=re.M)
    tree = _Py2Division().visit(ast.parse(src, filename=path))
    ast.fix_missing_locations(tree)
    mod = types.ModuleType(name)
    mod.__file__ = path
    mod.__dict__["_py2_div"] = _py2_div
    saved = {k: sys.modules.get(k) for k in ["tensorflow"] + list(extra_modules or {})}
    sys.modules["tensorflow"] = tf_module
    for k, m in (extra_modules or {}).items():
        sys.modules[k] = m
    try:
        exec(compile(tree, path, "exec"), mod.__dict__)
    finally:
        for k, m in saved.items():
            if m is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = m
    return mod


def run_reference_builder(ref_root: str, module: str, builder: str, x: torch.Tensor, values: Dict[str, torch.Tensor], training: bool,
                          batch_size: int):
    """builds + evaluates reference graph `builder` of `module` ('p3d' or 'gn.p3d_gn') on x with the given variable values.
    Returns (output tensor, names of the variables the reference code created, in creation order)."""
    import os

    tf = build_module(values)
    network = load_reference_module(os.path.join(ref_root, "utils", "network.py"), "utils.network", tf)
    utils_pkg = types.ModuleType("utils")
    utils_pkg.network = network
    if module == "p3d":
        mod = load_reference_module(os.path.join(ref_root, "p3d.py"), "p3d", tf, {"utils": utils_pkg, "utils.network": network})
    else:   # gn/p3d_gn.py does a flat `from network import *` (the GN driver runs with utils/ on sys.path)
        mod = load_reference_module(os.path.join(ref_root, "gn", "p3d_gn.py"), "p3d_gn", tf, {"network": network})
    with torch.no_grad():
        out = getattr(mod, builder)(_t(x), 0.0, batch_size, training)
    return torch.Tensor(out).as_subclass(torch.Tensor), list(G.created)
