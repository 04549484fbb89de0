"""Engine ops of the GroupNorm + CBAM model variant (gn/p3d_gn.py): per-sample GroupNorm statistics ->
fused affine/ReLU/add pass, and the CBAM-on-residual block tail.  Forward path (inference / parity); the
training backward of these two ops is the next item of the build plan."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _abi as A
from .engine import ConvOut, Engine, Param, T

GN_EPS = 1e-5   # utils/network.py:65
GN_GROUPS = 32


class GNState:
    def __init__(self, eng: Engine, N: int, Cc: int, S: int, gamma: Param, beta: Param):
        self.gamma, self.beta = gamma, beta
        self.G = min(GN_GROUPS, Cc)
        self.rows = A.lib.sap3d_sample_stats_rows(S, Cc, N)
        dev = eng.device
        self.part = torch.empty(N, self.rows, 3, Cc, device=dev, dtype=torch.float32)
        self.scale = torch.empty(N, Cc, device=dev, dtype=torch.float32)
        self.shift = torch.empty(N, Cc, device=dev, dtype=torch.float32)
        self.mean = torch.empty(N, self.G, device=dev, dtype=torch.float32)
        self.rstd = torch.empty(N, self.G, device=dev, dtype=torch.float32)


def _no_backward(name):
    def f():
        raise A.Sap3dError(f"backward of {name} (GroupNorm/CBAM graph) is not implemented in this round")
    return f


class GNActOp:
    """y = relu_out?( relu1?(GN1(a)) + relu2?(GN2(b) | b) ) with per-sample group statistics"""

    def __init__(self, eng: Engine, a: ConvOut, g1: GNState, relu1, b=None, g2: Optional[GNState] = None, relu2=False,
                 relu_out=False, name=""):
        self.eng, self.a, self.g1, self.relu1 = eng, a, g1, relu1
        self.b, self.g2, self.relu2, self.relu_out, self.name = b, g2, relu2, relu_out, name
        self.b_t = None if b is None else (b.raw if isinstance(b, ConvOut) else b)
        self.y = eng.tensor(a.raw.shape, name)
        eng.fwd_ops.append(self.fwd)
        eng.bwd_ops.append(_no_backward(name) if eng.training_graph else (lambda: None))

    def _stats(self, t: T, gs: GNState):
        e = self.eng
        N, S, Cc = t.shape[0], t.positions // t.shape[0], t.C
        A.check(A.lib.sap3d_sample_channel_partials(e.dt, A.ptr(t.buf), None, N, S, Cc, gs.rows, A.ptr(gs.part), e.stream), "gn partials")
        A.check(A.lib.sap3d_gn_finalize(A.ptr(gs.part), gs.rows, N, S, Cc, gs.G, A.ptr(gs.gamma.w), A.ptr(gs.beta.w), GN_EPS,
                                        A.ptr(gs.scale), A.ptr(gs.shift), A.ptr(gs.mean), A.ptr(gs.rstd), e.stream), "gn finalize")
        e._count(2)

    def fwd(self):
        e = self.eng
        self._stats(self.a.raw, self.g1)
        if self.g2 is not None:
            self._stats(self.b_t, self.g2)
        y = self.y
        S = y.positions // y.shape[0]
        A.check(A.lib.sap3d_affine_act(e.dt, A.ptr(self.a.raw.buf), A.ptr(self.g1.scale), A.ptr(self.g1.shift), int(self.relu1),
                                       A.ptr(self.b_t.buf) if self.b_t is not None else None,
                                       A.ptr(self.g2.scale) if self.g2 else None, A.ptr(self.g2.shift) if self.g2 else None,
                                       int(self.relu2), int(self.relu_out), A.ptr(y.buf), y.positions, y.C, S, e.stream),
                "gn affine_act " + self.name)
        e._count()


class CbamBlockTailOp:
    """out = relu(GN(c3) + cbam_block(residual))  (gn/p3d_gn.py:175-177; utils/network.py:198-274)"""

    def __init__(self, eng: Engine, c3: ConvOut, g3: GNState, residual: T, w0: Param, b0: Param, w1: Param, b1: Param,
                 w_sp: Param, name=""):
        self.eng, self.c3, self.g3, self.r, self.name = eng, c3, g3, residual, name
        self.w0, self.b0, self.w1, self.b1, self.w_sp = w0, b0, w1, b1, w_sp
        N, D, H, W, Cc = residual.shape
        dev = eng.device
        S = D * H * W
        self.rows = A.lib.sap3d_sample_stats_rows(S, Cc, N)
        self.part = torch.empty(N, self.rows, 3, Cc, device=dev, dtype=torch.float32)
        self.cscale = torch.empty(N, Cc, device=dev, dtype=torch.float32)
        self.sp = torch.empty(N, S, 2, device=dev, dtype=torch.float32)
        self.att = torch.empty(N, S, device=dev, dtype=torch.float32)
        self.y = eng.tensor(residual.shape, name)
        eng.fwd_ops.append(self.fwd)
        eng.bwd_ops.append(_no_backward(name) if eng.training_graph else (lambda: None))

    def fwd(self):
        e = self.eng
        N, D, H, W, Cc = self.r.shape
        S = D * H * W
        raw, g3 = self.c3.raw, self.g3
        A.check(A.lib.sap3d_sample_channel_partials(e.dt, A.ptr(raw.buf), None, N, S, Cc, g3.rows, A.ptr(g3.part), e.stream), "gn partials")
        A.check(A.lib.sap3d_gn_finalize(A.ptr(g3.part), g3.rows, N, S, Cc, g3.G, A.ptr(g3.gamma.w), A.ptr(g3.beta.w), GN_EPS,
                                        A.ptr(g3.scale), A.ptr(g3.shift), A.ptr(g3.mean), A.ptr(g3.rstd), e.stream), "gn finalize")
        A.check(A.lib.sap3d_cbam_fwd(e.dt, A.ptr(self.r.buf), N, D, H, W, Cc, self.w0.shape[1], A.ptr(self.w0.w), A.ptr(self.b0.w),
                                     A.ptr(self.w1.w), A.ptr(self.b1.w), A.ptr(self.w_sp.w), A.ptr(self.part), self.rows,
                                     A.ptr(self.cscale), A.ptr(self.sp), A.ptr(self.att), e.stream), "cbam_fwd " + self.name)
        A.check(A.lib.sap3d_cbam_merge(e.dt, A.ptr(raw.buf), A.ptr(g3.scale), A.ptr(g3.shift), A.ptr(self.r.buf), A.ptr(self.cscale),
                                       A.ptr(self.att), A.ptr(self.y.buf), N, S, Cc, e.stream), "cbam_merge " + self.name)
        e._count(7)


class ConcatOp:
    """materialised tf.concat([a, b], -1) (only needed for three-way concatenations)"""

    def __init__(self, eng: Engine, a: T, b: T, name=""):
        self.eng, self.a, self.b = eng, a, b
        self.y = eng.tensor((*a.shape[:4], a.C + b.C), name)
        eng.fwd_ops.append(self.fwd)
        eng.bwd_ops.append(_no_backward(name) if eng.training_graph else (lambda: None))

    def fwd(self):
        e = self.eng
        A.check(A.lib.sap3d_concat_channels(e.dt, A.ptr(self.a.buf), A.ptr(self.b.buf), A.ptr(self.y.buf), self.a.positions,
                                            self.a.C, self.b.C, e.stream), "concat")
        e._count()
