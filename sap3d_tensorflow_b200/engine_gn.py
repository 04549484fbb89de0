"""Engine ops of the GroupNorm + CBAM model variant (gn/p3d_gn.py): per-sample GroupNorm statistics ->
fused affine/ReLU/add pass, and the CBAM-on-residual block tail, forward and backward (training driver:
gn/train_p3d_gn_dataset.py:169-199)."""
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


def _reserve_ws(eng: Engine, N: int, S: int, Cc: int):
    """one shared backward workspace per engine (ops run in stream order), sized for the largest user"""
    need = A.lib.sap3d_gn_bwd_workspace(N, S, Cc)
    eng._gn_ws_bytes = max(getattr(eng, "_gn_ws_bytes", 0), need)


def _ws(eng: Engine) -> torch.Tensor:
    ws = getattr(eng, "_gn_ws", None)
    if ws is None or ws.numel() * 4 < eng._gn_ws_bytes:
        ws = eng._gn_ws = torch.zeros(eng._gn_ws_bytes // 4 + 16, device=eng.device, dtype=torch.float32)
    return ws


class GNActOp:
    """y = relu_out?( relu1?(GN1(a)) + relu2?(GN2(b) | b) ) with per-sample group statistics"""

    def __init__(self, eng: Engine, a: ConvOut, g1: GNState, relu1, b=None, g2: Optional[GNState] = None, relu2=False,
                 relu_out=False, name=""):
        self.eng, self.a, self.g1, self.relu1 = eng, a, g1, relu1
        self.b, self.g2, self.relu2, self.relu_out, self.name = b, g2, relu2, relu_out, name
        self.b_t = None if b is None else (b.raw if isinstance(b, ConvOut) else b)
        self.y = eng.tensor(a.raw.shape, name)
        eng.fwd_ops.append(self.fwd)
        eng.bwd_ops.append(self.bwd)
        if eng.training_graph:
            t = a.raw
            _reserve_ws(eng, t.shape[0], t.positions // t.shape[0], t.C)

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


    def bwd(self):
        e = self.eng
        if not self.y.gflag:
            return
        g1, g2, a_raw, b_t = self.g1, self.g2, self.a.raw, self.b_t
        acc_a = a_raw.take_acc()
        db_ptr, acc_b = None, 0
        if b_t is not None and b_t.needs_grad:
            acc_b = b_t.take_acc()
            db_ptr = A.ptr(b_t.ensure_grad())
        N, S, Cc = a_raw.shape[0], a_raw.positions // a_raw.shape[0], a_raw.C
        A.check(A.lib.sap3d_gn_act_bwd(
            e.dt, A.ptr(self.y.grad), A.ptr(a_raw.buf), A.ptr(g1.scale), A.ptr(g1.shift), A.ptr(g1.mean), A.ptr(g1.rstd),
            A.ptr(g1.gamma.w), int(self.relu1), A.ptr(b_t.buf) if b_t is not None else None,
            A.ptr(g2.scale) if g2 else None, A.ptr(g2.shift) if g2 else None, A.ptr(g2.mean) if g2 else None,
            A.ptr(g2.rstd) if g2 else None, A.ptr(g2.gamma.w) if g2 else None, int(self.relu2), int(self.relu_out), N, S, Cc, g1.G,
            A.ptr(a_raw.ensure_grad()), acc_a, db_ptr, acc_b, A.ptr(g1.gamma.g), A.ptr(g1.beta.g),
            A.ptr(g2.gamma.g) if g2 else None, A.ptr(g2.beta.g) if g2 else None, A.ptr(_ws(e)), e.stream), "gn_act_bwd " + self.name)
        e._count(3)


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
        self.hidden = w0.shape[1]
        self.save = torch.empty(N, 2 * Cc + 2 * self.hidden, device=dev, dtype=torch.float32) if eng.training_graph else None
        self.y = eng.tensor(residual.shape, name)
        eng.fwd_ops.append(self.fwd)
        eng.bwd_ops.append(self.bwd)
        if eng.training_graph:
            c3.raw.ensure_grad()
            _reserve_ws(eng, N, S, Cc)

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
                                     A.ptr(self.cscale), A.ptr(self.sp), A.ptr(self.att), A.ptr(self.save), e.stream), "cbam_fwd " + self.name)
        A.check(A.lib.sap3d_cbam_merge(e.dt, A.ptr(raw.buf), A.ptr(g3.scale), A.ptr(g3.shift), A.ptr(self.r.buf), A.ptr(self.cscale),
                                       A.ptr(self.att), A.ptr(self.y.buf), N, S, Cc, e.stream), "cbam_merge " + self.name)
        e._count(7)


    def bwd(self):
        e = self.eng
        if not self.y.gflag:
            return
        N, D, H, W, Cc = self.r.shape
        raw, g3, r = self.c3.raw, self.g3, self.r
        acc_c3 = raw.take_acc()
        dr_ptr, acc_r = None, 0
        if r.needs_grad:
            acc_r = r.take_acc()
            dr_ptr = A.ptr(r.ensure_grad())
        A.check(A.lib.sap3d_cbam_tail_bwd(
            e.dt, A.ptr(self.y.grad), A.ptr(self.y.buf), A.ptr(raw.buf), A.ptr(g3.scale), A.ptr(g3.mean), A.ptr(g3.rstd),
            A.ptr(g3.gamma.w), A.ptr(r.buf), N, D, H, W, Cc, g3.G, self.hidden, A.ptr(self.w0.w), A.ptr(self.w1.w), A.ptr(self.w_sp.w),
            A.ptr(self.cscale), A.ptr(self.sp), A.ptr(self.att), A.ptr(self.save), A.ptr(raw.ensure_grad()), acc_c3, dr_ptr, acc_r,
            A.ptr(g3.gamma.g), A.ptr(g3.beta.g), A.ptr(self.w0.g), A.ptr(self.b0.g), A.ptr(self.w1.g), A.ptr(self.b1.g),
            A.ptr(self.w_sp.g), A.ptr(_ws(e)), e.stream), "cbam_tail_bwd " + self.name)
        e._count(8)


class CbamOp:
    """stand-alone cbam_block / channel_attention / spatial_attention of utils/network.py:198-274 on an arbitrary feature map
    (the block tail above is the fused form the GN backbone uses).  mode: 'both' | 'channel' | 'spatial'.
    Forward: sap3d_cbam_fwd (attention maps) + sap3d_cbam_merge with no main branch (y = x * cscale * att, no ReLU).
    Backward: the tail kernels with the main branch switched off -- zero main-branch operand, an all-positive stand-in for the
    ReLU mask -- so the CBAM gradient code is the one the hot path exercises."""

    def __init__(self, eng: Engine, x: T, mode: str, w0=None, b0=None, w1=None, b1=None, w_sp=None, name=""):
        assert mode in ("both", "channel", "spatial")
        self.eng, self.x, self.mode, self.name = eng, x, mode, name
        self.w0, self.b0, self.w1, self.b1, self.w_sp = w0, b0, w1, b1, w_sp
        N, D, H, W, Cc = x.shape
        if Cc % 8 != 0:
            raise A.Sap3dError("cbam_block: channels must be a multiple of 8")
        dev, S = eng.device, D * H * W
        self.hidden = Cc // 8
        self.rows = A.lib.sap3d_sample_stats_rows(S, Cc, N)
        f32 = dict(device=dev, dtype=torch.float32)
        self.part = torch.zeros(N, self.rows, 3, Cc, **f32)
        self.cscale = torch.ones(N, Cc, **f32)          # stays 1 in 'spatial' mode
        self.sp = torch.zeros(N, S, 2, **f32)
        self.att = torch.ones(N, S, **f32)              # stays 1 in 'channel' mode
        self.save = torch.zeros(N, 2 * Cc + 2 * self.hidden, **f32)
        self.y = eng.tensor(x.shape, name)
        eng.fwd_ops.append(self.fwd)
        eng.bwd_ops.append(self.bwd)
        if eng.training_graph:
            _reserve_ws(eng, N, S, Cc)
            G = min(GN_GROUPS, Cc)
            self.G = G
            self.zero_act = torch.zeros(x.shape, device=dev, dtype=eng.tdt)      # main-branch operand (none)
            self.ones_act = torch.ones(x.shape, device=dev, dtype=eng.tdt)       # ReLU mask stand-in: everything passes
            self.zc = torch.zeros(N, Cc, **f32)
            self.zg = torch.zeros(N, G, **f32)
            self.zw = torch.zeros(max(Cc * self.hidden, 686), **f32)             # stand-in weights / gradient sinks
            self.sink = torch.zeros(max(Cc * self.hidden, 686), **f32)

    def fwd(self):
        e = self.eng
        N, D, H, W, Cc = self.x.shape
        ch = self.mode in ("both", "channel")
        spt = self.mode in ("both", "spatial")
        A.check(A.lib.sap3d_cbam_fwd(e.dt, A.ptr(self.x.buf), N, D, H, W, Cc, self.hidden,
                                     A.ptr(self.w0.w) if ch else None, A.ptr(self.b0.w) if ch else None,
                                     A.ptr(self.w1.w) if ch else None, A.ptr(self.b1.w) if ch else None,
                                     A.ptr(self.w_sp.w) if spt else None, A.ptr(self.part), self.rows, A.ptr(self.cscale),
                                     A.ptr(self.sp), A.ptr(self.att), A.ptr(self.save), e.stream), "cbam_fwd " + self.name)
        A.check(A.lib.sap3d_cbam_merge(e.dt, None, None, None, A.ptr(self.x.buf), A.ptr(self.cscale), A.ptr(self.att),
                                       A.ptr(self.y.buf), N, D * H * W, Cc, e.stream), "cbam apply " + self.name)
        e._count(5)

    def bwd(self):
        e = self.eng
        if not self.y.gflag or not self.x.needs_grad:
            return
        N, D, H, W, Cc = self.x.shape
        ch = self.mode in ("both", "channel")
        spt = self.mode in ("both", "spatial")
        acc = self.x.take_acc()
        g = lambda p, on: A.ptr(p.g) if on else A.ptr(self.sink)  # noqa: E731
        A.check(A.lib.sap3d_cbam_tail_bwd(
            e.dt, A.ptr(self.y.grad), A.ptr(self.ones_act), A.ptr(self.zero_act), A.ptr(self.zc), A.ptr(self.zg), A.ptr(self.zg),
            A.ptr(self.zc), A.ptr(self.x.buf), N, D, H, W, Cc, self.G, self.hidden,
            A.ptr(self.w0.w) if ch else A.ptr(self.zw), A.ptr(self.w1.w) if ch else A.ptr(self.zw),
            A.ptr(self.w_sp.w) if spt else A.ptr(self.zw), A.ptr(self.cscale), A.ptr(self.sp), A.ptr(self.att), A.ptr(self.save),
            None, 0, A.ptr(self.x.ensure_grad()), acc, A.ptr(self.sink), A.ptr(self.sink),
            g(self.w0, ch), g(self.b0, ch), g(self.w1, ch), g(self.b1, ch), A.ptr(self.w_sp.g) if spt else None,
            A.ptr(_ws(e)), e.stream), "cbam_bwd " + self.name)
        e._count(8)


class ConcatOp:
    """materialised tf.concat([a, b], -1) (only needed for three-way concatenations)"""

    def __init__(self, eng: Engine, a: T, b: T, name=""):
        self.eng, self.a, self.b = eng, a, b
        self.y = eng.tensor((*a.shape[:4], a.C + b.C), name)
        eng.fwd_ops.append(self.fwd)
        eng.bwd_ops.append(self.bwd)

    def bwd(self):
        e = self.eng
        if not self.y.gflag:
            return
        da = db = None
        acc_a = acc_b = 0
        if self.a.needs_grad:
            acc_a = self.a.take_acc()
            da = A.ptr(self.a.ensure_grad())
        if self.b.needs_grad:
            acc_b = self.b.take_acc()
            db = A.ptr(self.b.ensure_grad())
        A.check(A.lib.sap3d_split_channels(e.dt, A.ptr(self.y.grad), da, acc_a, db, acc_b, self.a.positions, self.a.C, self.b.C,
                                           e.stream), "split_channels")
        e._count()

    def fwd(self):
        e = self.eng
        A.check(A.lib.sap3d_concat_channels(e.dt, A.ptr(self.a.buf), A.ptr(self.b.buf), A.ptr(self.y.buf), self.a.positions,
                                            self.a.C, self.b.C, e.stream), "concat")
        e._count()
