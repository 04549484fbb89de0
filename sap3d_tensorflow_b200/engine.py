"""Static-schedule execution engine for the P3D saliency hot path.

The graph builders (p3d.py, gn/p3d_gn.py, network.py of this package — same public names as the
reference's) describe a model once against this engine; the engine owns the activation / gradient /
parameter buffers (torch tensors used purely as a device allocator) and a tape of launches into
libsap3d_b200.so.  forward()/backward()/adam run the tape on the current CUDA stream, so a whole
training step can be captured into ONE CUDA graph (no tracing compiler involved).

Replaces the TF-1.x graph executor walking the ~450 op kernels of one sess.run (train.py:217,
gen_pred.py:151).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _abi as A

BN_EPS = 1e-3        # tf.layers.batch_normalization defaults (TF-1.x)
BN_MOMENTUM = 0.99


def _dt(dtype: str):
    return (A.BF16, torch.bfloat16) if dtype == "bf16" else (A.F32, torch.float32)


class T:
    """activation tensor handle (NDHWC)"""

    __slots__ = ("shape", "buf", "grad", "name", "needs_grad", "gflag", "eng", "fork")

    def __init__(self, eng, shape, name="", needs_grad=None, torch_dtype=None):
        self.eng = eng
        self.shape = tuple(int(s) for s in shape)
        self.name = name
        self.buf = torch.empty(self.shape, device=eng.device, dtype=torch_dtype or eng.tdt)
        self.needs_grad = eng.training_graph if needs_grad is None else needs_grad
        self.grad = None
        self.gflag = False
        self.fork = None      # (ready event, gradient-complete event) of a cross-branch alias, see Engine.fork

    @classmethod
    def alias_of(cls, t: "T", name: str) -> "T":
        """a second handle on the same activation buffer with its OWN gradient buffer"""
        a = cls.__new__(cls)
        a.eng, a.shape, a.name, a.buf, a.needs_grad = t.eng, t.shape, name, t.buf, t.needs_grad
        a.grad, a.gflag, a.fork = None, False, None
        return a

    def ensure_grad(self):
        if self.grad is None:
            self.grad = torch.empty_like(self.buf)
        return self.grad

    @property
    def positions(self):
        return int(np.prod(self.shape[:-1]))

    @property
    def C(self):
        return self.shape[-1]

    def take_acc(self) -> int:
        """returns 1 if a previous consumer already wrote this tensor's gradient (then accumulate)"""
        acc = 1 if self.gflag else 0
        self.gflag = True
        return acc


class Param:
    __slots__ = ("name", "shape", "kind", "trainable", "numel", "offset", "w", "g", "fan")

    def __init__(self, name, shape, kind, trainable):
        self.name = name
        self.shape = tuple(int(s) for s in shape)
        self.kind = kind
        self.trainable = trainable
        self.numel = int(np.prod(self.shape))
        self.offset = -1
        self.w = None
        self.g = None


class NameScope:
    """TF-1.x default-name uniquification: conv3d, conv3d_1, ... per enclosing variable scope."""

    def __init__(self):
        self.counters: Dict[str, int] = {}

    def unique(self, scope: str, base: str) -> str:
        key = scope + "/" + base
        n = self.counters.get(key, 0)
        self.counters[key] = n + 1
        name = base if n == 0 else f"{base}_{n}"
        return (scope + "/" + name) if scope else name


class ConvOut:
    """raw (pre-normalisation) conv output plus the per-tile statistics its epilogue produced"""

    def __init__(self, raw: T, stats: Optional[torch.Tensor], rows: int, op=None):
        self.raw = raw
        self.stats = stats
        self.rows = rows
        self.op = op      # the producing _ConvOp (lets an inference-mode norm fold itself into the conv epilogue)


class NormState:
    def __init__(self, eng, C, gamma: Optional[Param], beta: Optional[Param], mm: Optional[Param], mv: Optional[Param]):
        self.gamma, self.beta, self.mm, self.mv = gamma, beta, mm, mv
        f = lambda: torch.empty(C, device=eng.device, dtype=torch.float32)  # noqa: E731
        self.scale, self.shift, self.mean, self.rstd = f(), f(), f(), f()


class _Tape(list):
    """op list of the engine; an op appended while a branch is selected (Engine.branch) runs on that branch's stream"""

    def __init__(self, eng):
        super().__init__()
        self._eng = eng

    def append(self, fn):
        b = self._eng._branch
        super().append(fn if b == 0 else self._eng._on_branch(fn, b))


class Engine:
    def __init__(self, dtype: str = "bf16", training_graph: bool = False, device: str = "cuda:0", conv_impl: int = A.IMPL_AUTO,
                 dropout_seed: int = 1234, per_sample_statistics: bool = False):
        if not torch.cuda.is_available():
            raise A.Sap3dError("sap3d_tensorflow_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.device = torch.device(device)
        self.dtype_name = dtype
        self.dt, self.tdt = _dt(dtype)
        self.training_graph = training_graph
        self.conv_impl = conv_impl
        self.names = NameScope()
        self.params: "OrderedDict[str, Param]" = OrderedDict()
        # ---- branches: independent sub-graphs on their own streams (p3d._unetpp: the decoder beside the backbone) ----------------
        # Ops recorded inside `with eng.branch(b)` run on branch_streams[b]; tensors cross between branches only through
        # fork() / consume_forked(), which carry the events.  OPT-IN (SAP3D_BRANCHES=1): measured on the B = 8 training step it
        # buys 0.12 ms of 17.15 (forward-only 5.72 -> 5.61 ms) -- the decoder's persistent kernels hold every SM while they run
        # (one 200 KB CTA per SM), so the backbone's small kernels wait for them instead of running beside them.
        self.branches_enabled = os.environ.get("SAP3D_BRANCHES", "0") == "1"
        self.branches_suspended = False
        self._branch = 0          # branch selected while BUILDING
        self._run_branch = 0      # branch whose op is RUNNING (selects the per-branch BatchNorm-backward workspace)
        self.branch_streams: Dict[int, torch.cuda.Stream] = {}
        self.fwd_ops: List = _Tape(self)
        self.bwd_ops: List = _Tape(self)
        self.tensors: List[T] = []
        self.convs: List = []
        self.taps: Dict[str, T] = {}
        self.finalized = False
        self.dropout_seed = dropout_seed
        self._dropout_ops = 0    # per-engine salt of the dropout masks (two engines with one seed draw the same masks)
        # inference graphs only: every batch-statistics BatchNorm takes its statistics PER CLIP, so a batch of B windows gives
        # what B single-window runs give (the reference's gen_pred.py feeds one window per sess.run while the backbone's
        # BatchNorm always uses batch statistics, p3d.py:140,350)
        if per_sample_statistics and training_graph:
            raise A.Sap3dError("per_sample_statistics is an inference-graph option")
        self.per_sample_bn = per_sample_statistics
        # BN moving averages follow the batch statistics only inside a training step (TF attaches UPDATE_OPS to train_op,
        # train.py:170-172; sess.run(pred) leaves them alone).  Forward-only execution passes momentum = 1: moving * 1 + batch * 0
        self.update_moving = False
        self.step = torch.zeros(1, device=self.device, dtype=torch.int32)
        self.loss_buf = torch.zeros(1, device=self.device, dtype=torch.float64)
        self._bwd_ws: Dict[int, torch.Tensor] = {}
        self.sync_bn = None      # parallel.SyncBatchNorm: batch statistics span the data-parallel replicas (eager mode)
        self._max_c = 8
        self.launches_fwd = 0
        self.launches_bwd = 0
        self._counting = None
        self._pack_table = None
        self.pre_pack_ops: List = []
        self.var_prefix = ""   # tf.variable_scope wrapped around a whole builder (gn/p3d_gn.py:490 'P3D')
        # data-parallel overlap: backward is cut at these marks (see mark_dp_split): (ops recorded so far, last parameter so far)
        self._split_marks: List[tuple] = []
        self.dp_segments: List[tuple] = []          # after finalize(): (ops_lo, ops_hi, grad_lo, grad_hi) in BACKWARD order
        self._split_ops: Optional[int] = None      # the LAST mark (the first cut backward meets); None = backward is not cut
        self.dp_split_offset: Optional[int] = None  # ... and the offset of the first gradient element behind it
        self.grad_source: Optional[torch.Tensor] = None   # bf16 copy of the gradients Adam should read instead of flat_g (DP exchange)
        # filter gradients run on side streams, off the critical path.  Several, taken round-robin: the ~200 small filter-gradient
        # launches of the backbone are independent of each other, and ONE stream serialised them into a 0.6 ms backlog that was still
        # draining after the main chain had finished (in-graph trace, r02).  SAP3D_WGRAD_STREAMS=1 restores the single stream.
        self.side_streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, int(os.environ.get("SAP3D_WGRAD_STREAMS", "3"))))]
        self.side_stream = self.side_streams[0]   # also the home of the padded-filter folds (stream order = dependency)
        self._wgrad_rr = 0
        self.aux_stream = torch.cuda.Stream(device=self.device, priority=-1)   # forward: the second of two independent convs
        self.use_side_stream = True

    # ------------------------------------------------------------------------------------------
    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    @property
    def bwd_ws(self) -> torch.Tensor:
        return self._bwd_ws[self._run_branch if self._run_branch in self._bwd_ws else 0]

    # ---- branches ------------------------------------------------------------------------------------------------------------
    def _branches_active(self) -> bool:
        # synchronised BatchNorm issues NCCL collectives from inside the ops: keep every rank's issue order identical.
        # branches_suspended: the backward pass is being captured as several CUDA graphs (overlapped data-parallel exchange);
        # an event recorded in one capture cannot be waited for in another, so everything runs on the main stream then.
        return self.branches_enabled and self.sync_bn is None and not self.branches_suspended

    def _on_branch(self, fn, b):
        def run():
            if not self._branches_active():
                return fn()
            prev = self._run_branch
            self._run_branch = b
            try:
                with torch.cuda.stream(self.branch_streams[b]):
                    fn()
            finally:
                self._run_branch = prev
        return run

    def branch(self, b: int):
        """context manager: ops recorded inside run on branch b's stream (0 = the main chain)"""
        eng = self

        class _Ctx:
            def __enter__(self_c):
                self_c.prev = eng._branch
                if eng.branches_enabled:
                    assert not eng.finalized
                    if b != 0 and b not in eng.branch_streams:
                        eng.branch_streams[b] = torch.cuda.Stream(device=eng.device, priority=-1)
                    eng._branch = b

            def __exit__(self_c, *exc):
                eng._branch = self_c.prev
                return False
        return _Ctx()

    def fork(self, t: T, name: str = "") -> T:
        """a handle on `t` for ONE consumer branch other than the current one: same activation buffer, own gradient buffer.
        Forward: records `ready` on the producer's stream.  Backward (runs on the producer's branch right before the producer's
        own backward): waits for the consumer branch's `grad done` event and adds the alias gradient into t's gradient."""
        if not self.branches_enabled:
            return t
        t2 = T.alias_of(t, name or (t.name + "/fork"))
        self.tensors.append(t2)
        ready, done = torch.cuda.Event(), torch.cuda.Event()
        t2.fork = (ready, done)
        dev = self.device

        def fwd():
            if self._branches_active():
                ready.record(torch.cuda.current_stream(dev))

        def bwd():
            if not t2.gflag or not t.needs_grad:
                return
            if self._branches_active():
                torch.cuda.current_stream(dev).wait_event(done)
            g = t.ensure_grad()
            if t.take_acc():
                g.add_(t2.grad)
            else:
                g.copy_(t2.grad)
            self._count()
        self.fwd_ops.append(fwd)
        self.bwd_ops.append(bwd)
        return t2

    def consume_forked(self, ts: Sequence[T]):
        """called on the consumer branch BEFORE its first op that reads the forked tensors `ts`: forward waits for their `ready`
        events; backward (reverse order: after every later op of this branch) records their `grad done` events."""
        ts = [t for t in ts if t.fork is not None]
        if not ts:
            return
        dev = self.device

        def fwd():
            if self._branches_active():
                cur = torch.cuda.current_stream(dev)
                for t in ts:
                    cur.wait_event(t.fork[0])

        def bwd():
            if self._branches_active():
                cur = torch.cuda.current_stream(dev)
                for t in ts:
                    t.fork[1].record(cur)
        self.fwd_ops.append(fwd)
        self.bwd_ops.append(bwd)

    def _fork_branches(self):
        if self._branches_active():
            main = torch.cuda.current_stream(self.device)
            for s in self.branch_streams.values():
                s.wait_stream(main)

    def _join_branches(self):
        if self._branches_active():
            main = torch.cuda.current_stream(self.device)
            for s in self.branch_streams.values():
                main.wait_stream(s)

    def tensor(self, shape, name="", needs_grad=None, torch_dtype=None) -> T:
        t = T(self, shape, name, needs_grad, torch_dtype)
        self.tensors.append(t)
        return t

    def tap(self, name: str, t: T) -> T:
        self.taps[name] = t
        return t

    def param(self, name, shape, kind, trainable=True) -> Param:
        name = self.var_prefix + name
        if name in self.params:
            p = self.params[name]
            assert p.shape == tuple(shape), (name, p.shape, shape)
            return p
        assert not self.finalized
        p = Param(name, shape, kind, trainable)
        self.params[name] = p
        return p

    def mark_dp_split(self):
        """called by the builders between the stages of the backbone (and before the decoder): everything created AFTER a mark
        finishes its backward BEFORE everything created ahead of it and lies behind it in the flat gradient buffer, so the
        data-parallel exchange of a finished tail segment runs while the gradients of the earlier layers are still being
        computed.  SAP3D_DP_MARKS (digits, default all) selects which of the builder's marks are kept: "2" = only the third."""
        idx = getattr(self, "_mark_calls", 0)
        self._mark_calls = idx + 1
        keep = os.environ.get("SAP3D_DP_MARKS")
        if keep is not None and str(idx) not in keep:
            return
        self._split_marks.append((len(self.bwd_ops), next(reversed(self.params)) if self.params else None))

    def _count(self, n=1):
        if self._counting == "fwd":
            self.launches_fwd += n
        elif self._counting == "bwd":
            self.launches_bwd += n

    # ------------------------------------------------------------------------------------------
    # parameter storage: ONE flat fp32 buffer each for weights / grads / adam m / adam v
    # ------------------------------------------------------------------------------------------
    def finalize(self):
        assert not self.finalized
        off = 0
        ordered = [p for p in self.params.values() if p.trainable] + [p for p in self.params.values() if not p.trainable]
        n_train = 0
        for p in ordered:
            p.offset = off
            off += (p.numel + 63) // 64 * 64
            if p.trainable:
                n_train = off
        self.n_train = n_train
        if self._split_marks:
            created = list(self.params)
            cuts = []   # (ops index, gradient offset), ascending; marks with no trainable parameter on one side are dropped
            for n_ops, last in self._split_marks:
                k = created.index(last) + 1 if last in created else 0
                later = [n for n in created[k:] if self.params[n].trainable]
                o = self.params[later[0]].offset if later else None
                if o is None or o == 0 or n_ops == 0 or (cuts and (o <= cuts[-1][1] or n_ops <= cuts[-1][0])):
                    continue
                cuts.append((n_ops, o))
            if cuts:
                ops_b = [0] + [c[0] for c in cuts] + [None]
                g_b = [0] + [c[1] for c in cuts] + [n_train]
                self.dp_segments = [(ops_b[i], ops_b[i + 1], g_b[i], g_b[i + 1]) for i in reversed(range(len(cuts) + 1))]
                self._split_ops, self.dp_split_offset = cuts[-1]
        self.flat_w = torch.zeros(off, device=self.device, dtype=torch.float32)
        self.flat_g = torch.zeros(max(n_train, 1), device=self.device, dtype=torch.float32) if self.training_graph else None
        self.flat_m = torch.zeros(max(n_train, 1), device=self.device, dtype=torch.float32) if self.training_graph else None
        self.flat_v = torch.zeros(max(n_train, 1), device=self.device, dtype=torch.float32) if self.training_graph else None
        for p in ordered:
            p.w = self.flat_w[p.offset:p.offset + p.numel].view(p.shape)
            if p.trainable and self.training_graph:
                p.g = self.flat_g[p.offset:p.offset + p.numel].view(p.shape)
        if self.training_graph:
            nbytes = A.lib.sap3d_affine_act_bwd_workspace(self._max_c)
            # one BatchNorm-backward workspace per branch: ops of different branches run at the same time
            self._bwd_ws = {b: torch.zeros(nbytes // 4 + 16, device=self.device, dtype=torch.float32) for b in [0] + sorted(self.branch_streams)}
        self.finalized = True
        self.init_params_tf(0)

    def init_params_tf(self, seed: int = 0):
        """TensorFlow's default initialisers for every variable kind (p3d.py:12; tf.layers defaults)."""
        rng = np.random.RandomState(seed)
        for p in self.params.values():
            shp = p.shape
            if p.kind in ("glorot", "glorot_t"):
                rf = int(np.prod(shp[:-2])) if len(shp) > 2 else 1
                lim = math.sqrt(6.0 / ((shp[-1] + shp[-2]) * rf))
                v = rng.uniform(-lim, lim, size=shp)
            elif p.kind == "xavier1d":  # get_conv_weight(name+'_bias',[C],0): xavier on a rank-1 shape
                lim = math.sqrt(6.0 / (2 * shp[0]))
                v = rng.uniform(-lim, lim, size=shp)
            elif p.kind == "vscale":  # variance_scaling_initializer(): truncated normal, factor 2, FAN_IN
                rf = int(np.prod(shp[:-2])) if len(shp) > 2 else 1
                std = math.sqrt(2.0 / (shp[-2] * rf)) / 0.8796
                v = np.clip(rng.normal(0, std, size=shp), -2 * std, 2 * std)
            elif p.kind in ("ones", "var"):
                v = np.ones(shp)
            elif p.kind in ("zeros", "mean", "bias", "sa_gamma"):
                v = np.zeros(shp)
            else:
                raise ValueError(p.kind)
            p.w.copy_(torch.tensor(v, dtype=torch.float32))
        self.pack_weights()

    def load_params(self, values: Dict[str, "np.ndarray | torch.Tensor"], strict: bool = True):
        for name, p in self.params.items():
            if name not in values:
                if strict:
                    raise KeyError(f"missing variable {name}")
                continue
            v = values[name]
            v = v if isinstance(v, torch.Tensor) else torch.tensor(np.asarray(v))
            p.w.copy_(v.to(torch.float32).reshape(p.shape))
        extra = set(values) - set(self.params)
        if strict and extra:
            raise KeyError(f"unknown variables {sorted(extra)[:5]} ...")
        self.pack_weights()

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {n: p.w.detach().cpu().clone() for n, p in self.params.items()}

    def pack_weights(self):
        """fp32 master filters -> bf16 tensor-core operands for every conv, in ONE kernel launch"""
        for f in self.pre_pack_ops:   # derived parameters (zero-padded filters) must be current before packing
            f()
        if self.dt != A.BF16 or not self.convs:
            return
        if self._pack_table is None:
            ents = []
            start = 0
            for c in self.convs:
                two = (A.PackEntry * 2)()
                n = A.lib.sap3d_conv_pack_entries(C.byref(c.desc), A.ptr(c.w.w), A.ptr(c.wf), A.ptr(c.wd), two)
                for i in range(n):
                    e = two[i]
                    e.start = start      # index space only: starts (and the total) are multiples of the kernel's 4096-element span
                    start += (e.rows_pad * e.taps * e.cols + 4095) // 4096 * 4096
                    ents.append(bytes(e))
            raw = b"".join(ents)
            self._pack_n, self._pack_total = len(ents), start
            # (a graph of im2col-form convs only -- the stem-only frame-cache graph -- has nothing to pre-pack)
            self._pack_table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device) if ents else False
        if self._pack_table is False:
            return
        A.check(A.lib.sap3d_pack_multi(A.ptr(self._pack_table), self._pack_n, self._pack_total, self.stream), "pack_multi")

    # ------------------------------------------------------------------------------------------
    # ops
    # ------------------------------------------------------------------------------------------
    def conv(self, xs: Sequence[T], cout: int, kernel, strides, w: Param, b: Optional[Param] = None, transposed=False,
             want_stats=True, name="", out_f32=False, bias_grad=True) -> ConvOut:
        """bias_grad=False: the bias feeds a batch-statistics norm, whose mean subtraction makes dL/dbias exactly
        zero in exact arithmetic (TF computes rounding noise there); the gradient is left at zero."""
        op = _ConvOp(self, list(xs), cout, tuple(kernel), tuple(strides), w, b, transposed, want_stats, name, out_f32)
        op.bias_grad = bias_grad
        self.convs.append(op)
        self.fwd_ops.append(op.fwd)
        self.bwd_ops.append(op.bwd)
        return op.out

    def conv_padded_cout(self, xs: Sequence[T], cout_pad: int, kernel, strides, w: Param, b: Optional[Param], name="") -> ConvOut:
        """conv whose TF variable has cout = w.shape[-1] output channels but which is EXECUTED with cout_pad channels
        (zero filters / biases for the padding): keeps 16- and 32-channel projections on the tensor-core path in forward,
        data-gradient and filter-gradient.  Output channels >= cout are exactly zero."""
        op = _PaddedParams(self, w, b, cout_pad)
        self.bwd_ops.append(op.bwd)   # registered BEFORE the conv: the bwd list is walked in reverse, so it runs right after it
        return self.conv(xs, cout_pad, kernel, strides, op.wp, op.bp, transposed=False, want_stats=False, name=name)

    def norm_state(self, C, gamma=None, beta=None, mm=None, mv=None) -> NormState:
        self._max_c = max(self._max_c, C)
        return NormState(self, C, gamma, beta, mm, mv)

    def norm_act(self, a: ConvOut, n1: Optional[NormState], train1: bool, relu1: bool, b=None, n2: Optional[NormState] = None,
                 train2: bool = False, relu2: bool = False, relu_out: bool = False, name="") -> T:
        op = _NormActOp(self, a, n1, train1, relu1, b, n2, train2, relu2, relu_out, name)
        self.fwd_ops.append(op.fwd)
        self.bwd_ops.append(op.bwd)
        return op.y

    def maxpool(self, x: T, ksize, strides, same=True, name="") -> T:
        op = _PoolOp(self, x, ksize, strides, same, name)
        self.fwd_ops.append(op.fwd)
        self.bwd_ops.append(op.bwd)
        return op.y

    def dropout(self, x: T, rate: float, name="") -> T:
        if rate <= 0.0:
            return x
        op = _DropoutOp(self, x, rate, name)
        self.fwd_ops.append(op.fwd)
        self.bwd_ops.append(op.bwd)
        return op.y

    def attn_core(self, g: T, f: T, h: T, name="") -> T:
        op = _AttnCoreOp(self, g, f, h, name)
        self.fwd_ops.append(op.fwd)
        self.bwd_ops.append(op.bwd)
        return op.o

    def gate(self, o: T, x: T, gamma: Param, name="") -> T:
        op = _GateOp(self, o, x, gamma, name)
        self.fwd_ops.append(op.fwd)
        self.bwd_ops.append(op.bwd)
        return op.y

    def head(self, x: T, w: Param, b: Param, ksize, stride, sigmoid=True, name="") -> "_HeadOp":
        op = _HeadOp(self, x, w, b, ksize, stride, sigmoid, name)
        self.fwd_ops.append(op.fwd)
        self.bwd_ops.append(op.bwd)
        return op

    def logits_loss(self, co: ConvOut, name="") -> "_LogitsLossOp":
        """network output produced by an ordinary convolution (fp32 logits) + the smooth-L1 loss in training graphs"""
        op = _LogitsLossOp(self, co, name)
        self.bwd_ops.append(op.bwd)
        return op

    # ------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------
    def forward(self):
        self._counting = "fwd"
        self.launches_fwd = 0
        self._fork_branches()
        for f in self.fwd_ops:
            f()
        self._join_branches()
        self._counting = None

    def backward(self, part: Optional[int] = None):
        """part None: the whole backward pass; k: segment k of dp_segments (0 = the ops created after the last mark_dp_split --
        head, decoder, last stage --, ..., the last one = stem + first stage).  Each part ends by joining the filter-gradient side
        stream, so its share of the flat gradient buffer is complete."""
        assert self.training_graph
        self._counting = "bwd"
        if part is not None and not self.dp_segments:
            part = None
        if part in (None, 0):
            self.launches_bwd = 0
            for t in self.tensors:
                t.gflag = False
            self.flat_g.zero_()
            self._count()
        if part is None:
            ops = self.bwd_ops
        else:
            lo, hi = self.dp_segments[part][:2]
            ops = self.bwd_ops[lo:hi]
        self._fork_branches()
        for f in reversed(ops):
            f()
        self._join_branches()
        if self.use_side_stream:
            for ss in self.side_streams:
                torch.cuda.current_stream(self.device).wait_stream(ss)   # join the filter-gradient branches
        self._counting = None

    def adam(self, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8, grad_scale=1.0):
        if self.grad_source is not None:     # all-reduced bf16 buckets of the data-parallel exchange, read in place
            A.check(A.lib.sap3d_adam_step_g(A.ptr(self.flat_w), A.ptr(self.grad_source), A.BF16, A.ptr(self.flat_m), A.ptr(self.flat_v),
                                            self.n_train, A.ptr(self.step), lr, b1, b2, eps, grad_scale, self.stream), "adam")
        else:
            A.check(A.lib.sap3d_adam_step(A.ptr(self.flat_w), A.ptr(self.flat_g), A.ptr(self.flat_m), A.ptr(self.flat_v),
                                          self.n_train, A.ptr(self.step), lr, b1, b2, eps, grad_scale, self.stream), "adam")
        self.pack_weights()

    def begin_step(self):
        A.check(A.lib.sap3d_step_increment(A.ptr(self.step), self.stream), "step_increment")
        self.loss_buf.zero_()


# ==================================================================================================
class _ConvOp:
    def __init__(self, eng: Engine, xs, cout, kernel, strides, w: Param, b, transposed, want_stats, name, out_f32):
        self.eng = eng
        self.xs = xs
        self.w, self.b = w, b
        self.name = name
        N, D, H, W, _ = xs[0].shape
        cin = [x.C for x in xs]
        self.desc = A.make_conv_desc(eng.dt, N, D, H, W, cin, cout, kernel, strides, transposed, b is not None, out_f32,
                                     eng.conv_impl)
        Do, Ho, Wo = A.conv_out_dims(self.desc)
        raw = eng.tensor((N, Do, Ho, Wo, cout), name + "/raw", torch_dtype=torch.float32 if out_f32 else None)
        rows = A.lib.sap3d_conv_stats_rows(C.byref(self.desc)) if want_stats else 0
        stats = torch.zeros(rows, 2, cout, device=eng.device, dtype=torch.float32) if want_stats else None
        self.out = ConvOut(raw, stats, rows, self)
        self.fused_affine = None   # (NormState, relu): inference-mode BN (+ReLU) applied in the conv epilogue
        self.fused_bn = None       # _NormActOp: training-mode BatchNorm finished inside this conv's launch (sap3d_conv_fwd_bn)
        self.aux = False           # forward launch on the aux stream (independent sibling branch, joined by its consumer)
        self.use_tc = eng.dt == A.BF16
        nf = A.lib.sap3d_conv_packed_elems(C.byref(self.desc), 0)
        nd = A.lib.sap3d_conv_packed_elems(C.byref(self.desc), 1)
        self.wf = torch.zeros(nf, device=eng.device, dtype=torch.bfloat16) if self.use_tc else None
        self.wd = torch.zeros(nd, device=eng.device, dtype=torch.bfloat16) if (self.use_tc and eng.training_graph) else None

    def pack(self):
        if self.use_tc:
            A.check(A.lib.sap3d_conv_pack_weights(C.byref(self.desc), A.ptr(self.w.w), A.ptr(self.wf), A.ptr(self.wd),
                                                  self.eng.stream), "pack " + self.name)

    def fwd(self):
        e = self.eng
        if self.aux and e.use_side_stream:
            # fork: this conv and the next op on the main stream both only depend on what has been enqueued so far; the
            # few-CTA backbone convs of parallel branches (ST_B's S and T, the projection shortcut) then share the GPU
            main = torch.cuda.current_stream(e.device)
            e.aux_stream.wait_stream(main)
            with torch.cuda.stream(e.aux_stream):
                self._fwd()
            return
        self._fwd()

    def _fwd(self):
        e = self.eng
        x1 = self.xs[1].buf if len(self.xs) > 1 else None
        if self.fused_affine is not None:
            ns, relu = self.fused_affine
            A.check(A.lib.sap3d_bn_finalize(None, 0, self.out.raw.C, 1.0, A.ptr(ns.gamma.w), A.ptr(ns.beta.w), A.ptr(ns.mm.w),
                                            A.ptr(ns.mv.w), 0, BN_MOMENTUM, BN_EPS, A.ptr(ns.scale), A.ptr(ns.shift), A.ptr(ns.mean),
                                            A.ptr(ns.rstd), e.stream), "bn_finalize(moving) " + self.name)
            A.check(A.lib.sap3d_conv_fwd_affine(C.byref(self.desc), A.ptr(self.xs[0].buf), A.ptr(x1), A.ptr(self.w.w), A.ptr(self.wf),
                                                A.ptr(self.b.w) if self.b is not None else None, A.ptr(ns.scale), A.ptr(ns.shift),
                                                int(relu), A.ptr(self.out.raw.buf), e.stream), "conv_fwd_affine " + self.name)
            e._count(2)
            return
        nop = self.fused_bn
        if nop is not None and nop.fuse_active():
            ns = nop.n1
            f = A.BnFuse(A.ptr(ns.gamma.w), A.ptr(ns.beta.w), A.ptr(ns.mm.w), A.ptr(ns.mv.w), BN_MOMENTUM if e.update_moving else 1.0, BN_EPS,
                         A.ptr(ns.scale), A.ptr(ns.shift), A.ptr(ns.mean), A.ptr(ns.rstd), int(nop.relu1),
                         A.ptr(nop.b_t.buf) if nop.b_t is not None else None, int(nop.relu_out), A.ptr(nop.y.buf))
            A.check(A.lib.sap3d_conv_fwd_bn(C.byref(self.desc), A.ptr(self.xs[0].buf), A.ptr(x1), A.ptr(self.w.w), A.ptr(self.wf),
                                            A.ptr(self.b.w) if self.b is not None else None, A.ptr(self.out.raw.buf),
                                            A.ptr(self.out.stats), C.byref(f), e.stream), "conv_fwd_bn " + self.name)
            e._count()
            return
        A.check(A.lib.sap3d_conv_fwd(C.byref(self.desc), A.ptr(self.xs[0].buf), A.ptr(x1), A.ptr(self.w.w), A.ptr(self.wf),
                                     A.ptr(self.b.w) if self.b is not None else None, A.ptr(self.out.raw.buf),
                                     A.ptr(self.out.stats), e.stream), "conv_fwd " + self.name)
        e._count()

    def bwd(self):
        e = self.eng
        raw = self.out.raw
        if not raw.gflag:
            return
        dy = raw.grad
        if (len(self.xs) == 2 and all(x.needs_grad for x in self.xs) and self.xs[0] is not self.xs[1]
                and A.lib.sap3d_conv_dgrad2_supported(C.byref(self.desc)) == 1):
            # fused-concat conv with two equal segments: both data gradients from ONE launch (dy read once per tap)
            a0, a1 = self.xs[0].take_acc(), self.xs[1].take_acc()
            A.check(A.lib.sap3d_conv_dgrad2(C.byref(self.desc), A.ptr(dy), A.ptr(self.w.w), A.ptr(self.wd), A.ptr(self.xs[0].ensure_grad()),
                                            a0, A.ptr(self.xs[1].ensure_grad()), a1, e.stream), "conv_dgrad2 " + self.name)
            e._count()
        else:
            for si, x in enumerate(self.xs):
                if not x.needs_grad:
                    continue
                acc = x.take_acc()
                A.check(A.lib.sap3d_conv_dgrad(C.byref(self.desc), si, A.ptr(dy), A.ptr(self.w.w), A.ptr(self.wd),
                                               A.ptr(x.ensure_grad()), acc, e.stream), "conv_dgrad " + self.name)
                e._count()
        x1 = self.xs[1].buf if len(self.xs) > 1 else None
        main = torch.cuda.current_stream(e.device)
        if e.use_side_stream:
            # fork: the filter gradient only needs dy (ready on `main`) and the saved inputs; it is joined before Adam.  Filters that
            # are folded afterwards (zero-padded attention projections, _PaddedFilter.bwd on side_streams[0]) stay on that stream.
            ss = e.side_streams[0] if isinstance(self.w, _TempParam) else e.side_streams[e._wgrad_rr % len(e.side_streams)]
            e._wgrad_rr += 1
            ss.wait_stream(main)
            st = ss.cuda_stream
        else:
            st = e.stream
        A.check(A.lib.sap3d_conv_wgrad(C.byref(self.desc), A.ptr(self.xs[0].buf), A.ptr(x1), A.ptr(dy), A.ptr(self.w.g),
                                       A.ptr(self.b.g) if (self.b is not None and self.bias_grad) else None, A.ptr(self.wf), st),
                "conv_wgrad " + self.name)
        e._count(len(self.xs) + (1 if (self.b is not None and self.bias_grad) else 0))


class _NormActOp:
    """y = relu_out?( relu1?(norm1(a)) + relu2?(norm2(b) | b) )"""

    def __init__(self, eng, a: ConvOut, n1, train1, relu1, b, n2, train2, relu2, relu_out, name):
        self.eng = eng
        self.a, self.n1, self.train1, self.relu1 = a, n1, train1, relu1
        self.b, self.n2, self.train2, self.relu2 = b, n2, train2, relu2
        self.relu_out = relu_out
        self.name = name
        self.b_t: Optional[T] = None if b is None else (b.raw if isinstance(b, ConvOut) else b)
        # inference graphs: a moving-statistics BatchNorm (+ReLU) with no second operand folds into the conv epilogue
        # (the raw tensor then IS the normalised output; nothing is launched here)
        self.folded = (not eng.training_graph and n1 is not None and not train1 and b is None and n2 is None and not relu_out
                       and a.op is not None and a.op.fused_affine is None and not a.raw.shape[-1] % 8
                       and all(t is not a.raw for t in eng.taps.values())
                       and A.lib.sap3d_conv_fwd_on_tensor_cores(C.byref(a.op.desc)) == 1 and not a.op.desc.out_f32)
        if self.folded:
            a.op.fused_affine = (n1, relu1)
            self.y = a.raw
            return
        self.y = eng.tensor(a.raw.shape, name)
        if eng.training_graph:
            a.raw.ensure_grad()
        # training graphs, on request (SAP3D_CONV_FUSE_BN=1; measured slower than the two launches, see conv_tc.cuh): a
        # batch-statistics BatchNorm (+ReLU, + plain residual) whose convolution keeps every CTA resident (backbone stages 2-3) is
        # finished inside that convolution's launch; nothing is launched here then
        self.fused_conv = None
        if (eng.training_graph and n1 is not None and train1 and n2 is None and not relu2 and not isinstance(b, ConvOut)
                and a.op is not None and a.op.use_tc and not a.op.aux and a.op.fused_bn is None and a.op.fused_affine is None
                and eng.dt == A.BF16 and a.stats is not None and os.environ.get("SAP3D_CONV_FUSE_BN", "0") == "1"
                and (self.b_t is None or (self.b_t.shape == a.raw.shape and self.b_t is not a.raw))
                and A.lib.sap3d_conv_fwd_bn_supported(C.byref(a.op.desc)) == 1):
            a.op.fused_bn = self
            self.fused_conv = a.op

    def fuse_active(self) -> bool:
        """the producing convolution finishes this norm (decided per run: synchronised / per-sample statistics need the
        separate launches)"""
        e = self.eng
        return self.fused_conv is not None and e.sync_bn is None and not e.per_sample_bn

    def _finalize(self, co: ConvOut, ns: NormState, training: bool):
        e = self.eng
        cnt = float(co.raw.positions)
        if training and e.sync_bn is not None:
            # synchronised BatchNorm: the per-tile (sum, sum of squares) rows are summed over the replicas, then finalised
            # against the global position count -- the statistics of the reference's single-device batch
            e.sync_bn.all_reduce(co.stats)
            cnt *= e.sync_bn.world
        A.check(A.lib.sap3d_bn_finalize(A.ptr(co.stats), co.rows, co.raw.C, cnt, A.ptr(ns.gamma.w), A.ptr(ns.beta.w),
                                        A.ptr(ns.mm.w), A.ptr(ns.mv.w), int(training), BN_MOMENTUM if e.update_moving else 1.0, BN_EPS,
                                        A.ptr(ns.scale), A.ptr(ns.shift), A.ptr(ns.mean), A.ptr(ns.rstd), e.stream), "bn_finalize " + self.name)
        e._count()

    FUSE_MAX_ROWS = 128   # statistics rows up to which finalize is folded into the apply launch (whole backbone)

    def _fwd_fused(self) -> bool:
        n1, n2, a, e = self.n1, self.n2, self.a, self.eng
        if e.sync_bn is not None and (self.train1 or self.train2):
            return False
        if n1 is None or (self.train1 and a.rows > self.FUSE_MAX_ROWS):
            return False
        b_co = self.b if isinstance(self.b, ConvOut) else None
        if n2 is not None and (b_co is None or (self.train2 and b_co.rows > self.FUSE_MAX_ROWS)):
            return False
        A.check(A.lib.sap3d_bn_apply_fused(
            e.dt, A.ptr(a.raw.buf), A.ptr(a.stats), a.rows, A.ptr(n1.gamma.w), A.ptr(n1.beta.w), A.ptr(n1.mm.w), A.ptr(n1.mv.w),
            int(self.train1), A.ptr(n1.scale), A.ptr(n1.shift), A.ptr(n1.mean), A.ptr(n1.rstd), int(self.relu1),
            A.ptr(self.b_t.buf) if self.b_t is not None else None, int(n2 is not None),
            A.ptr(b_co.stats) if n2 else None, b_co.rows if n2 else 0, A.ptr(n2.gamma.w) if n2 else None,
            A.ptr(n2.beta.w) if n2 else None, A.ptr(n2.mm.w) if n2 else None, A.ptr(n2.mv.w) if n2 else None, int(self.train2),
            A.ptr(n2.scale) if n2 else None, A.ptr(n2.shift) if n2 else None, A.ptr(n2.mean) if n2 else None,
            A.ptr(n2.rstd) if n2 else None, int(self.relu2), int(self.relu_out), A.ptr(self.y.buf), self.y.positions, self.y.C,
            float(a.raw.positions), BN_MOMENTUM if e.update_moving else 1.0, BN_EPS, e.stream), "bn_apply_fused " + self.name)
        e._count()
        return True

    def _per_sample_affine(self, t: T, ns: NormState, slot: int):
        """per-clip batch statistics: scale / shift [N][C] from the stored raw tensor -- GroupNorm's kernels with one channel per
        group and BatchNorm's epsilon (moving averages are not touched: this is an inference-only mode)"""
        e = self.eng
        N, S, Cc = t.shape[0], t.positions // t.shape[0], t.C
        if not hasattr(self, "_ps"):
            self._ps = {}
        if slot not in self._ps:
            rows = A.lib.sap3d_sample_stats_rows(S, Cc, N)
            f = lambda *shape: torch.empty(*shape, device=e.device, dtype=torch.float32)  # noqa: E731
            self._ps[slot] = (rows, f(N, rows, 3, Cc), f(N, Cc), f(N, Cc), f(N, Cc), f(N, Cc))
        rows, part, scale, shift, mean, rstd = self._ps[slot]
        A.check(A.lib.sap3d_sample_channel_partials(e.dt, A.ptr(t.buf), None, N, S, Cc, rows, A.ptr(part), e.stream), "bn per-sample partials")
        A.check(A.lib.sap3d_gn_finalize(A.ptr(part), rows, N, S, Cc, Cc, A.ptr(ns.gamma.w), A.ptr(ns.beta.w), BN_EPS, A.ptr(scale), A.ptr(shift),
                                        A.ptr(mean), A.ptr(rstd), e.stream), "bn per-sample finalize")
        e._count(2)
        return scale, shift

    def _fwd_per_sample(self):
        e, n1, n2 = self.eng, self.n1, self.n2
        y = self.y
        N, S = y.shape[0], y.positions // y.shape[0]
        # one launch when every norm involved takes per-clip batch statistics and the (clip, 64-channel) slab is small -- the
        # whole backbone; SAP3D_SAMPLE_NORM_FUSED=0 restores partials + finalize + apply
        if (n1 is not None and self.train1 and (n2 is None or self.train2) and os.environ.get("SAP3D_SAMPLE_NORM_FUSED", "1") != "0"
                and A.lib.sap3d_sample_norm_apply_supported(e.dt, S, y.C) == 1):
            A.check(A.lib.sap3d_sample_norm_apply(
                e.dt, A.ptr(self.a.raw.buf), A.ptr(n1.gamma.w), A.ptr(n1.beta.w), int(self.relu1),
                A.ptr(self.b_t.buf) if self.b_t is not None else None, A.ptr(n2.gamma.w) if n2 else None, A.ptr(n2.beta.w) if n2 else None,
                int(self.relu2), int(self.relu_out), A.ptr(y.buf), N, S, y.C, BN_EPS, e.stream), "sample_norm_apply " + self.name)
            e._count()
            return
        s1 = t1 = s2 = t2 = None
        if n1 is not None:
            if self.train1:
                s1, t1 = self._per_sample_affine(self.a.raw, n1, 0)
            else:   # moving statistics: the same [C] vector for every clip, broadcast to [N][C]
                self._finalize(self.a, n1, False)
                N = self.a.raw.shape[0]
                s1, t1 = n1.scale.repeat(N).contiguous(), n1.shift.repeat(N).contiguous()
        if n2 is not None:
            if self.train2:
                s2, t2 = self._per_sample_affine(self.b_t, n2, 1)
            else:
                self._finalize(self.b, n2, False)
                N = self.b_t.shape[0]
                s2, t2 = n2.scale.repeat(N).contiguous(), n2.shift.repeat(N).contiguous()
        y = self.y
        A.check(A.lib.sap3d_affine_act(e.dt, A.ptr(self.a.raw.buf), A.ptr(s1), A.ptr(t1), int(self.relu1),
                                       A.ptr(self.b_t.buf) if self.b_t is not None else None, A.ptr(s2), A.ptr(t2), int(self.relu2),
                                       int(self.relu_out), A.ptr(y.buf), y.positions, y.C, y.positions // y.shape[0], e.stream),
                "affine_act (per-sample statistics) " + self.name)
        e._count()
        self._keep = (s1, t1, s2, t2)

    def fwd(self):
        e = self.eng
        if self.folded or self.fuse_active():
            return
        if isinstance(self.b, ConvOut) and self.b.op is not None and self.b.op.aux and e.use_side_stream:
            torch.cuda.current_stream(e.device).wait_stream(e.aux_stream)   # join the sibling branch
        if e.per_sample_bn and ((self.n1 is not None and self.train1) or (self.n2 is not None and self.train2)):
            return self._fwd_per_sample()
        if self._fwd_fused():
            return
        if self.n1 is not None:
            self._finalize(self.a, self.n1, self.train1)
        if self.n2 is not None:
            self._finalize(self.b, self.n2, self.train2)
        n1, n2 = self.n1, self.n2
        A.check(A.lib.sap3d_affine_act(e.dt, A.ptr(self.a.raw.buf), A.ptr(n1.scale) if n1 else None,
                                       A.ptr(n1.shift) if n1 else None, int(self.relu1),
                                       A.ptr(self.b_t.buf) if self.b_t is not None else None,
                                       A.ptr(n2.scale) if n2 else None, A.ptr(n2.shift) if n2 else None, int(self.relu2),
                                       int(self.relu_out), A.ptr(self.y.buf), self.y.positions, self.y.C, 0, e.stream),
                "affine_act " + self.name)
        e._count()

    def bwd(self):
        e = self.eng
        if self.folded or not self.y.gflag:
            return
        n1, n2 = self.n1, self.n2
        a_raw = self.a.raw
        acc_a = a_raw.take_acc()
        db_ptr, acc_b = None, 0
        if self.b_t is not None and self.b_t.needs_grad:
            acc_b = self.b_t.take_acc()
            db_ptr = A.ptr(self.b_t.ensure_grad())
        bs1 = n1 is not None and self.train1
        bs2 = n2 is not None and self.train2
        args = (
            e.dt, A.ptr(self.y.grad), A.ptr(a_raw.buf),
            A.ptr(n1.scale) if n1 else None, A.ptr(n1.shift) if n1 else None,
            A.ptr(n1.mean) if bs1 else None, A.ptr(n1.rstd) if bs1 else None, int(self.relu1),
            A.ptr(self.b_t.buf) if self.b_t is not None else None,
            A.ptr(n2.scale) if n2 else None, A.ptr(n2.shift) if n2 else None,
            A.ptr(n2.mean) if bs2 else None, A.ptr(n2.rstd) if bs2 else None, int(self.relu2),
            int(self.relu_out), self.y.positions, self.y.C,
            A.ptr(a_raw.ensure_grad()), acc_a, db_ptr, acc_b,
            A.ptr(n1.gamma.g) if n1 else None, A.ptr(n1.beta.g) if n1 else None,
            A.ptr(n2.gamma.g) if n2 else None, A.ptr(n2.beta.g) if n2 else None,
            A.ptr(e.bwd_ws), e.stream)
        if e.sync_bn is not None and (bs1 or bs2):
            # the per-channel sums of the BN backward span the replicas too; d(gamma), d(beta) stay LOCAL sums (the gradient
            # exchange adds them up with every other gradient)
            cnt = float(self.y.positions) * e.sync_bn.world
            A.check(A.lib.sap3d_affine_act_bwd_sync(*args, cnt, 1), "affine_act_bwd reduce " + self.name)
            e.sync_bn.all_reduce(e.bwd_ws[:4 * self.y.C])
            A.check(A.lib.sap3d_affine_act_bwd_sync(*args, cnt, 2), "affine_act_bwd apply " + self.name)
        else:
            A.check(A.lib.sap3d_affine_act_bwd(*args), "affine_act_bwd " + self.name)
        e._count(3)


class _PoolOp:
    def __init__(self, eng, x: T, ksize, strides, same, name):
        self.eng, self.x = eng, x
        self.k, self.s, self.same = A.i3(ksize), A.i3(strides), int(same)
        N, D, H, W, Cc = x.shape
        out = (C.c_int32 * 3)()
        A.check(A.lib.sap3d_maxpool3d_out_dims(D, H, W, self.k, self.s, self.same, out), "maxpool dims")
        self.y = eng.tensor((N, out[0], out[1], out[2], Cc), name, needs_grad=x.needs_grad)
        self.name = name
        # window-local arg-max (uint8) recorded by the forward pass of training graphs for the gather-form backward
        self.amax = torch.empty(self.y.shape, device=eng.device, dtype=torch.uint8) if (eng.training_graph and x.needs_grad) else None

    def fwd(self):
        e, x = self.eng, self.x
        N, D, H, W, Cc = x.shape
        A.check(A.lib.sap3d_maxpool3d_fwd(e.dt, A.ptr(x.buf), N, D, H, W, Cc, self.k, self.s, self.same, A.ptr(self.y.buf),
                                          A.ptr(self.amax), e.stream), "maxpool_fwd " + self.name)
        e._count()

    def bwd(self):
        e, x = self.eng, self.x
        if not self.y.gflag or not x.needs_grad:
            return
        N, D, H, W, Cc = x.shape
        acc = x.take_acc()
        A.check(A.lib.sap3d_maxpool3d_bwd(e.dt, A.ptr(x.buf), A.ptr(self.y.grad), N, D, H, W, Cc, self.k, self.s, self.same,
                                          A.ptr(self.amax), A.ptr(x.ensure_grad()), acc, e.stream), "maxpool_bwd " + self.name)
        e._count()


class _DropoutOp:
    def __init__(self, eng, x: T, rate, name):
        self.eng, self.x, self.rate, self.name = eng, x, float(rate), name
        self.y = eng.tensor(x.shape, name)
        self.index = eng._dropout_ops          # position of this op among the engine's dropout ops
        eng._dropout_ops += 1

    @property
    def seed(self):
        """(engine seed, op index): read at launch / capture time, so attach_data_parallel can fold the rank into
        eng.dropout_seed before the graphs are captured (replicas must not draw the same masks)"""
        return (self.eng.dropout_seed * 1000003 + self.index) & 0x7FFFFFFFFFFFFFFF

    def fwd(self):
        e = self.eng
        A.check(A.lib.sap3d_dropout(e.dt, A.ptr(self.x.buf), A.ptr(self.y.buf), self.x.buf.numel(), self.rate, self.seed,
                                    A.ptr(e.step), 0, e.stream), "dropout " + self.name)
        e._count()

    def bwd(self):
        e = self.eng
        if not self.y.gflag or not self.x.needs_grad:
            return
        acc = self.x.take_acc()
        A.check(A.lib.sap3d_dropout(e.dt, A.ptr(self.y.grad), A.ptr(self.x.ensure_grad()), self.x.buf.numel(), self.rate,
                                    self.seed, A.ptr(e.step), acc, e.stream), "dropout_bwd " + self.name)
        e._count()


class _HeadOp:
    """final 1-channel transposed conv (+ sigmoid) and, in training graphs, the smooth-L1 loss"""

    def __init__(self, eng, x: T, w: Param, b: Param, ksize, stride, sigmoid, name):
        self.eng, self.x, self.w, self.b = eng, x, w, b
        self.k, self.stride, self.sigmoid, self.name = A.i3(ksize), int(stride), sigmoid, name
        N, D, H, W, _ = x.shape
        shp = (N, D * stride, H * stride, W * stride, 1)
        self.logits = torch.empty(shp, device=eng.device, dtype=torch.float32)
        self.pred = torch.empty(shp, device=eng.device, dtype=torch.float32)
        self.target = torch.zeros(shp[:-1], device=eng.device, dtype=torch.float32) if eng.training_graph else None
        self.dlogits = torch.empty(shp, device=eng.device, dtype=torch.float32) if eng.training_graph else None
        # tensor-core form ([positions x C] x [C x 27] GEMM + col2im) for the k3 s2 heads of the bf16 graphs
        self.loss_sigma, self.loss_w_in, self.loss_w_out = 1.0, 1.0, 1.0   # network.smooth_l1_loss arguments (train.py:159)
        self.use_tc = eng.dt == A.BF16 and x.C % 64 == 0 and tuple(ksize) == (3, 3, 3) and self.stride == 2
        if self.use_tc:
            self.ws = torch.empty(A.lib.sap3d_head_tc_workspace(N, D, H, W, x.C) // 4 + 16, device=eng.device, dtype=torch.float32)

    def fwd(self):
        e, x = self.eng, self.x
        N, D, H, W, Cc = x.shape
        if self.use_tc:
            A.check(A.lib.sap3d_head_tc_fwd(A.ptr(x.buf), N, D, H, W, Cc, A.ptr(self.w.w), A.ptr(self.b.w), A.ptr(self.logits),
                                            A.ptr(self.pred) if self.sigmoid else None, A.ptr(self.ws), e.stream), "head_tc_fwd " + self.name)
            e._count(3)
            return
        A.check(A.lib.sap3d_head_fwd(e.dt, A.ptr(x.buf), N, D, H, W, Cc, self.k, self.stride, A.ptr(self.w.w), A.ptr(self.b.w),
                                     A.ptr(self.logits), A.ptr(self.pred) if self.sigmoid else None, e.stream),
                "head_fwd " + self.name)
        e._count()

    def bwd(self):
        e, x = self.eng, self.x
        N, D, H, W, Cc = x.shape
        A.check(A.lib.sap3d_loss_smooth_l1_ex(A.ptr(self.logits), A.ptr(self.target), self.logits.numel(), int(self.sigmoid), None,
                                              A.ptr(self.dlogits), A.ptr(e.loss_buf), A.ptr(self.b.g), self.loss_sigma, self.loss_w_in,
                                              self.loss_w_out, e.stream), "loss")
        acc = x.take_acc()
        if self.use_tc:
            A.check(A.lib.sap3d_head_tc_bwd(A.ptr(self.dlogits), A.ptr(x.buf), N, D, H, W, Cc, A.ptr(x.ensure_grad()), acc, A.ptr(self.w.g),
                                            A.ptr(self.ws), e.stream, None), "head_tc_bwd " + self.name)
            e._count(6)
            return
        A.check(A.lib.sap3d_head_bwd(e.dt, A.ptr(self.dlogits), A.ptr(x.buf), N, D, H, W, Cc, self.k, self.stride,
                                     A.ptr(self.w.w), A.ptr(x.ensure_grad()), acc, A.ptr(self.w.g), e.stream),
                "head_bwd " + self.name)
        e._count(3)

    @property
    def output(self):
        return self.pred if self.sigmoid else self.logits


class _LogitsLossOp:
    """final tf.layers.conv3d -> 1 channel (gn/p3d_gn.py:538; the conv itself is an engine conv with fp32 output) and, in
    training graphs, smooth_l1_loss on it (gn/train_p3d_gn_dataset.py:186)"""

    def __init__(self, eng, co: ConvOut, name):
        self.eng, self.raw, self.name = eng, co.raw, name
        assert self.raw.buf.dtype == torch.float32, "logits_loss needs a conv built with out_f32=True"
        self.loss_sigma, self.loss_w_in, self.loss_w_out = 1.0, 1.0, 1.0
        shp = self.raw.shape
        self.logits = self.raw.buf
        self.target = torch.zeros(shp[:-1], device=eng.device, dtype=torch.float32) if eng.training_graph else None
        self.dlogits = torch.empty(shp, device=eng.device, dtype=torch.float32) if eng.training_graph else None
        if eng.training_graph:
            self.raw.grad = torch.empty(shp, device=eng.device, dtype=eng.tdt)   # the conv's backward reads dy in storage dtype

    def bwd(self):
        e = self.eng
        A.check(A.lib.sap3d_loss_smooth_l1_ex(A.ptr(self.logits), A.ptr(self.target), self.logits.numel(), 0, None,
                                              A.ptr(self.dlogits), A.ptr(e.loss_buf), None, self.loss_sigma, self.loss_w_in,
                                              self.loss_w_out, e.stream), "loss " + self.name)
        self.raw.take_acc()
        if e.dt == A.F32:
            self.raw.grad.copy_(self.dlogits)
        else:
            A.check(A.lib.sap3d_cast(A.F32, A.ptr(self.dlogits), A.ptr(self.raw.grad), self.dlogits.numel(), e.stream), "cast dlogits")
        e._count(2)

    @property
    def output(self):
        return self.logits


class _PaddedParams:
    """zero-padded fp32 copies of a filter [..., cin, cout] -> [..., cin, cout_pad] and its bias, refreshed before every
    weight packing; gradients of the padded copies are folded back into the real variables after the conv's backward"""

    def __init__(self, eng: Engine, w: Param, b: Optional[Param], cout_pad: int):
        self.eng, self.w, self.b, self.cp = eng, w, b, cout_pad
        self.c = w.shape[-1]
        self.rows = w.numel // self.c
        dev = eng.device
        mk = lambda shape, nm: _TempParam(nm, shape, dev, eng.training_graph)  # noqa: E731
        self.wp = mk((*w.shape[:-1], cout_pad), w.name + "#pad")
        self.bp = mk((cout_pad,), b.name + "#pad") if b is not None else None
        eng.pre_pack_ops.append(self.refresh)
        eng.fwd_ops.append(self.zero_grads)

    def refresh(self):
        st = self.eng.stream
        A.check(A.lib.sap3d_pad_channels(A.F32, A.ptr(self.w.w), A.ptr(self.wp.w), self.rows, self.c, self.cp, 0, 0, st), "pad filter")
        if self.b is not None:
            A.check(A.lib.sap3d_pad_channels(A.F32, A.ptr(self.b.w), A.ptr(self.bp.w), 1, self.c, self.cp, 0, 0, st), "pad bias")

    def zero_grads(self):
        if self.wp.g is not None:
            self.wp.g.zero_()
            if self.bp is not None:
                self.bp.g.zero_()

    def bwd(self):
        e = self.eng
        # the filter gradient was produced on the side stream: fold it back there (stream order = dependency)
        st = e.side_stream.cuda_stream if e.use_side_stream else e.stream
        A.check(A.lib.sap3d_pad_channels(A.F32, A.ptr(self.wp.g), A.ptr(self.w.g), self.rows, self.c, self.cp, 1, 1, st), "unpad dW")
        if self.b is not None:
            A.check(A.lib.sap3d_pad_channels(A.F32, A.ptr(self.bp.g), A.ptr(self.b.g), 1, self.c, self.cp, 1, 1, st), "unpad db")
        e._count(2)


class _TempParam:
    """parameter-shaped buffers that are not TF variables (not in the flat optimizer state)"""

    def __init__(self, name, shape, device, with_grad):
        self.name, self.shape = name, tuple(shape)
        self.numel = int(np.prod(self.shape))
        self.w = torch.zeros(self.shape, device=device, dtype=torch.float32)
        self.g = torch.zeros(self.shape, device=device, dtype=torch.float32) if with_grad else None
        self.trainable = False


class _GateOp:
    """y = o * gamma + x  (utils/network.py:191-192)"""

    def __init__(self, eng, o: T, x: T, gamma: Param, name):
        self.eng, self.o, self.x, self.gamma, self.name = eng, o, x, gamma, name
        self.y = eng.tensor(x.shape, name)

    def fwd(self):
        e = self.eng
        A.check(A.lib.sap3d_gate_fwd(e.dt, A.ptr(self.o.buf), A.ptr(self.x.buf), A.ptr(self.gamma.w), A.ptr(self.y.buf),
                                     self.y.buf.numel(), e.stream), "gate_fwd " + self.name)
        e._count()

    def bwd(self):
        e = self.eng
        if not self.y.gflag:
            return
        self.o.take_acc()
        dx_ptr, acc = None, 0
        if self.x.needs_grad:
            acc = self.x.take_acc()
            dx_ptr = A.ptr(self.x.ensure_grad())
        A.check(A.lib.sap3d_gate_bwd(e.dt, A.ptr(self.y.grad), A.ptr(self.o.buf), A.ptr(self.gamma.w), A.ptr(self.o.ensure_grad()),
                                     dx_ptr, acc, A.ptr(self.gamma.g), self.y.buf.numel(), e.stream), "gate_bwd " + self.name)
        e._count()


class _AttnCoreOp:
    """o[b] = softmax(g[b] f[b]^T) h[b]  (utils/network.py:184-186).  Large aligned problems run on the
    tcgen05 GEMMs (logits -> row softmax -> P.V), everything else on the generic CUDA-core kernels."""

    def __init__(self, eng, g: T, f: T, h: T, name):
        self.eng, self.g, self.f, self.h, self.name = eng, g, f, h, name
        B = g.shape[0]
        self.B = B
        self.Nq = int(np.prod(g.shape[1:4]))
        self.Nk = int(np.prod(f.shape[1:4]))
        self.dk, self.dv = g.C, h.C
        self.o = eng.tensor((*g.shape[:4], self.dv), name + "/o")
        dev = eng.device
        self.use_tc = eng.dt == A.BF16 and self.dv % 64 == 0 and self.dk % 8 == 0
        self.Nkp = (self.Nk + 63) // 64 * 64 if self.use_tc else self.Nk   # keys padded (zero probabilities) for K % 64
        self.ldb = self.Nkp
        tr = eng.training_graph
        # fused flash-style kernels (csrc/flash_attn.cu): the [Nq][Nk] score / probability matrices are never materialised
        self.use_flash = self.use_tc and self.dk == 64 and self.dv in (128, 256) and self.FLASH and (not tr or self.dv == 128)
        if self.use_flash:
            self.lse = torch.empty(B, self.Nq, device=dev, dtype=torch.float32)
            if tr:
                self.flash_ws = torch.empty(A.lib.sap3d_flash_attn_bwd_workspace(B, self.Nq, self.Nk, self.dv) // 4 + 16, device=dev,
                                            dtype=torch.float32)
            return
        self.beta = torch.zeros(B, self.Nq, self.ldb, device=dev, dtype=eng.tdt)
        if self.use_tc:
            self.dkp = (self.dk + 63) // 64 * 64
            self.pad = self.dkp != self.dk   # (network.attention hands over 64-padded projections: no pad kernels)
            bf = torch.bfloat16
            self.gp = torch.zeros(B, self.Nq, self.dkp, device=dev, dtype=bf) if self.pad else None
            self.fp = torch.zeros(B, self.Nk, self.dkp, device=dev, dtype=bf) if self.pad else None
            self.logits = torch.empty(self.Nq, self.Nkp, device=dev, dtype=torch.float32)
            self.vt = torch.zeros(B, self.dv, self.Nkp, device=dev, dtype=bf)
            if tr:
                self.ds = torch.empty(self.Nq, self.Nkp, device=dev, dtype=bf)
                self.ft = torch.zeros(self.dkp, self.Nkp, device=dev, dtype=bf)
                self.dv32 = torch.empty(self.Nkp, self.dv, device=dev, dtype=torch.float32)
                self.dk32 = torch.empty(self.Nkp, self.dkp, device=dev, dtype=torch.float32)
                self.dgp = torch.empty(B, self.Nq, self.dkp, device=dev, dtype=bf) if self.pad else None
                self.dfp = torch.empty(B, self.Nk, self.dkp, device=dev, dtype=bf) if self.pad else None
        elif tr:
            self.ds = torch.empty_like(self.beta)
        if self.use_tc and self.BATCHED:
            # all samples in one launch per product (sap3d_gemm_nt_batched): score / gradient buffers for the whole batch, and the
            # transposed operands that turn the two P^T.X products into NT form (query axis zero-padded to a multiple of 64)
            self.Nqp = (self.Nq + 63) // 64 * 64
            bf = torch.bfloat16
            self.logits = torch.empty(B, self.Nq, self.Nkp, device=dev, dtype=torch.float32)
            if tr:
                self.ds = torch.empty(B, self.Nq, self.Nkp, device=dev, dtype=bf)
                self.ft = torch.zeros(B, self.dkp, self.Nkp, device=dev, dtype=bf)
                self.pt = torch.zeros(B, self.Nkp, self.Nqp, device=dev, dtype=bf)     # beta^T, then dS^T
                self.dot = torch.zeros(B, self.dv, self.Nqp, device=dev, dtype=bf)     # dO^T
                self.gt = torch.zeros(B, self.dkp, self.Nqp, device=dev, dtype=bf)     # g^T
                self.dv32 = self.dk32 = None

    FLASH = True
    # one launch per product for the whole batch (sap3d_gemm_nt_batched); SAP3D_ATTN_BATCHED=0 restores the per-sample loops.
    # Measured r02: 20.04 -> 18.96 ms per B=8 training step, 1577 -> 1232 device activities
    BATCHED = os.environ.get("SAP3D_ATTN_BATCHED", "1") == "1"

    def _nt(self, a, lda, sa, b, ldb, sb, rows_b, c, ldc, sc, M, N, K, out_f32, what):
        A.check(A.lib.sap3d_gemm_nt_batched(A.ptr(a), lda, sa, A.ptr(b), ldb, sb, rows_b, A.ptr(c), ldc, sc, M, N, K, self.B, out_f32, 0,
                                            self.eng.stream), what + " " + self.name)

    def _fwd_tc_batched(self):
        e, st = self.eng, self.eng.stream
        B, Nq, Nk, Nkp, dkp, dv = self.B, self.Nq, self.Nk, self.Nkp, self.dkp, self.dv
        if self.pad:
            A.check(A.lib.sap3d_pad_channels(e.dt, A.ptr(self.g.buf), A.ptr(self.gp), B * Nq, self.dk, dkp, 0, 0, st), "pad g")
            A.check(A.lib.sap3d_pad_channels(e.dt, A.ptr(self.f.buf), A.ptr(self.fp), B * Nk, self.dk, dkp, 0, 0, st), "pad f")
            e._count(2)
        gq = self.gp if self.pad else self.g.buf
        fk = self.fp if self.pad else self.f.buf
        A.check(A.lib.sap3d_transpose(e.dt, A.ptr(self.h.buf), A.ptr(self.vt), B, Nk, dv, dv, Nkp, Nk * dv, dv * Nkp, st), "transpose h")
        self._nt(gq, dkp, Nq * dkp, fk, dkp, Nk * dkp, Nk, self.logits, Nkp, Nq * Nkp, Nq, Nkp, dkp, 1, "QK^T")
        A.check(A.lib.sap3d_softmax_rows(A.F32, A.ptr(self.logits), A.ptr(self.beta), B * Nq, Nk, Nkp, Nkp, st), "softmax")
        self._nt(self.beta, Nkp, Nq * Nkp, self.vt, Nkp, dv * Nkp, dv, self.o.buf, dv, Nq * dv, Nq, dv, Nkp, 0, "PV")
        e._count(4)

    def _bwd_tc_batched(self):
        e, st = self.eng, self.eng.stream
        B, Nq, Nk, Nkp, Nqp, dkp, dv, dk = self.B, self.Nq, self.Nk, self.Nkp, self.Nqp, self.dkp, self.dv, self.dk
        gq = self.gp if self.pad else self.g.buf
        fk = self.fp if self.pad else self.f.buf
        do = self.o.grad
        dh = self.h.ensure_grad()
        dg = self.dgp if self.pad else self.g.ensure_grad()
        df = self.dfp if self.pad else self.f.ensure_grad()
        tr = lambda src, dst, R, Cc, ld_in, ld_out, what: A.check(  # noqa: E731
            A.lib.sap3d_transpose(e.dt, A.ptr(src), A.ptr(dst), B, R, Cc, ld_in, ld_out, R * ld_in, Cc * ld_out, st), what)
        # dV = P^T dO  as  (P^T)[Nk x Nq] . (dO^T)[dv x Nq]^T
        tr(self.beta, self.pt, Nq, Nkp, Nkp, Nqp, "transpose P")
        tr(do, self.dot, Nq, dv, dv, Nqp, "transpose dO")
        self._nt(self.pt, Nqp, Nkp * Nqp, self.dot, Nqp, dv * Nqp, dv, dh, dv, Nk * dv, Nk, dv, Nqp, 0, "dV")
        # dP = dO V^T ;  dS = P o (dP - rowsum(dP o P))
        self._nt(do, dv, Nq * dv, self.h.buf, dv, Nk * dv, Nk, self.ds, Nkp, Nq * Nkp, Nq, Nkp, dv, 0, "dP")
        A.check(A.lib.sap3d_softmax_bwd_rows(A.ptr(self.beta), A.ptr(self.ds), B * Nq, Nk, Nkp, st), "softmax bwd")
        # dQ = dS F  as  dS[Nq x Nk] . (F^T)[dk x Nk]^T
        tr(fk, self.ft, Nk, dkp, dkp, Nkp, "transpose f")
        self._nt(self.ds, Nkp, Nq * Nkp, self.ft, Nkp, dkp * Nkp, dkp, dg, dkp, Nq * dkp, Nq, dkp, Nkp, 0, "dQ")
        # dK = dS^T G  as  (dS^T)[Nk x Nq] . (G^T)[dk x Nq]^T
        tr(self.ds, self.pt, Nq, Nkp, Nkp, Nqp, "transpose dS")
        tr(gq, self.gt, Nq, dkp, dkp, Nqp, "transpose g")
        self._nt(self.pt, Nqp, Nkp * Nqp, self.gt, Nqp, dkp * Nqp, dkp, df, dkp, Nk * dkp, Nk, dkp, Nqp, 0, "dK")
        e._count(10)
        if self.pad:
            A.check(A.lib.sap3d_pad_channels(e.dt, A.ptr(self.dgp), A.ptr(self.g.ensure_grad()), B * Nq, dk, dkp, 1, 0, st), "unpad dg")
            A.check(A.lib.sap3d_pad_channels(e.dt, A.ptr(self.dfp), A.ptr(self.f.ensure_grad()), B * Nk, dk, dkp, 1, 0, st), "unpad df")
            e._count(2)

    def _bwd_flash(self):
        e = self.eng
        A.check(A.lib.sap3d_flash_attn_bwd(A.ptr(self.g.buf), A.ptr(self.f.buf), A.ptr(self.h.buf), A.ptr(self.o.buf), A.ptr(self.o.grad),
                                           A.ptr(self.lse), A.ptr(self.g.ensure_grad()), A.ptr(self.f.ensure_grad()),
                                           A.ptr(self.h.ensure_grad()), self.B, self.Nq, self.Nk, self.dk, self.dv, A.ptr(self.flash_ws),
                                           e.stream), "flash_attn_bwd " + self.name)
        e._count(6)

    def _fwd_flash(self):
        e = self.eng
        A.check(A.lib.sap3d_flash_attn_fwd(A.ptr(self.g.buf), A.ptr(self.f.buf), A.ptr(self.h.buf), A.ptr(self.o.buf), A.ptr(self.lse),
                                           self.B, self.Nq, self.Nk, self.dk, self.dv, e.stream), "flash_attn_fwd " + self.name)
        e._count()

    # -- tensor-core path -------------------------------------------------------------------------
    def _fwd_tc(self):
        e, st = self.eng, self.eng.stream
        B, Nq, Nk, Nkp, dkp, dv = self.B, self.Nq, self.Nk, self.Nkp, self.dkp, self.dv
        if self.pad:
            A.check(A.lib.sap3d_pad_channels(e.dt, A.ptr(self.g.buf), A.ptr(self.gp), B * Nq, self.dk, dkp, 0, 0, st), "pad g")
            A.check(A.lib.sap3d_pad_channels(e.dt, A.ptr(self.f.buf), A.ptr(self.fp), B * Nk, self.dk, dkp, 0, 0, st), "pad f")
            e._count(2)
        gq = self.gp if self.pad else self.g.buf.view(B, Nq, dkp)
        fk = self.fp if self.pad else self.f.buf.view(B, Nk, dkp)
        hv = self.h.buf.view(B, Nk, dv)
        ob = self.o.buf.view(B, Nq, dv)
        A.check(A.lib.sap3d_transpose(e.dt, A.ptr(hv), A.ptr(self.vt), B, Nk, dv, dv, Nkp, Nk * dv, dv * Nkp, st), "transpose h")
        e._count()
        for b in range(B):
            A.check(A.lib.sap3d_gemm_nt(A.ptr(gq[b]), dkp, A.ptr(fk[b]), dkp, Nk, A.ptr(self.logits), Nkp, Nq, Nkp, dkp, 1, 0, st), "QK^T")
            A.check(A.lib.sap3d_softmax_rows(A.F32, A.ptr(self.logits), A.ptr(self.beta[b]), Nq, Nk, Nkp, Nkp, st), "softmax")
            A.check(A.lib.sap3d_gemm_nt(A.ptr(self.beta[b]), Nkp, A.ptr(self.vt[b]), Nkp, dv, A.ptr(ob[b]), dv, Nq, dv, Nkp, 0, 0, st), "PV")
        e._count(3 * B)

    def _bwd_tc(self):
        e, st = self.eng, self.eng.stream
        B, Nq, Nk, Nkp, dkp, dv, dk = self.B, self.Nq, self.Nk, self.Nkp, self.dkp, self.dv, self.dk
        gq = self.gp if self.pad else self.g.buf.view(B, Nq, dkp)
        fk = self.fp if self.pad else self.f.buf.view(B, Nk, dkp)
        hv = self.h.buf.view(B, Nk, dv)
        do = self.o.grad.view(B, Nq, dv)
        dh = self.h.ensure_grad().view(B, Nk, dv)
        dg = self.dgp if self.pad else self.g.ensure_grad().view(B, Nq, dkp)
        df = self.dfp if self.pad else self.f.ensure_grad().view(B, Nk, dkp)
        for b in range(B):
            self.dv32.zero_()
            A.check(A.lib.sap3d_gemm_tn(A.ptr(self.beta[b]), Nkp, A.ptr(do[b]), dv, A.ptr(self.dv32), dv, Nkp, dv, Nq, st), "dV")
            A.check(A.lib.sap3d_cast(A.F32, A.ptr(self.dv32), A.ptr(dh[b]), Nk * dv, st), "cast dV")
            A.check(A.lib.sap3d_gemm_nt(A.ptr(do[b]), dv, A.ptr(hv[b]), dv, Nk, A.ptr(self.ds), Nkp, Nq, Nkp, dv, 0, 0, st), "dP")
            A.check(A.lib.sap3d_softmax_bwd_rows(A.ptr(self.beta[b]), A.ptr(self.ds), Nq, Nk, Nkp, st), "softmax bwd")
            A.check(A.lib.sap3d_transpose(e.dt, A.ptr(fk[b]), A.ptr(self.ft), 1, Nk, dkp, dkp, Nkp, 0, 0, st), "transpose f")
            A.check(A.lib.sap3d_gemm_nt(A.ptr(self.ds), Nkp, A.ptr(self.ft), Nkp, dkp, A.ptr(dg[b]), dkp, Nq, dkp, Nkp, 0, 0, st), "dQ")
            self.dk32.zero_()
            A.check(A.lib.sap3d_gemm_tn(A.ptr(self.ds), Nkp, A.ptr(gq[b]), dkp, A.ptr(self.dk32), dkp, Nkp, dkp, Nq, st), "dK")
            A.check(A.lib.sap3d_cast(A.F32, A.ptr(self.dk32), A.ptr(df[b]), Nk * dkp, st), "cast dK")
        e._count(10 * B)
        if self.pad:
            A.check(A.lib.sap3d_pad_channels(e.dt, A.ptr(self.dgp), A.ptr(self.g.ensure_grad()), B * Nq, dk, dkp, 1, 0, st), "unpad dg")
            A.check(A.lib.sap3d_pad_channels(e.dt, A.ptr(self.dfp), A.ptr(self.f.ensure_grad()), B * Nk, dk, dkp, 1, 0, st), "unpad df")
            e._count(2)

    # -- dispatch ---------------------------------------------------------------------------------
    def fwd(self):
        e = self.eng
        if self.use_flash:
            return self._fwd_flash()
        if self.use_tc:
            return self._fwd_tc_batched() if self.BATCHED else self._fwd_tc()
        A.check(A.lib.sap3d_attention_fwd(e.dt, A.ptr(self.g.buf), A.ptr(self.f.buf), A.ptr(self.h.buf), A.ptr(self.beta),
                                          A.ptr(self.o.buf), self.B, self.Nq, self.Nk, self.dk, self.dv, self.dk, self.dk, self.dv,
                                          self.ldb, self.dv, e.stream), "attention_fwd " + self.name)
        e._count(2)

    def bwd(self):
        e = self.eng
        if not self.o.gflag:
            return
        for t in (self.g, self.f, self.h):
            if t.take_acc():
                raise A.Sap3dError("attention operands must have a single consumer")
        if self.use_flash:
            return self._bwd_flash()
        if self.use_tc:
            return self._bwd_tc_batched() if self.BATCHED else self._bwd_tc()
        A.check(A.lib.sap3d_attention_bwd(e.dt, A.ptr(self.g.buf), A.ptr(self.f.buf), A.ptr(self.h.buf), A.ptr(self.beta),
                                          A.ptr(self.o.grad), A.ptr(self.ds), A.ptr(self.g.ensure_grad()), A.ptr(self.f.ensure_grad()),
                                          A.ptr(self.h.ensure_grad()), self.B, self.Nq, self.Nk, self.dk, self.dv, self.dk, self.dk,
                                          self.dv, self.ldb, self.dv, e.stream), "attention_bwd " + self.name)
        e._count(3)
