"""placeholder / Session: the TF-1.x-shaped front door used by the drivers (train.py:141-217,
gen_pred.py:45-46,151, test.py:157-160), backed by the static engine and CUDA graphs.

    x = placeholder([B, 16, 112, 112, 3], dtype="bf16", training_graph=True)
    pred = p3d.p3d_unetplusplus_ds(x, 0.5, B, True)
    sess = Session(pred)
    loss = sess.train_step(clips, targets)          # fwd + smooth-L1 + bwd + Adam (train.py:217)
    sal  = sess.run(clips)                           # forward only (gen_pred.py:151)
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _abi as A
from .engine import Engine, T


def placeholder(shape, dtype: str = "bf16", training_graph: bool = False, conv_impl: int = A.IMPL_AUTO, device="cuda:0",
                dropout_seed: int = 1234, per_sample_statistics: bool = False) -> T:
    """tf.placeholder(tf.float32, [B, 16, H, W, 3]) (train.py:143): creates the engine and its input tensor.
    per_sample_statistics (inference graphs): batch-statistics BatchNorm layers normalise every clip on its own, so that
    a batch of B windows reproduces B single-window sess.run calls (gen_pred.py:45,151 feeds one window at a time)."""
    eng = Engine(dtype, training_graph, device, conv_impl, dropout_seed, per_sample_statistics)
    x = eng.tensor(shape, "input", needs_grad=False)
    eng.input = x
    eng.input_f32 = torch.zeros(tuple(shape), device=eng.device, dtype=torch.float32)
    return x


class Session:
    def __init__(self, head, lr: float = 1e-4):
        head = getattr(head, "head", head)     # a network.Loss (smooth_l1_loss(pred, y, 1, 1, sigma)) wraps the output op
        self.head = head
        self.eng: Engine = head.eng
        self.lr = lr
        if not self.eng.finalized:
            self.eng.finalize()
        self.graph_fwd: Optional[torch.cuda.CUDAGraph] = None
        self.graph_train: Optional[torch.cuda.CUDAGraph] = None
        self.grad_hook = None  # called between backward and adam (data-parallel all-reduce)
        self.grad_scale = 1.0

    # ---- input staging ---------------------------------------------------------------------------
    def _feed(self, x: torch.Tensor):
        e = self.eng
        if x.device.type != "cuda":
            e.input_f32.copy_(x, non_blocking=True)
        elif x.data_ptr() != e.input_f32.data_ptr():
            e.input_f32.copy_(x)
        self._stage_input()

    def _stage_input(self):
        e = self.eng
        if e.dt == A.F32:
            e.input.buf.copy_(e.input_f32)
        else:
            A.check(A.lib.sap3d_cast(A.F32, A.ptr(e.input_f32), A.ptr(e.input.buf), e.input_f32.numel(), e.stream), "cast input")

    # ---- eager execution -------------------------------------------------------------------------
    def forward_eager(self):
        self._stage_input()
        self.eng.forward()

    def _train_front(self, split: bool = False):
        e = self.eng
        self._stage_input()
        e.begin_step()
        e.update_moving = True      # UPDATE_OPS run with train_op only (train.py:170-172)
        try:
            e.forward()
        finally:
            e.update_moving = False
        e.backward(0 if split else None)

    def _overlap(self) -> bool:
        """data-parallel exchange overlapped with the tail of backward: needs the builder's split mark and a hook that can
        start on a partial gradient buffer"""
        return (len(self.eng.dp_segments) > 1 and self.grad_hook is not None and hasattr(self.grad_hook, "start")
                and getattr(self.grad_hook, "overlap", True))

    def _train_back(self):
        self.eng.adam(self.lr, grad_scale=self.grad_scale)

    def train_eager(self):
        self._train_front()
        if self.grad_hook is not None:
            self.grad_hook(self.eng)
        self._train_back()

    # ---- CUDA-graph capture ----------------------------------------------------------------------
    def capture(self, train: bool):
        """captures one forward, or one training step as two graphs (fwd+bwd | optimizer) so that the
        data-parallel gradient exchange can run between them.  One eager run first so lazily-allocated
        gradient buffers exist; the optimiser state is snapshotted around it."""
        torch.cuda.synchronize()
        e = self.eng
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            if train:
                snap = (e.flat_w.clone(), e.flat_m.clone(), e.flat_v.clone(), e.step.clone())
                self._train_front()
                self._train_back()
                torch.cuda.synchronize()
                e.flat_w.copy_(snap[0]); e.flat_m.copy_(snap[1]); e.flat_v.copy_(snap[2]); e.step.copy_(snap[3])
                e.pack_weights()
            else:
                self.forward_eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        # the critical path (forward, data gradients, norms) is captured on a HIGH-priority stream; the filter-gradient
        # branch forks onto the engine's default-priority side stream, so its many-CTA kernels only fill SMs the
        # few-CTA backbone kernels of the main chain leave idle instead of queueing in front of them
        cap = torch.cuda.Stream(device=e.device, priority=-1)
        if train:
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            split = self._overlap()
            e.branches_suspended = bool(split)   # several backward graphs: no cross-graph events (Engine._branches_active)
            with torch.cuda.graph(ga, stream=cap):
                self._train_front(split)
            gms = []
            if split:   # backward of the earlier segments, one graph each: replayed while the finished tail segments of the
                for k in range(1, len(e.dp_segments)):   # gradient buffer are being all-reduced
                    gm = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gm, pool=ga.pool(), stream=cap):
                        e.backward(k)
                    gms.append(gm)
            with torch.cuda.graph(gb, pool=ga.pool(), stream=cap):
                self._train_back()
            self.graph_train = (ga, gb, gms)
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cap):
                self.forward_eager()
            self.graph_fwd = g
        torch.cuda.synchronize()

    # ---- public API ------------------------------------------------------------------------------
    def run(self, x: torch.Tensor, graph: bool = False) -> torch.Tensor:
        """forward pass; returns the [B,16,H,W,1] fp32 saliency tensor (device).  x=None consumes the batch staged by prefetch().
        The returned tensor is the head's output BUFFER: the next run / train_step overwrites it (clone it to keep it)."""
        if x is None:
            if not hasattr(self, "_staged"):
                raise A.Sap3dError("run(None): no batch has been staged with prefetch()")
            self._take_prefetched(False)
            self._stage_input()
        else:
            self._feed(x)
        if graph:
            if self.eng.sync_bn is not None:
                raise A.Sap3dError("synchronised BatchNorm runs in eager mode only (run(graph=False))")
            if self.graph_fwd is None:
                self.capture(train=False)
            self.graph_fwd.replay()
        else:
            self.eng.forward()
        return self.head.output

    def train_step(self, x: torch.Tensor, y: torch.Tensor, graph: bool = False) -> torch.Tensor:
        """one iteration of train.py:217: returns the (device, fp64) loss scalar tensor -- the engine's loss BUFFER, which the
        next step overwrites (read it with .item() or clone it)."""
        e = self.eng
        if x is None:
            if not hasattr(self, "_staged"):
                raise A.Sap3dError("train_step(None, None): no batch has been staged with prefetch()")
            self._take_prefetched(True)      # batch staged by prefetch()
        else:
            if x.device.type != "cuda":
                e.input_f32.copy_(x, non_blocking=True)
            elif x.data_ptr() != e.input_f32.data_ptr():
                e.input_f32.copy_(x)
            self.head.target.copy_(y.reshape(self.head.target.shape), non_blocking=True)
        if graph:
            if e.sync_bn is not None:
                raise A.Sap3dError("synchronised BatchNorm runs in eager mode only (train_step(graph=False))")
            if self.graph_train is None:
                self.capture(train=True)
            ga, gb, gms = self.graph_train
            ga.replay()
            if gms:
                for k, gm in enumerate(gms):
                    self.grad_hook.start(e, k)   # all-reduce of the finished segment k starts on the comm stream ...
                    gm.replay()                  # ... while the layers ahead of it run their backward
                self.grad_hook.finish(e)         # last segment (stem + first stage), then wait for all of them
            elif self.grad_hook is not None:
                self.grad_hook(e)
            gb.replay()
        else:
            self.train_eager()
        return e.loss_buf

    # ---- input prefetch (the reference's tensorpack PrefetchDataZMQ / BatchData pipeline, train.py:120-135) -------------
    def prefetch(self, x: torch.Tensor, y: Optional[torch.Tensor] = None):
        """starts the host->device copy of the NEXT batch (pinned host tensors) on a copy stream, so that it overlaps the
        step that is currently running; `train_step(None, None)` / `run(None)` then consume the staged batch."""
        e = self.eng
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=e.device)
            self._stage_x = torch.empty_like(e.input_f32)
            self._stage_y = torch.empty_like(self.head.target) if getattr(self.head, "target", None) is not None else None
            self._staged = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(e.device))
        self._copy_stream.wait_event(self._consumed)           # the previous staged batch has been copied out
        with torch.cuda.stream(self._copy_stream):
            self._stage_x.copy_(x, non_blocking=True)
            if y is not None and self._stage_y is not None:
                self._stage_y.copy_(y.reshape(self._stage_y.shape), non_blocking=True)
            self._staged.record(self._copy_stream)

    def _take_prefetched(self, with_target: bool):
        e = self.eng
        cur = torch.cuda.current_stream(e.device)
        cur.wait_event(self._staged)
        e.input_f32.copy_(self._stage_x)                        # device-to-device, ~10 us
        if with_target and self._stage_y is not None:
            self.head.target.copy_(self._stage_y)
        self._consumed.record(cur)

    def tap(self, name: str) -> torch.Tensor:
        return self.eng.taps[name].buf

    def variables(self) -> Dict[str, torch.Tensor]:
        return {n: p.w for n, p in self.eng.params.items()}

    def gradients(self) -> Dict[str, torch.Tensor]:
        return {n: p.g for n, p in self.eng.params.items() if p.trainable}

    # ---- checkpoints (tf.train.Saver of train.py:180-185,266-267; gen_pred.py:57-64) -------------------------------------
    def save(self, prefix: str, include_optimizer: bool = False, max_to_keep: int = 10) -> str:
        """writes a TensorFlow tensor-bundle checkpoint of the reference's Saver var_list -- the trainable variables plus every
        moving_mean / moving_variance, under the reference's variable names -- and updates the directory's `checkpoint` file.
        `include_optimizer` adds tf.train.AdamOptimizer's slots (`<var>/Adam`, `<var>/Adam_1`, beta1_power, beta2_power),
        which the reference's saver leaves out (its resumed runs restart Adam)."""
        from . import checkpoint as ckpt
        e = self.eng
        torch.cuda.synchronize(e.device)
        tensors = {n: p.w.detach().cpu().numpy() for n, p in e.params.items()}
        if include_optimizer:
            if e.flat_m is None:
                raise RuntimeError("include_optimizer needs a training graph")
            m, v = e.flat_m.cpu().numpy(), e.flat_v.cpu().numpy()
            for n, p in e.params.items():
                if p.trainable:
                    sm, sv = ckpt.adam_slot_names(n)
                    tensors[sm] = m[p.offset:p.offset + p.numel].reshape(p.shape)
                    tensors[sv] = v[p.offset:p.offset + p.numel].reshape(p.shape)
            t = int(e.step.item())
            tensors["beta1_power"] = np.float32(0.9 ** (t + 1))      # TF multiplies the accumulators AFTER each apply
            tensors["beta2_power"] = np.float32(0.999 ** (t + 1))
            tensors["sap3d/adam_step"] = np.int64(t)                 # exact step; beta1_power underflows fp32 resolution late
        return ckpt.save(prefix, tensors, update_state=True, max_to_keep=max_to_keep)

    def restore(self, prefix_or_dir: str, strict: bool = True) -> str:
        """`saver.restore(sess, path)`: loads every variable of this graph by name from a tensor-bundle checkpoint (a prefix, or
        a directory whose `checkpoint` file names the newest one).  Adam slots are restored when present; otherwise the
        optimiser state is left as it is (the reference's behaviour).  Invalidates nothing: the graphs read the same buffers."""
        from . import checkpoint as ckpt
        e = self.eng
        prefix = prefix_or_dir
        if os.path.isdir(prefix_or_dir):
            prefix = ckpt.latest_checkpoint(prefix_or_dir)
            if prefix is None:
                raise ckpt.CheckpointError(f"no checkpoint state in {prefix_or_dir}")
        avail = {n for n, _ in ckpt.list_variables(prefix)}
        missing = [n for n in e.params if n not in avail]
        if strict and missing:
            raise KeyError(f"checkpoint {prefix} lacks {len(missing)} variables, e.g. {missing[:3]}")
        want = [n for n in e.params if n in avail]
        slots = []
        if e.flat_m is not None:
            for n, p in e.params.items():
                if p.trainable:
                    sm, sv = ckpt.adam_slot_names(n)
                    if sm in avail and sv in avail:
                        slots.append((p, sm, sv))
        extra = [s for _, a, b in slots for s in (a, b)] + [n for n in ("beta1_power", "sap3d/adam_step") if n in avail]
        vals = ckpt.load(prefix, want + extra)
        torch.cuda.synchronize(e.device)
        e.load_params({n: vals[n] for n in want}, strict=False)
        if slots and len(slots) == sum(1 for p in e.params.values() if p.trainable):
            for p, sm, sv in slots:
                e.flat_m[p.offset:p.offset + p.numel].copy_(torch.from_numpy(vals[sm].reshape(-1).copy()))
                e.flat_v[p.offset:p.offset + p.numel].copy_(torch.from_numpy(vals[sv].reshape(-1).copy()))
            if "sap3d/adam_step" in vals:
                t = int(vals["sap3d/adam_step"])
            elif "beta1_power" in vals:
                t = max(int(round(np.log(float(vals["beta1_power"])) / np.log(0.9))) - 1, 0)
            else:
                t = 0
            e.step.fill_(t)
        torch.cuda.synchronize(e.device)
        return prefix
