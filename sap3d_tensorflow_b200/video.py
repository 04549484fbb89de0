"""Whole-video inference as gen_pred.py:88-168 does it: frames are preprocessed (BGR -> RGB, minus the channel mean,
cv2.resize to 112 x 112, / 255; gen_pred.py:113-118 == dataflow.py:194-209), every window of 16 consecutive frames is one
clip, the FIRST window contributes all 16 saliency maps and every later window only its last map (gen_pred.py:152-166).
The reference runs one window per sess.run; here consecutive windows are stacked into batches of the Session's size and
go through the captured CUDA graph, and the preprocessing is one kernel over all frames (sap3d_preprocess_frames)."""
from __future__ import annotations

import ctypes as C
from typing import Iterator, Tuple

import numpy as np
import torch

from . import _abi as A

MEAN_RGB = (90.0, 102.0, 98.0)   # gen_pred.py MEAN_VALUE = [98, 102, 90] (BGR) reversed
WINDOW = 16


def preprocess_frames(frames_bgr_u8, size: int = 112, dtype: str = "f32") -> torch.Tensor:
    """[T, h, w, 3] uint8 BGR (cv2.imread order) -> [T, size, size, 3] network-input frames on the device"""
    f = frames_bgr_u8 if isinstance(frames_bgr_u8, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(frames_bgr_u8))
    f = f.to(device="cuda", dtype=torch.uint8).contiguous()
    T, h, w, c = f.shape
    assert c == 3
    out = torch.empty(T, size, size, 3, device=f.device, dtype=torch.bfloat16 if dtype == "bf16" else torch.float32)
    mean = (C.c_float * 3)(*MEAN_RGB)
    A.check(A.lib.sap3d_preprocess_frames(A.ptr(f), T, h, w, mean, A.BF16 if dtype == "bf16" else A.F32, A.ptr(out), size, size,
                                          torch.cuda.current_stream().cuda_stream), "preprocess_frames")
    return out


def window_starts(n_frames: int) -> range:
    """start indices of the windows gen_pred.py runs (name_index <= len - 15, 1-based)"""
    return range(0, max(0, n_frames - WINDOW + 1))


def predict_video(sess, frames: torch.Tensor, graph: bool = True) -> Iterator[Tuple[int, torch.Tensor]]:
    """frames [T, H, W, 3] preprocessed (device, fp32).  Yields (frame index, saliency map [H, W]) in the order and with
    the selection rule of gen_pred.py: frames 0..15 from the first window, then the last frame of every later window."""
    B = sess.eng.input.shape[0]
    if B > 1 and not sess.eng.per_sample_bn:
        # the backbone's BatchNorm always uses batch statistics (p3d.py:140,350) and gen_pred.py feeds ONE window per
        # sess.run: stacking windows would normalise every layer over B clips (and over the padding copies of the last batch)
        raise A.Sap3dError("predict_video with a batch of windows needs placeholder(..., per_sample_statistics=True); "
                           "without it use a batch-1 session (the reference's gen_pred.py placeholder is [1,16,112,112,3])")
    starts = list(window_starts(frames.shape[0]))
    for lo in range(0, len(starts), B):
        chunk = starts[lo:lo + B]
        batch = torch.stack([frames[s:s + WINDOW] for s in chunk] + [frames[chunk[-1]:chunk[-1] + WINDOW]] * (B - len(chunk)))
        sal = sess.run(batch.float(), graph=graph)          # [B, 16, H, W, 1]
        for j, s in enumerate(chunk):
            if s == 0:
                for k in range(WINDOW):
                    yield k, sal[j, k, :, :, 0].clone()
            else:
                yield s + WINDOW - 1, sal[j, WINDOW - 1, :, :, 0].clone()


def stem_activations(stem_sess, frames: torch.Tensor) -> torch.Tensor:
    """frames [T, H, W, 3] (preprocessed) -> the per-frame stem output [T, H/2, W/2, 64] (p3d.p3d_stem: conv 1x7x7 + moving-statistics
    BatchNorm + ReLU), computed once per frame in chunks of the stem session's batch size"""
    F = stem_sess.eng.input.shape[0]
    assert stem_sess.eng.input.shape[1] == 1, "the stem session takes single frames: placeholder([F, 1, H, W, 3])"
    T = frames.shape[0]
    out = None
    for lo in range(0, T, F):
        chunk = frames[lo:lo + F]
        if chunk.shape[0] < F:
            chunk = torch.cat([chunk, chunk[-1:].expand(F - chunk.shape[0], *chunk.shape[1:])])
        act = stem_sess.run(chunk.float().unsqueeze(1), graph=True)          # [F, 1, h, w, 64]
        if out is None:
            out = torch.empty(T, *act.shape[2:], device=act.device, dtype=act.dtype)
        n = min(F, T - lo)
        out[lo:lo + n] = act[:n, 0]
    return out


def predict_video_cached(stem_sess, sess, frames: torch.Tensor, graph: bool = True) -> Iterator[Tuple[int, torch.Tensor]]:
    """predict_video with the per-frame stem cache: consecutive windows of gen_pred.py:88-135 share 15 of their 16 frames, and the
    stem (conv 1x7x7 s(1,2,2) + BatchNorm on moving statistics + ReLU, p3d.py:343-345 with training=False) is frame-local, so its
    [H/2, W/2, 64] output is computed once per frame (stem_sess = Session(p3d.p3d_stem(placeholder([F,1,H,W,3])))) and the window
    graph (sess, built on a placeholder of stem activations [B,16,H/2,W/2,64]) starts at the temporal pools.  Same maps as
    predict_video; 1/16 of the stem work per window."""
    B = sess.eng.input.shape[0]
    if sess.eng.input.shape[-1] != 64:
        raise A.Sap3dError("predict_video_cached: the window session must be built on a placeholder of stem activations [B,16,H/2,W/2,64]")
    if B > 1 and not sess.eng.per_sample_bn:
        raise A.Sap3dError("predict_video_cached with a batch of windows needs placeholder(..., per_sample_statistics=True)")
    acts = stem_activations(stem_sess, frames)
    starts = list(window_starts(frames.shape[0]))
    for lo in range(0, len(starts), B):
        chunk = starts[lo:lo + B]
        batch = torch.stack([acts[s:s + WINDOW] for s in chunk] + [acts[chunk[-1]:chunk[-1] + WINDOW]] * (B - len(chunk)))
        sal = sess.run(batch.float(), graph=graph)
        for j, s in enumerate(chunk):
            if s == 0:
                for k in range(WINDOW):
                    yield k, sal[j, k, :, :, 0].clone()
            else:
                yield s + WINDOW - 1, sal[j, WINDOW - 1, :, :, 0].clone()
