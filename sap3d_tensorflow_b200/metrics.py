"""Saliency metrics with the public names of the reference's utils/metrics.py (CC :227, SIM :258, NSS :200,
KLdiv :338), evaluated by one fused CUDA kernel per batch of maps (csrc/metrics.cu).  The drivers score the
LAST frame of every clip against the density / fixation maps (train.py:254-259, test.py:167-176) and average
the non-NaN values (test.py:177-181): `evaluate_clips` does exactly that on the device."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _abi as A


def _dev(x) -> torch.Tensor:
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    return t.to(device="cuda", dtype=torch.float32).contiguous()


def saliency_metrics(pred, density, fixation=None) -> torch.Tensor:
    """pred / density / fixation: [n, H, W] (or [H, W]).  Returns a device tensor [n, 4] (fp64): CC, SIM, NSS, KLdiv.
    NSS is NaN when no fixation map is given."""
    p, d = _dev(pred), _dev(density)
    if p.dim() == 2:
        p, d = p[None], d[None]
    f = None
    if fixation is not None:
        f = _dev(fixation)
        f = f[None] if f.dim() == 2 else f
    n, elems = p.shape[0], p[0].numel()
    out = torch.empty(n, 4, device=p.device, dtype=torch.float64)
    A.check(A.lib.sap3d_saliency_metrics(A.ptr(p), A.ptr(d), A.ptr(f), n, elems, elems, elems, elems, A.ptr(out),
                                         torch.cuda.current_stream().cuda_stream), "saliency_metrics")
    return out


def CC(saliency_map1, saliency_map2) -> float:
    return float(saliency_metrics(saliency_map1, saliency_map2)[0, 0])


def SIM(saliency_map1, saliency_map2) -> float:
    return float(saliency_metrics(saliency_map1, saliency_map2)[0, 1])


def NSS(saliency_map, fixation_map) -> float:
    return float(saliency_metrics(saliency_map, saliency_map, fixation_map)[0, 2])


def KLdiv(saliencyMap, fixationMap) -> float:
    return float(saliency_metrics(saliencyMap, fixationMap)[0, 3])


def _nan_sum_count(vals: torch.Tensor) -> Dict[str, torch.Tensor]:
    """NaN-filtered per-metric (sum, count) of a [B, m] fp64 table (test.py:177-181)"""
    n, m = vals.shape
    sums = torch.empty(m, device=vals.device, dtype=torch.float64)
    cnts = torch.empty(m, device=vals.device, dtype=torch.float64)
    A.check(A.lib.sap3d_nan_sum_count(A.ptr(vals), n, m, A.ptr(sums), A.ptr(cnts), torch.cuda.current_stream().cuda_stream), "nan_sum_count")
    return {"values": vals, "sum": sums, "count": cnts}


def resize_bilinear(maps, out_hw) -> torch.Tensor:
    """cv2.resize(map, (W, H)) (INTER_LINEAR) for [n, h, w] float maps — the upsampling test.py:168 applies to every
    predicted frame before scoring (112 x 112 -> 1080 x 960)"""
    m = _dev(maps)
    m = m[None] if m.dim() == 2 else m
    n, h, w = m.shape
    H, W = int(out_hw[0]), int(out_hw[1])
    out = torch.empty(n, H, W, device=m.device, dtype=torch.float32)
    A.check(A.lib.sap3d_resize_bilinear(A.ptr(m), n, h, w, A.ptr(out), H, W, torch.cuda.current_stream().cuda_stream), "resize_bilinear")
    return out


def saliency_auc(saliency, fixation, jitter: bool = False, n_rep: int = 100, step_size: float = 0.1, seed: int = 0) -> torch.Tensor:
    """[n, 2] (fp64): AUC_Judd, AUC_Borji of n (saliency, fixation) map pairs (utils/metrics.py:25-154).  NaN without
    fixations.  Borji's random locations and the optional jitter are counter hashes of `seed` (reproducible)."""
    s, f = _dev(saliency), _dev(fixation)
    if s.dim() == 2:
        s, f = s[None], f[None]
    n, elems = s.shape[0], s[0].numel()
    ws = torch.empty(A.lib.sap3d_saliency_auc_workspace(n, n_rep) // 4 + 16, device=s.device, dtype=torch.float32)
    out = torch.empty(n, 2, device=s.device, dtype=torch.float64)
    A.check(A.lib.sap3d_saliency_auc(A.ptr(s), A.ptr(f), n, elems, int(jitter), n_rep, float(step_size), int(seed), A.ptr(out), A.ptr(ws),
                                     torch.cuda.current_stream().cuda_stream), "saliency_auc")
    return out


def AUC_Judd(saliency_map, fixation_map, jitter=True, seed=0) -> float:
    """utils/metrics.py:25 (jitter defaults to True there too; train.py:259 and test.py:174 rely on the default).  The
    reference draws the jitter from numpy's global RNG (irreproducible); here it is a counter hash of `seed`."""
    return float(saliency_auc(saliency_map, fixation_map, jitter=jitter, seed=seed)[0, 0])


def AUC_Borji(saliency_map, fixation_map, n_rep=100, step_size=0.1, seed=0) -> float:
    return float(saliency_auc(saliency_map, fixation_map, n_rep=n_rep, step_size=step_size, seed=seed)[0, 1])


def evaluate_clips_test_time(pred: torch.Tensor, density: torch.Tensor, fixation: torch.Tensor, seed: int = 0) -> Dict[str, torch.Tensor]:
    """test.py:164-183: the last frame of every clip is upsampled to the ground-truth resolution (cv2.resize) and scored
    with CC, SIM, AUC_Judd, AUC_Borji, NSS.  pred [B,16,h,w,1]; density / fixation [B,H,W].  Returns the [B,5] values
    (CC, SIM, AUC_Judd, AUC_Borji, NSS) and the NaN-filtered (sum, count) pairs a sharded evaluation all-reduces."""
    p = pred.reshape(pred.shape[0], pred.shape[1], pred.shape[2], pred.shape[3])[:, -1].contiguous()
    up = resize_bilinear(p, density.shape[-2:])
    base = saliency_metrics(up, density, fixation)
    auc = saliency_auc(up, fixation, seed=seed)
    vals = torch.empty(base.shape[0], 5, device=base.device, dtype=torch.float64)      # column gather = memory plumbing
    vals[:, 0], vals[:, 1], vals[:, 2], vals[:, 3], vals[:, 4] = base[:, 0], base[:, 1], auc[:, 0], auc[:, 1], base[:, 2]
    return _nan_sum_count(vals)


def evaluate_clips(pred: torch.Tensor, density: torch.Tensor, fixation: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """pred [B,16,H,W,1] (Session.run output), density/fixation [B,16,H,W] or [B,H,W]: metrics of the last frame of
    every clip.  Returns per-metric (sum over non-NaN clips, count) pairs — the quantities a sharded evaluation
    all-reduces — plus the [B,4] per-clip values."""
    p = pred.reshape(pred.shape[0], pred.shape[1], pred.shape[2], pred.shape[3])[:, -1]
    d = density[:, -1] if density.dim() == 4 else density
    f = None if fixation is None else (fixation[:, -1] if fixation.dim() == 4 else fixation)
    return _nan_sum_count(saliency_metrics(p, d, f))
