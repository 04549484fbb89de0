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


def evaluate_clips(pred: torch.Tensor, density: torch.Tensor, fixation: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """pred [B,16,H,W,1] (Session.run output), density/fixation [B,16,H,W] or [B,H,W]: metrics of the last frame of
    every clip.  Returns per-metric (sum over non-NaN clips, count) pairs — the quantities a sharded evaluation
    all-reduces — plus the [B,4] per-clip values."""
    p = pred.reshape(pred.shape[0], pred.shape[1], pred.shape[2], pred.shape[3])[:, -1]
    d = density[:, -1] if density.dim() == 4 else density
    f = None if fixation is None else (fixation[:, -1] if fixation.dim() == 4 else fixation)
    vals = saliency_metrics(p, d, f)
    ok = ~torch.isnan(vals)
    sums = torch.where(ok, vals, torch.zeros_like(vals)).sum(0)
    return {"values": vals, "sum": sums, "count": ok.sum(0).to(torch.float64)}
