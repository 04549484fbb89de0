"""sap3d_tensorflow_b200 — B200-native hot path of the P3D video-saliency model.

Host code mirrors the reference's graph-builder surface (p3d.py, gn/p3d_gn.py, utils/network.py,
utils/metrics.py); all arithmetic runs in hand-written sm_100a kernels behind the C ABI of
include/sap3d.h (lib/libsap3d_b200.so).  There is no CPU or PyTorch fallback path.
"""
from . import _abi  # noqa: F401  (raises ImportError when the CUDA library has not been built)
from . import checkpoint, dataflow, gn, metrics, network, p3d, video  # noqa: F401
from .session import Session, placeholder  # noqa: F401

__all__ = ["Session", "placeholder", "p3d", "network", "metrics", "gn", "video", "checkpoint", "dataflow"]
