"""TEST INFRASTRUCTURE — NumPy restatement of the counter-based dropout mask used by sap3d_dropout
(splitmix64 of (seed, element index); csrc/head.cu:dropout_keep).  TF's own Philox stream for
tf.layers.dropout (p3d.py:392) is not reproducible, so parity tests inject this mask into the oracle."""
import numpy as np


def keep_mask(seed: int, n: int, rate: float) -> np.ndarray:
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        h = idx + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        h ^= h >> np.uint64(30)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(27)
        h *= np.uint64(0x94D049BB133111EB)
        h ^= h >> np.uint64(31)
        u = (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return u >= np.float32(rate)
