"""Saliency metrics: (CPU) the NumPy oracle against golden vectors produced by the reference's own
utils/metrics.py; (GPU) the fused CUDA kernel against the oracle and the golden vectors, metric values within
1e-3 (BASELINE.json tolerance; observed ~1e-6)."""
import os

import numpy as np
import pytest

from oracle import metrics_oracle as MO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_golden.npz")


def test_oracle_matches_reference_golden_vectors():
    g = np.load(GOLD)
    assert g["values"].shape == (12, 4)
    for i in range(g["values"].shape[0]):
        v = MO.all_metrics(g["pred"][i], g["density"][i], g["fixation"][i])
        np.testing.assert_allclose(v, g["values"][i], rtol=2e-6, atol=2e-6)


def test_oracle_against_live_reference_when_present():
    import sys
    import types

    ref_dir = "/root/reference/utils"
    if not os.path.isdir(ref_dir):
        pytest.skip("reference tree not present (GPU box)")
    sk = types.ModuleType("skimage")
    sk.img_as_float = lambda x: x
    sk.exposure = types.ModuleType("skimage.exposure")
    tr = types.ModuleType("skimage.transform")
    tr.resize = None
    sys.modules.update({"skimage": sk, "skimage.exposure": sk.exposure, "skimage.transform": tr})
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    sys.path.insert(0, ref_dir)
    import metrics as ref  # the reference's utils/metrics.py

    rng = np.random.RandomState(7)
    for _ in range(3):
        p, d = rng.rand(40, 50).astype(np.float32), rng.rand(40, 50).astype(np.float32)
        f = (rng.rand(40, 50) < 0.02).astype(np.float32)
        f[3, 4] = 1
        assert abs(ref.CC(p, d) - MO.CC(p, d)) < 1e-6
        assert abs(ref.SIM(p, d) - MO.SIM(p, d)) < 1e-6
        assert abs(ref.NSS(p, f) - MO.NSS(p, f)) < 1e-6


def test_metric_properties():
    rng = np.random.RandomState(0)
    p = rng.rand(32, 32)
    assert abs(MO.CC(p, p) - 1) < 1e-12 and abs(MO.CC(p, -p) + 1) < 1e-12
    assert abs(MO.SIM(p, p) - 1) < 1e-12
    assert abs(MO.CC(p, 3 * p + 2) - 1) < 1e-12      # CC is invariant to affine rescaling
    assert MO.KLdiv(p, p) < 0.05                        # only the uint8 quantisation of map1 remains


@pytest.mark.gpu
def test_cuda_metrics_match_golden_and_oracle(lib_built):
    import torch
    from sap3d_tensorflow_b200 import metrics as M

    g = np.load(GOLD)
    vals = M.saliency_metrics(g["pred"], g["density"], g["fixation"]).cpu().numpy()
    np.testing.assert_allclose(vals, g["values"], rtol=1e-3, atol=1e-3)
    assert np.abs(vals - g["values"]).max() < 1e-4
    # test.py-sized maps (1080 x 960) and the scalar API
    rng = np.random.RandomState(3)
    p, d = rng.rand(1080, 960).astype(np.float32), rng.rand(1080, 960).astype(np.float32)
    f = (rng.rand(1080, 960) < 0.001).astype(np.float32)
    ref = MO.all_metrics(p, d, f)
    got = [M.CC(p, d), M.SIM(p, d), M.NSS(p, f), M.KLdiv(p, d)]
    np.testing.assert_allclose(got, ref, rtol=1e-3, atol=1e-3)
    # flat prediction -> NaN exactly like NumPy (the drivers filter NaNs, test.py:177-181)
    flat = np.full((16, 16), 0.5, dtype=np.float32)
    v = M.saliency_metrics(flat, rng.rand(16, 16).astype(np.float32)).cpu().numpy()[0]
    assert np.isnan(v[0])
    ev = M.evaluate_clips(torch.rand(3, 16, 8, 8, 1, device="cuda"), torch.rand(3, 16, 8, 8, device="cuda"), (torch.rand(3, 16, 8, 8, device="cuda") > 0.7).float())
    assert ev["values"].shape == (3, 4) and float(ev["count"][0]) == 3
