"""Generates tests/golden/metrics_golden.npz by running the REFERENCE's utils/metrics.py (imported from
/root/reference; build container only) on seeded synthetic maps.  skimage is absent from the image and is
stubbed (the metrics under test never call it); scipy.misc.imresize (removed from SciPy) is provided by the
restatement in oracle/metrics_oracle.py, so KLdiv runs the reference's own function body.

    python tests/golden/make_metrics_golden.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import metrics_oracle as MO  # noqa: E402

sk = types.ModuleType("skimage")
sk.img_as_float = lambda x: x
sk.exposure = types.ModuleType("skimage.exposure")
tr = types.ModuleType("skimage.transform")
tr.resize = None
sys.modules.update({"skimage": sk, "skimage.exposure": sk.exposure, "skimage.transform": tr})
import scipy  # noqa: E402

misc = types.ModuleType("scipy.misc")
misc.imresize = MO.imresize
sys.modules["scipy.misc"] = misc
scipy.misc = misc
if not hasattr(np, "float_"):
    np.float_ = np.float64  # metric_utils.py:43 uses the NumPy-1 alias
sys.path.insert(0, "/root/reference/utils")
import metrics as ref  # noqa: E402


def make_case(rng, h, w, kind):
    yy, xx = np.mgrid[0:h, 0:w]
    if kind == "blobs":
        pred = sum(np.exp(-((yy - rng.uniform(0, h)) ** 2 + (xx - rng.uniform(0, w)) ** 2) / (2 * rng.uniform(3, 15) ** 2)) for _ in range(3))
        dens = sum(np.exp(-((yy - rng.uniform(0, h)) ** 2 + (xx - rng.uniform(0, w)) ** 2) / (2 * rng.uniform(3, 15) ** 2)) for _ in range(3))
    elif kind == "uniform":
        pred, dens = rng.uniform(0, 1, (h, w)), rng.uniform(0, 1, (h, w))
    else:  # sigmoid-like network output vs 8-bit ground truth
        pred = 1.0 / (1.0 + np.exp(-rng.normal(0, 2, (h, w))))
        dens = rng.randint(0, 256, (h, w)) / 255.0
    fix = (rng.uniform(0, 1, (h, w)) < 0.01).astype(np.float32)
    fix[rng.randint(h), rng.randint(w)] = 1.0
    return pred.astype(np.float32), dens.astype(np.float32), fix


def main():
    rng = np.random.RandomState(1234)
    preds, denss, fixs, vals = [], [], [], []
    for i in range(12):
        p, d, f = make_case(rng, 112, 112, ["blobs", "uniform", "sigmoid"][i % 3])
        vals.append([ref.CC(p, d), ref.SIM(p, d), ref.NSS(p, f), ref.KLdiv(p, d)])
        preds.append(p); denss.append(d); fixs.append(f)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics_golden.npz")
    np.savez_compressed(out, pred=np.stack(preds), density=np.stack(denss), fixation=np.stack(fixs), values=np.array(vals, dtype=np.float64))
    print("wrote", out, np.array(vals)[:3])
    make_auc_golden()


def make_auc_golden():
    """AUC_Judd (jitter off) / AUC_Borji (rand_sampler = the counter hash of oracle.metrics_oracle.hash_sampler) from the
    REFERENCE's own functions, and cv2.resize itself for the test-time upsampling (test.py:168)."""
    import cv2

    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    rng = np.random.RandomState(4321)
    sal, fix, vals = [], [], []
    for i in range(8):
        p, _, f = make_case(rng, 56, 64, ["blobs", "uniform", "sigmoid"][i % 3])
        if i == 3:
            p = np.round(p * 8) / 8          # heavy ties in the saliency values
        judd = ref.AUC_Judd(p.copy(), f, jitter=False)
        borji = ref.AUC_Borji(p.copy(), f, n_rep=100, step_size=0.1, rand_sampler=MO.hash_sampler(5))
        sal.append(p.astype(np.float32)); fix.append(f); vals.append([judd, borji])
    small = np.stack([make_case(rng, 28, 28, "sigmoid")[0] for _ in range(3)])
    big = np.stack([cv2.resize(m, (120, 135)) for m in small])          # (W, H) = (120, 135): same 4.29 / 4.82 ratios as 112 -> 960 x 1080
    big2 = cv2.resize(small[0], (45, 17))                               # downscale path
    # input preprocessing exactly as gen_pred.py:113-118 writes it (cv2.resize on the mean-subtracted float image)
    frames = rng.randint(0, 256, (2, 90, 160, 3)).astype(np.uint8)          # 16:9 frames, BGR as cv2.imread returns them
    mean_value = np.array([98, 102, 90], dtype=np.float32)[::-1][None, ...]
    pre = []
    for fr in frames:
        im = fr[:, :, ::-1]
        im = im - mean_value
        im = cv2.resize(im, (112, 112))
        pre.append((im / 255.).astype(np.float32))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics_auc_golden.npz")
    np.savez_compressed(out, sal=np.stack(sal), fix=np.stack(fix), values=np.array(vals, dtype=np.float64), resize_src=small,
                        resize_135x120=big, resize_17x45=big2, frames_bgr=frames, frames_pre=np.stack(pre))
    print("wrote", out, np.array(vals))


if __name__ == "__main__":
    main()
