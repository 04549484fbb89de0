"""TEST INFRASTRUCTURE — NumPy (float64) restatement of the reference's saliency metrics.

PINNED for CC / SIM / NSS: checked against the reference's own utils/metrics.py (imported from
/root/reference with a 3-line skimage stub) through the golden vectors of tests/golden/metrics_golden.npz
(generator: tests/golden/make_metrics_golden.py).  KLdiv cannot run in the reference as shipped
(`from scipy.misc import imresize`, metrics.py:348, was removed from SciPy): its golden values come from the
reference function body executed with the restated `imresize` below, so KLdiv is pinned only up to that
restatement (bytescale to uint8 + PIL bilinear resize, the documented scipy<=1.2 behaviour).

Reference: utils/metrics.py CC :227-250, SIM :258-287, NSS :200-224, KLdiv :338-361;
utils/metric_utils.py normalize :10-53.
"""
import numpy as np


def normalize(x, method="standard"):
    x = np.asarray(x, dtype=np.float64)
    if method == "standard":
        return (x - x.mean()) / x.std()
    if method == "range":
        return (x - x.min()) / (x.max() - x.min())
    if method == "sum":
        return x / float(x.sum())
    raise ValueError(method)


def CC(map1, map2):
    a = normalize(map1, "standard").ravel()
    b = normalize(map2, "standard").ravel()
    return float(np.corrcoef(a, b)[0, 1])


def SIM(map1, map2):
    a = normalize(normalize(map1, "range"), "sum")
    b = normalize(normalize(map2, "range"), "sum")
    return float(np.minimum(a, b).sum())


def NSS(saliency_map, fixation_map):
    s = normalize(saliency_map, "standard")
    f = np.asarray(fixation_map) > 0.5
    return float(np.mean(s[f]))


def bytescale(data):
    """scipy.misc.bytescale (scipy <= 1.2) with default cmin/cmax/high/low"""
    data = np.asarray(data)
    cmin, cmax = data.min(), data.max()
    cscale = cmax - cmin
    if cscale == 0:
        cscale = 1
    scale = 255.0 / cscale
    bytedata = (data - cmin) * scale
    return (bytedata.clip(0, 255) + 0.5).astype(np.uint8)


def imresize(arr, size):
    """scipy.misc.imresize(arr, shape) for a 2-D float array: bytescale -> PIL mode 'L' -> bilinear resize"""
    from PIL import Image

    im = Image.fromarray(bytescale(arr), mode="L")
    im = im.resize((size[1], size[0]), resample=Image.BILINEAR)
    return np.asarray(im)


def KLdiv(saliencyMap, fixationMap):
    map1 = np.asarray(saliencyMap).astype(np.float32)
    map2 = np.asarray(fixationMap).astype(np.float32)
    map1 = imresize(map1, np.shape(map2))
    if map1.any():
        map1 = map1 / map1.sum()
    if map2.any():
        map2 = map2 / map2.sum()
    eps = 2.2204e-16
    score = map2 * np.log(eps + map2 / (map1 + eps))
    return float(score.sum())


def all_metrics(pred, density, fixation):
    """[CC, SIM, NSS, KLdiv] as the drivers use them: CC/SIM/KLdiv against the density map, NSS against the
    fixation map (train.py:254-259, test.py:167-176)"""
    return np.array([CC(pred, density), SIM(pred, density), NSS(pred, fixation), KLdiv(pred, density)], dtype=np.float64)
