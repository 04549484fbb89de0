"""TEST INFRASTRUCTURE — NumPy (float64) restatement of the reference's saliency metrics.

PINNED for CC / SIM / NSS: checked against the reference's own utils/metrics.py (imported from
/root/reference with a 3-line skimage stub) through the golden vectors of tests/golden/metrics_golden.npz
(generator: tests/golden/make_metrics_golden.py).  KLdiv cannot run in the reference as shipped
(`from scipy.misc import imresize`, metrics.py:348, was removed from SciPy): its golden values come from the
reference function body executed with the restated `imresize` below, so KLdiv is pinned only up to that
restatement (bytescale to uint8 + PIL bilinear resize, the documented scipy<=1.2 behaviour).

AUC_Judd (jitter off) and AUC_Borji (with the counter-hash `hash_sampler` passed as the reference's own `rand_sampler`
argument) are PINNED the same way; `resize_bilinear` restates cv2.resize(..., INTER_LINEAR) and is pinned against cv2
itself (tests/golden/metrics_auc_golden.npz).

Reference: utils/metrics.py CC :227-250, SIM :258-287, NSS :200-224, KLdiv :338-361, AUC_Judd :25-89, AUC_Borji :92-154;
utils/metric_utils.py normalize :10-53; test.py:164-183 (cv2.resize to 1080 x 960 before scoring).
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


# ---- test-time path: cv2.resize + AUC metrics (test.py:164-183) ---------------------------------------------------
def resize_bilinear(src, out_hw):
    """cv2.resize(src, (W, H), interpolation=cv2.INTER_LINEAR) for a 2-D float32 map: half-pixel centres, coordinates
    clamped to the border, horizontal pass then vertical pass in float32"""
    src = np.asarray(src, dtype=np.float32)
    h, w = src.shape
    H, W = out_hw

    def coords(n_out, n_in):
        f = (np.arange(n_out) + 0.5) * (n_in / n_out) - 0.5          # float64 coordinate; the FRACTION is rounded to float32
        i0 = np.floor(f).astype(np.int64)
        f = (f - i0).astype(np.float32)
        lo = i0 < 0
        i0[lo], f[lo] = 0, 0
        hi = i0 >= n_in - 1
        i0[hi], f[hi] = n_in - 1, 0
        return i0, np.minimum(i0 + 1, n_in - 1), f

    x0, x1, fx = coords(W, w)
    y0, y1, fy = coords(H, h)
    one = np.float32(1)
    rows = src[:, x0] * (one - fx)[None, :] + src[:, x1] * fx[None, :]          # [h, W]
    return (rows[y0] * (one - fy)[:, None] + rows[y1] * fy[:, None]).astype(np.float32)


_GOLD = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(x):
    with np.errstate(over="ignore"):
        x = np.asarray(x, dtype=np.uint64) + _GOLD
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def hash_sampler(seed: int):
    """a `rand_sampler` for AUC_Borji (utils/metrics.py:118-122,143): location of (fixation i, repetition r) =
    splitmix64(seed*GOLD + i*n_rep + r) mod n_pixels — the definition csrc/metrics_auc.cu uses"""
    def sampler(S, F, n_rep, n_fix):
        with np.errstate(over="ignore"):
            i = np.arange(n_fix, dtype=np.uint64)[:, None]
            r = np.arange(n_rep, dtype=np.uint64)[None, :]
            key = np.uint64(seed) * _GOLD + i * np.uint64(n_rep) + r
        return S[(splitmix64(key) % np.uint64(len(S))).astype(np.int64)]
    return sampler


def _trapz(y, x):
    y, x = np.asarray(y, dtype=np.float64), np.asarray(x, dtype=np.float64)
    return float(np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) * 0.5))


def AUC_Judd(saliency_map, fixation_map):
    """utils/metrics.py:25-89 with jitter=False (the jitter is unseeded np.random noise)"""
    S = np.asarray(saliency_map).ravel()
    F = (np.asarray(fixation_map) > 0.5).ravel()
    if not F.any():
        return float("nan")
    S_fix = S[F]
    n_fix, n_pixels = len(S_fix), len(S)
    thresholds = sorted(S_fix, reverse=True)
    tp, fp = np.zeros(n_fix + 2), np.zeros(n_fix + 2)
    tp[-1] = fp[-1] = 1
    for k, th in enumerate(thresholds):
        above = np.sum(S >= th)
        tp[k + 1] = (k + 1) / float(n_fix)
        fp[k + 1] = (above - k - 1) / float(n_pixels - n_fix)
    return _trapz(tp, fp)


def AUC_Borji(saliency_map, fixation_map, n_rep=100, step_size=0.1, seed=0):
    """utils/metrics.py:92-154 with rand_sampler = hash_sampler(seed)"""
    S = np.asarray(saliency_map)
    F = (np.asarray(fixation_map) > 0.5).ravel()
    if not F.any():
        return float("nan")
    S = ((S - S.min()) / (S.max() - S.min())).ravel()          # normalize(method='range'), dtype preserved
    S_fix = S[F]
    n_fix = len(S_fix)
    S_rand = hash_sampler(seed)(S, F, n_rep, n_fix)
    auc = np.zeros(n_rep)
    for rep in range(n_rep):
        thresholds = np.r_[0:np.max(np.r_[S_fix, S_rand[:, rep]]):step_size][::-1]
        tp, fp = np.zeros(len(thresholds) + 2), np.zeros(len(thresholds) + 2)
        tp[-1] = fp[-1] = 1
        for k, th in enumerate(thresholds):
            tp[k + 1] = np.sum(S_fix >= th) / float(n_fix)
            fp[k + 1] = np.sum(S_rand[:, rep] >= th) / float(n_fix)
        auc[rep] = _trapz(tp, fp)
    return float(np.mean(auc))


def preprocess_frame(frame_bgr_u8, size=112):
    """gen_pred.py:113-118 / dataflow.py:194-209: BGR uint8 -> RGB, minus [90, 102, 98], cv2.resize (INTER_LINEAR) on the
    float image, / 255 (restated with resize_bilinear per channel; pinned against cv2 in metrics_auc_golden.npz)"""
    rgb = np.asarray(frame_bgr_u8)[:, :, ::-1].astype(np.float32) - np.array([90.0, 102.0, 98.0], dtype=np.float32)
    out = np.stack([resize_bilinear(rgb[:, :, c], (size, size)) for c in range(3)], axis=-1)
    return (out / np.float32(255.0)).astype(np.float32)
