"""TEST INFRASTRUCTURE — independent float64 NumPy direct-loop restatement of the TF conv / transposed
conv / max-pool definitions, used only to cross-check oracle/tf_semantics.py (so the torch-based oracle
is not a single point of failure).  Deliberately slow and literal."""
import numpy as np


def same_pad(i, k, s):
    o = -(-i // s)
    total = max((o - 1) * s + k - i, 0)
    return o, total // 2


def conv3d_same(x, w, strides, bias=None):
    n, d, h, ww, ci = x.shape
    kd, kh, kw, _, co = w.shape
    (od, pd), (oh, ph), (ow, pw) = [same_pad(i, k, s) for i, k, s in zip((d, h, ww), (kd, kh, kw), strides)]
    y = np.zeros((n, od, oh, ow, co), dtype=np.float64)
    for a in range(od):
        for b in range(oh):
            for c in range(ow):
                acc = np.zeros((n, co))
                for i in range(kd):
                    zd = a * strides[0] + i - pd
                    if zd < 0 or zd >= d:
                        continue
                    for j in range(kh):
                        zh = b * strides[1] + j - ph
                        if zh < 0 or zh >= h:
                            continue
                        for l in range(kw):
                            zw = c * strides[2] + l - pw
                            if zw < 0 or zw >= ww:
                                continue
                            acc += x[:, zd, zh, zw, :].astype(np.float64) @ w[i, j, l].astype(np.float64)
                y[:, a, b, c, :] = acc
    if bias is not None:
        y += bias
    return y


def conv3d_transpose_same(x, w, strides, bias=None):
    """w: [kd,kh,kw,Cout,Cin]; y[p] = sum x[i] w[k], p = i*s + k - pb, pb = max(k-s,0)//2, size I*s"""
    n, d, h, ww, ci = x.shape
    kd, kh, kw, co, _ = w.shape
    od, oh, ow = d * strides[0], h * strides[1], ww * strides[2]
    pb = [max(k - s, 0) // 2 for k, s in zip((kd, kh, kw), strides)]
    y = np.zeros((n, od, oh, ow, co), dtype=np.float64)
    for a in range(d):
        for b in range(h):
            for c in range(ww):
                xv = x[:, a, b, c, :].astype(np.float64)
                for i in range(kd):
                    p0 = a * strides[0] + i - pb[0]
                    if p0 < 0 or p0 >= od:
                        continue
                    for j in range(kh):
                        p1 = b * strides[1] + j - pb[1]
                        if p1 < 0 or p1 >= oh:
                            continue
                        for l in range(kw):
                            p2 = c * strides[2] + l - pb[2]
                            if p2 < 0 or p2 >= ow:
                                continue
                            y[:, p0, p1, p2, :] += xv @ w[i, j, l].astype(np.float64).T
    if bias is not None:
        y += bias
    return y


def max_pool3d_same(x, ksize, strides):
    n, d, h, ww, c = x.shape
    (od, pd), (oh, ph), (ow, pw) = [same_pad(i, k, s) for i, k, s in zip((d, h, ww), ksize, strides)]
    y = np.full((n, od, oh, ow, c), -np.inf)
    for a in range(od):
        for b in range(oh):
            for e in range(ow):
                for i in range(ksize[0]):
                    zd = a * strides[0] + i - pd
                    if zd < 0 or zd >= d:
                        continue
                    for j in range(ksize[1]):
                        zh = b * strides[1] + j - ph
                        if zh < 0 or zh >= h:
                            continue
                        for l in range(ksize[2]):
                            zw = e * strides[2] + l - pw
                            if zw < 0 or zw >= ww:
                                continue
                            y[:, a, b, e, :] = np.maximum(y[:, a, b, e, :], x[:, zd, zh, zw, :])
    return y
