"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of Pillow's 8-bit bicubic resampler, the
third-party arithmetic behind the reference's `image.resize(size, Image.BICUBIC)` calls in `slice_image`
(`modeling_minicpmv/modeling_minicpmv.py:509,519,531`; Pillow pinned at 10.1.0 by the reference's requirements.txt,
12.x installed here — same algorithm).

Follows Pillow `src/libImaging/Resample.c`: `bicubic_filter`, `precompute_coeffs`, `normalize_coeffs_8bpc`,
`ImagingResampleHorizontal_8bpc`, `ImagingResampleVertical_8bpc`, `ImagingResampleInner` (horizontal pass first, only
over the rows [ybox_first, ybox_last) the vertical pass reads; unchanged axes are skipped; equal sizes = copy, as
`Image.resize` returns `self.copy()`).

Pinned: tests/test_frontend_host.py compares `resize_bicubic` with `PIL.Image.resize` bit for bit over a sweep of
up/down-scaling shapes, so parity for this row is anchored on the real dependency, not on this file.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size):
    """-> ksize, bounds [(xmin, n)], double coefficients [out][ksize] (rows normalised to sum 1)."""
    filterscale = scale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds, kk = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        ww = 0.0
        for x in range(xmax):
            w = bicubic_filter((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        bounds.append((xmin, xmax))
        kk.append(k)
    return ksize, bounds, kk


def normalize_coeffs_8bpc(kk):
    one = float(1 << PRECISION_BITS)
    return [[int(-0.5 + v * one) if v < 0 else int(0.5 + v * one) for v in row] for row in kk]


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _pass(img, bounds, kk, axis):
    """One separable pass over `axis` (0 = vertical, 1 = horizontal) of a uint8 [H,W,C] array."""
    src = img.astype(np.int64)
    out_n = len(bounds)
    shape = list(img.shape)
    shape[axis] = out_n
    out = np.empty(shape, dtype=np.uint8)
    for i, ((lo, n), k) in enumerate(zip(bounds, kk)):
        acc = np.full(out[(i,) if axis == 0 else (slice(None), i)].shape, 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for j in range(n):
            acc = acc + (src[lo + j] if axis == 0 else src[:, lo + j]) * k[j]
        # Pillow accumulates in 32-bit ints; the sums stay far inside int32 for normalised kernels
        if axis == 0:
            out[i] = _clip8(acc)
        else:
            out[:, i] = _clip8(acc)
    return out


def resize_bicubic(img, out_w, out_h):
    """uint8 [H,W,3] -> uint8 [out_h,out_w,3], the pixels `PIL.Image.fromarray(img).resize((out_w,out_h), BICUBIC)` returns."""
    in_h, in_w = img.shape[:2]
    if (in_w, in_h) == (out_w, out_h):
        return img.copy()
    _, bh, kh = precompute_coeffs(in_w, out_w)
    _, bv, kv = precompute_coeffs(in_h, out_h)
    kh, kv = normalize_coeffs_8bpc(kh), normalize_coeffs_8bpc(kv)
    cur = img
    if out_w != in_w:
        first = bv[0][0]
        last = bv[-1][0] + bv[-1][1]
        cur = _pass(img[first:last], bh, kh, axis=1)
        bv = [(lo - first, n) for lo, n in bv]
    if out_h != in_h:
        cur = _pass(cur, bv, kv, axis=0)
    return cur
