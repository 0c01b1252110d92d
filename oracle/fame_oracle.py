"""numpy restatement of the reference's FAME mask pipeline (TEST INFRASTRUCTURE; utils/transform/fame.py:29-110).

The reference delegates Gaussian blur and RGB->HSV to kornia, which is not installed here (no network), so the reference module
itself cannot be imported: parity against kornia's OUTPUTS is unpinned.  This file restates those two published definitions
independently of devias_b200/fame.py (scipy.ndimage for the blur, the textbook piecewise hue formula), is pinned together with the
product on hand-computed known answers of the definitions (tests/golden/fame_known_answers.json, sources stated there), and
follows the reference line by line for everything else; tests/test_fame.py compares the two implementations."""
import numpy as np
from scipy import ndimage


def gaussian_blur(img, ksize, sigma):
    """img [H, W]; normalised separable Gaussian, mirror border (kornia 'reflect')"""
    x = np.arange(ksize, dtype=np.float64) - ksize // 2
    k = np.exp(-x ** 2 / (2 * sigma ** 2)); k /= k.sum()
    out = ndimage.correlate1d(img.astype(np.float64), k, axis=1, mode='mirror')
    return ndimage.correlate1d(out, k, axis=0, mode='mirror')


def rgb_to_hsv(img, eps=1e-8):
    """img [3, H, W] in [0, 1] -> h [0, 2 pi), s, v   (kornia.color.rgb_to_hsv conventions)"""
    r, g, b = img.astype(np.float64)
    mx, mn = img.max(0).astype(np.float64), img.min(0).astype(np.float64)
    d = mx - mn
    s = d / (mx + eps)
    dd = np.where(d == 0, 1.0, d)
    h = np.where(mx == r, (g - b) / dd, np.where(mx == g, 2.0 + (b - r) / dd, 4.0 + (r - g) / dd))
    h = 2.0 * np.pi * ((h / 6.0) % 1.0)
    return h, s, mx


def norm01(m, eps=1e-8):
    m = m - m.min()
    return m / (m.max() + eps)


def soft_mask(diff, ksize, sigma):
    """fame.py:92-94: blurred, min-max normalised motion map of one clip"""
    return norm01(gaussian_blur(diff, ksize, sigma))


def refine(mask, clip, ksize, sigma, eps=1e-8):
    """fame.py:43-79 for one clip [3, T, H, W] (de-normalised): continuous refined mask BEFORE the top-k binarisation"""
    H, W = mask.shape
    h, s, v = rgb_to_hsv(clip.mean(1))
    order = np.argsort(-mask.reshape(-1), kind='stable')
    fg = order[:int(0.5 * H * W)]
    bg = np.argsort(mask.reshape(-1), kind='stable')[:int(0.1 * H * W)]
    hx = (s * np.cos(h * 2 * np.pi) + 1) / 2
    hy = (s * np.sin(h * 2 * np.pi) + 1) / 2
    cm = (np.round(hx * 9 + 1) + (np.round(hy * 9 + 1) - 1) * 10 + (np.round(v * 9 + 1) - 1) * 100).astype(np.int64).reshape(-1)
    dfg = np.bincount(cm[fg], minlength=1000).astype(np.float64)
    dbg = np.bincount(cm[bg], minlength=1000).astype(np.float64) + 1
    dfg /= dfg.sum() + eps
    dbg /= dbg.sum() + eps
    r = dfg[cm] / (dbg[cm] + dfg[cm])
    return norm01(gaussian_blur(r.reshape(H, W), ksize, sigma))


def binarise(m, beta):
    """fame.py:80-85: the beta H W largest entries become 1"""
    flat = m.reshape(-1)
    out = np.zeros_like(flat)
    out[np.argsort(-flat, kind='stable')[:int(beta * flat.size)]] = 1.0
    return out.reshape(m.shape)
