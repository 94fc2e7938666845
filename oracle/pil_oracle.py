"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the Pillow arithmetic CALD uses.

The reference calls Pillow (an un-vendored third-party dependency; 12.2.0 is the
installed version of record) for two augmentations:

* ``img.resize((ow, oh), Image.BILINEAR)``        cald/cald_helper.py:53
* ``img.rotate(5, expand=True)`` (NEAREST) then ``img.resize((w, h))`` (BICUBIC
  default)                                        cald/cald_helper.py:153, 215

Pillow's source is not under /root/reference; this file restates its published
algorithm (libImaging/Resample.c: two-pass separable convolution on u8 with
22-bit fixed-point coefficients and support scaling; libImaging/Geometry.c:
nearest-neighbour affine in 16.16 fixed point; PIL/Image.py rotate(): matrix and
expanded size).  tests/test_pil_oracle.py pins it bit-exactly against the
installed Pillow on random images.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bilinear(x):
    x = np.abs(x)
    return np.where(x < 1.0, 1.0 - x, 0.0)


def _bicubic(x, a=-0.5):
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1,
                    np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))


FILTERS = {"bilinear": (_bilinear, 1.0), "bicubic": (_bicubic, 2.0)}


def precompute_coeffs(in_size, out_size, filt):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for box = full image.

    Returns (xmin[out], xcnt[out], kk[out, ksize] int32).
    """
    fn, support = FILTERS[filt]
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, dtype=np.int32)
    xcnt = np.zeros(out_size, dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        if lo < 0:
            lo = 0
        hi = int(center + support + 0.5)
        if hi > in_size:
            hi = in_size
        n = hi - lo
        w = fn((np.arange(n) + lo - center + 0.5) * ss).astype(np.float64)
        # Resample.c accumulates ww sequentially in double
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = w / ww
        # C casts double -> int by truncation toward zero
        k = np.where(w < 0, np.ceil(-0.5 + w * (1 << PRECISION_BITS)), np.floor(0.5 + w * (1 << PRECISION_BITS)))
        kk[xx, :n] = k.astype(np.int32)
        xmin[xx] = lo
        xcnt[xx] = n
    return xmin, xcnt, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _pass(img, out_size, filt, axis):
    """One separable pass along ``axis`` (1 = horizontal, 0 = vertical) of an HxWxC u8 image."""
    in_size = img.shape[axis]
    xmin, xcnt, kk = precompute_coeffs(in_size, out_size, filt)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        n = xcnt[xx]
        acc = np.tensordot(kk[xx, :n].astype(np.int64), src[xmin[xx]:xmin[xx] + n], axes=(0, 0))
        out[xx] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
    return np.moveaxis(out, 0, axis)


def resize(img, out_w, out_h, filt):
    """``Image.resize((out_w, out_h), filt)`` on an HxWxC u8 array (horizontal pass first)."""
    h, w = img.shape[:2]
    x = img
    if out_w != w:
        x = _pass(x, out_w, filt, 1)
    if out_h != h:
        x = _pass(x, out_h, filt, 0)
    return np.ascontiguousarray(x)


def rotate_matrix(w, h, angle_deg):
    """PIL/Image.py rotate(expand=True): inverse affine (dest -> src) and expanded size."""
    cx, cy = w / 2, h / 2
    ang = -math.radians(angle_deg)
    m = [round(math.cos(ang), 15), round(math.sin(ang), 15), 0.0,
         round(-math.sin(ang), 15), round(math.cos(ang), 15), 0.0]

    def tf(x, y):
        a, b, c, d, e, f = m
        return a * x + b * y + c, d * x + e * y + f

    m[2], m[5] = tf(-cx, -cy)
    m[2] += cx
    m[5] += cy
    xs, ys = [], []
    for x, y in ((0, 0), (w, 0), (w, h), (0, h)):
        tx, ty = tf(x, y)
        xs.append(tx)
        ys.append(ty)
    nw = math.ceil(max(xs)) - math.floor(min(xs))
    nh = math.ceil(max(ys)) - math.floor(min(ys))
    m[2], m[5] = tf(-(nw - w) / 2.0, -(nh - h) / 2.0)
    return m, nw, nh


def fixed_affine_coeffs(m):
    """Geometry.c affine_fixed(): 16.16 fixed point, FIX(v) = floor(v * 65536 + 0.5)."""
    def fix(v):
        return int(math.floor(v * 65536.0 + 0.5))
    a0, a1, a3, a4 = fix(m[0]), fix(m[1]), fix(m[3]), fix(m[4])
    a2 = fix(m[2] + m[0] * 0.5 + m[1] * 0.5)
    a5 = fix(m[5] + m[3] * 0.5 + m[4] * 0.5)
    return a0, a1, a2, a3, a4, a5


def rotate_nearest_expand(img, angle_deg):
    """``Image.rotate(angle, expand=True)`` (NEAREST, fill 0) on an HxWxC u8 array."""
    h, w = img.shape[:2]
    m, nw, nh = rotate_matrix(w, h, angle_deg)
    a0, a1, a2, a3, a4, a5 = fixed_affine_coeffs(m)
    xo = np.arange(nw, dtype=np.int64)[None, :]
    yo = np.arange(nh, dtype=np.int64)[:, None]
    xx = a2 + a1 * yo + a0 * xo
    yy = a5 + a4 * yo + a3 * xo
    xin = xx >> 16
    yin = yy >> 16
    ok = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
    out = np.zeros((nh, nw, img.shape[2]), dtype=np.uint8)
    out[ok] = img[yin[ok], xin[ok]]
    return out


def cald_rotate_image(img, angle_deg=5):
    """The image half of cald_helper.rotate (cald_helper.py:153, 215)."""
    h, w = img.shape[:2]
    r = rotate_nearest_expand(img, angle_deg)
    return resize(r, w, h, "bicubic"), r.shape[1], r.shape[0]


def cald_resize_image(img, ratio):
    """The image half of cald_helper.resize (cald_helper.py:47-53)."""
    h, w = img.shape[:2]
    return resize(img, int(w * ratio), int(h * ratio), "bilinear")


# ---------------------------------------------------------------- colour enhancement (cald_helper.py:56-69)
# torchvision F.adjust_brightness / adjust_contrast / adjust_saturation on a PIL image are PIL.ImageEnhance
# Brightness / Contrast / Color: Image.blend(degenerate, image, factor) with (libImaging/Blend.c)
#     temp = (float)((int)in1 + alpha * ((int)in2 - (int)in1))        alpha is a C float
#     0 <= alpha <= 1: out = (UINT8)temp            otherwise: clip to [0, 255], then (UINT8)temp
# and degenerate = black (Brightness) | grey of the rounded mean luma (Contrast) | per-pixel luma (Color), where
# luma = convert("L"): (R*19595 + G*38470 + B*7471 + 0x8000) >> 16 (libImaging/Convert.c rgb2l).
def luma(img):
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(deg, img, alpha):
    a = np.float32(alpha)
    in1 = deg.astype(np.int32)
    in2 = img.astype(np.int32)
    temp = in1.astype(np.float32) + a * (in2 - in1).astype(np.float32)
    if 0.0 <= alpha <= 1.0:
        return temp.astype(np.int32).astype(np.uint8)
    return np.where(temp <= 0, 0, np.where(temp >= 255, 255, temp.astype(np.int32))).astype(np.uint8)


def adjust_brightness(img, factor):
    return blend(np.zeros_like(img), img, factor)


def contrast_mean(img):
    """ImageEnhance.Contrast: int(ImageStat.Stat(image.convert("L")).mean[0] + 0.5)  (python float mean)"""
    l = luma(img)
    return int(float(int(l.astype(np.int64).sum())) / float(l.size) + 0.5)


def adjust_contrast(img, factor):
    return blend(np.full_like(img, contrast_mean(img)), img, factor)


def adjust_saturation(img, factor):
    return blend(np.repeat(luma(img)[..., None], 3, axis=2), img, factor)


def cald_color_adjust(img, factor):
    """cald_helper.ColorAdjust: brightness -> contrast -> saturation with the same factor, on u8."""
    return adjust_saturation(adjust_contrast(adjust_brightness(img, factor), factor), factor)
