"""Python bindings of the stage-level entry points (include/cald_b200_ops.h).

Used by the stage-wise parity tests; every function runs the hand-written CUDA
kernel on cuda:0 with host numpy buffers in and out.
"""
import ctypes

import numpy as np

from ._lib import lib, check_ops

_f32p = ctypes.POINTER(ctypes.c_float)


def _p(a):
    return a.ctypes.data_as(_f32p) if a is not None else None


def conv2d(x_nhwc, weight_oihw, bias=None, stride=1, relu=False, res=None, res_mode=0, prec=0, impl=0,
           block_n=0, kc=-1):
    x = np.ascontiguousarray(x_nhwc, dtype=np.float32)
    w = np.ascontiguousarray(weight_oihw, dtype=np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    n, h, wd, cin = x.shape
    cout, cin2, k, _ = w.shape
    assert cin == cin2
    ho, wo = ((h + 1) // 2, (wd + 1) // 2) if stride == 2 else (h, wd)
    out = np.empty((n, ho, wo, cout), dtype=np.float32)
    r = None
    rh = rw = 0
    if res is not None:
        r = np.ascontiguousarray(res, dtype=np.float32)
        rh, rw = r.shape[1:3]
        if res_mode == 0:
            res_mode = 1
    check_ops(lib().cald_op_conv2d(_p(x), n, h, wd, cin, _p(w), _p(b), cout, k, stride, int(relu), _p(r), res_mode,
                                   rh, rw, prec, impl, block_n, kc, _p(out)))
    return out


def conv2d_dual(x_nhwc, weight_oihw, bias, x2_nhwc, weight2_oi11, bias2, stride2=1, relu=True):
    """act(conv(x, w) + b + conv1x1(x2[::s, ::s], w2) + b2) in one launch (fused projection shortcut)."""
    x = np.ascontiguousarray(x_nhwc, dtype=np.float32)
    w = np.ascontiguousarray(weight_oihw, dtype=np.float32)
    x2 = np.ascontiguousarray(x2_nhwc, dtype=np.float32)
    w2 = np.ascontiguousarray(weight2_oi11, dtype=np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    b2 = None if bias2 is None else np.ascontiguousarray(bias2, dtype=np.float32)
    n, h, wd, cin = x.shape
    cout, _, k, _ = w.shape
    _, h2, w2d, cin2 = x2.shape
    out = np.empty((n, h, wd, cout), dtype=np.float32)
    check_ops(lib().cald_op_conv2d_dual(_p(x), n, h, wd, cin, _p(w), _p(b), cout, k, _p(x2), h2, w2d, cin2, _p(w2),
                                        _p(b2), stride2, int(relu), _p(out)))
    return out


def aug_image(kind, img_u8):
    """Device Pillow-exact augmentation image: kind 2 = smaller_resize (0.8, BILINEAR), 3 = rotation (5 deg)."""
    img = np.ascontiguousarray(img_u8, dtype=np.uint8)
    h, w = img.shape[:2]
    out = np.zeros((h * w * 3,), dtype=np.uint8)
    oh, ow = ctypes.c_int(0), ctypes.c_int(0)
    L = lib()
    rc = L.cald_op_aug_image(kind, img.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), h, w,
                             out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), ctypes.byref(oh), ctypes.byref(ow))
    if rc != 0:
        L.cald_last_error.restype = ctypes.c_char_p
        L.cald_last_error.argtypes = [ctypes.c_void_p]
        raise RuntimeError(L.cald_last_error(None).decode())
    return out[:oh.value * ow.value * 3].reshape(oh.value, ow.value, 3)


def color_adjust(img_u8, factor):
    """Device ColorAdjust (cald_helper.py:65-69): PIL brightness -> contrast -> saturation, u8 in, u8 out."""
    img = np.ascontiguousarray(img_u8, dtype=np.uint8)
    h, w = img.shape[:2]
    out = np.zeros_like(img)
    L = lib()
    L.cald_op_color_adjust.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                       ctypes.POINTER(ctypes.c_uint8)]
    rc = L.cald_op_color_adjust(img.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), h, w, float(factor),
                                out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    if rc != 0:
        L.cald_last_error.restype = ctypes.c_char_p
        L.cald_last_error.argtypes = [ctypes.c_void_p]
        raise RuntimeError(L.cald_last_error(None).decode())
    return out

