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
           phase_out=False, block_n=0, kc=-1):
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
                                   rh, rw, prec, impl, int(phase_out), block_n, kc, _p(out)))
    return out
