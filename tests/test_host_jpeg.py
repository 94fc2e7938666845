"""Host side of the pool ingest (no GPU): the C ABI's marker walk (cald_jpeg_info, include/cald_b200.h) against Pillow's
reading of the same files, and its refusals."""
import ctypes
import io

import numpy as np
import pytest
from PIL import Image

from cald_b200._lib import lib


def _info(data):
    L = lib()
    L.cald_jpeg_info.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.c_size_t, ctypes.POINTER(ctypes.c_int),
                                 ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    L.cald_last_error.restype = ctypes.c_char_p
    L.cald_last_error.argtypes = [ctypes.c_void_p]
    a = np.frombuffer(data, dtype=np.uint8)
    h, w, c = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    rc = L.cald_jpeg_info(a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), a.size, ctypes.byref(h), ctypes.byref(w),
                          ctypes.byref(c))
    return rc, h.value, w.value, c.value, L.cald_last_error(None).decode()


def _jpeg(img, **kw):
    buf = io.BytesIO()
    img.save(buf, format="JPEG", **kw)
    return buf.getvalue()


@pytest.mark.parametrize("h,w,kw", [(375, 500, dict(quality=90)), (500, 375, dict(quality=75, subsampling=1)),
                                    (37, 51, dict(quality=60, subsampling=0)), (1, 1, dict()),
                                    (64, 48, dict(restart_marker_blocks=2, optimize=True))])
def test_frame_size_and_components(h, w, kw):
    rs = np.random.RandomState(h + w)
    data = _jpeg(Image.fromarray(rs.randint(0, 256, (h, w, 3)).astype(np.uint8)), **kw)
    rc, gh, gw, gc, err = _info(data)
    assert rc == 0, err
    assert (gh, gw, gc) == (h, w, 3)
    gray = _jpeg(Image.fromarray(rs.randint(0, 256, (h, w)).astype(np.uint8)))
    assert _info(gray)[:4] == (0, h, w, 1)


def test_refusals_carry_a_message():
    img = Image.fromarray(np.zeros((32, 32, 3), np.uint8))
    rc, *_, err = _info(_jpeg(img, progressive=True))
    assert rc != 0 and "progressive" in err
    rc, *_, err = _info(_jpeg(img.convert("CMYK")))
    assert rc != 0 and "CMYK" in err
    rc, *_, err = _info(b"\x89PNG\r\n\x1a\n" + b"\0" * 32)
    assert rc != 0 and "not a JPEG" in err
    good = _jpeg(img)
    rc, *_, err = _info(good[:20])                       # cut inside the headers
    assert rc != 0
    # every prefix of a valid file is either parsed or refused -- never a crash or an out-of-bounds read
    for cut in range(2, len(good), 7):
        _info(good[:cut])


@pytest.mark.parametrize("h,w,kw", [(375, 500, dict(quality=90, subsampling=2)), (333, 499, dict(quality=85, subsampling=1)),
                                    (200, 300, dict(quality=95, subsampling=0)), (37, 51, dict(quality=60)),
                                    (1, 1, dict()), (241, 322, dict(quality=80, restart_marker_blocks=5)),
                                    (120, 160, dict(quality=30, optimize=True)), (90, 130, dict(quality=80, gray=True))])
def test_host_entropy_walk_equals_the_oracle(h, w, kw):
    """The host half of the default ingest path (engine.cu:UploadPipe walks the scans with jpeg_walk on host threads):
    the quantised coefficients of every block against oracle/jpeg_oracle.py's restatement of jdhuff.c, which
    tests/test_jpeg_oracle.py pins to Pillow end to end."""
    from oracle import jpeg_oracle as jo
    from cald_b200 import synth
    kw = dict(kw)
    img = synth.synth_image(900 + h, h, w)
    rs = np.random.RandomState(w)
    img = np.clip(img.astype(int) + rs.randint(-25, 25, img.shape), 0, 255).astype(np.uint8)
    pil = Image.fromarray(img.mean(-1).astype(np.uint8)) if kw.pop("gray", False) else Image.fromarray(img)
    data = _jpeg(pil, **kw)
    info = jo.parse(data)
    want, _ = jo.decode_coefficients(data, info)
    L = lib()
    L.cald_jpeg_coefficients.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.c_size_t, ctypes.POINTER(ctypes.c_int16),
                                         ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_int),
                                         ctypes.POINTER(ctypes.c_int)]
    a = np.frombuffer(data, dtype=np.uint8)
    n = ctypes.c_size_t(0)
    bw, bh = (ctypes.c_int * 3)(), (ctypes.c_int * 3)()
    fp = a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
    assert L.cald_jpeg_coefficients(fp, a.size, None, 0, ctypes.byref(n), bw, bh) != 0      # size query: too small
    assert n.value == sum(c.size for c in want)
    out = np.full(n.value + 8, 12345, dtype=np.int16)
    assert L.cald_jpeg_coefficients(fp, a.size, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), n.value,
                                    ctypes.byref(n), bw, bh) == 0, L.cald_last_error(None)
    assert (out[n.value:] == 12345).all()                                                   # nothing written past the end
    off = 0
    for q, c in enumerate(want):
        assert (bh[q], bw[q]) == c.shape[:2]
        got = out[off:off + c.size].reshape(c.shape)
        assert np.array_equal(got, c), (q, np.argwhere(got != c)[:4])
        off += c.size
