"""Pins oracle/pil_oracle.py bit-exactly against the installed Pillow / torchvision PIL ops that
cald/cald_helper.py calls (resize 47-53, rotate 135-223, ColorAdjust 65-69)."""
import numpy as np
import pytest
from PIL import Image

from oracle import pil_oracle as po

SIZES = [(97, 133), (200, 300), (64, 48), (375, 500), (33, 250)]


def _img(seed, h, w):
    return np.random.RandomState(seed).randint(0, 256, (h, w, 3)).astype(np.uint8)


@pytest.mark.parametrize("h,w", SIZES)
@pytest.mark.parametrize("ratio", [0.8, 1.2, 0.7, 0.9])
def test_resize_bilinear_exact(h, w, ratio):
    img = _img(h * 7 + w, h, w)
    want = np.asarray(Image.fromarray(img).resize((int(w * ratio), int(h * ratio)), Image.BILINEAR))
    assert np.array_equal(po.cald_resize_image(img, ratio), want)


@pytest.mark.parametrize("h,w", SIZES)
def test_rotate_expand_then_bicubic_exact(h, w):
    img = _img(h + w, h, w)
    pil = Image.fromarray(img)
    want = np.asarray(pil.rotate(5, expand=True).resize((w, h)))
    got = po.cald_rotate_image(img, 5)
    got = got[0] if isinstance(got, tuple) else got
    assert np.array_equal(got, want)


@pytest.mark.parametrize("h,w", SIZES[:3])
@pytest.mark.parametrize("factor", [1.5, 2, 3, 0.5, 5])
def test_color_adjust_exact(h, w, factor):
    import torchvision.transforms.functional as F
    img = _img(3 * h + w, h, w)
    pil = Image.fromarray(img)
    b = F.adjust_brightness(pil, factor)
    assert np.array_equal(po.adjust_brightness(img, factor), np.asarray(b))
    c = F.adjust_contrast(b, factor)
    assert np.array_equal(po.adjust_contrast(np.asarray(b), factor), np.asarray(c))
    s = F.adjust_saturation(c, factor)
    assert np.array_equal(po.adjust_saturation(np.asarray(c), factor), np.asarray(s))
    assert np.array_equal(po.cald_color_adjust(img, factor), np.asarray(s))
