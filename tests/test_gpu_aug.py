"""Stage parity (bit-exact, u8): device Pillow-exact resize / rotate kernels vs the oracle restatement
(which tests/test_oracle_kat.py pins bit-exactly against the installed Pillow)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(97, 133), (200, 300), (300, 200), (375, 500), (800, 1333)])
def test_device_pil_kernels_bit_exact(hw):
    from cald_b200 import ops, synth
    from oracle import pil_oracle as po
    h, w = hw
    rs = np.random.RandomState(h * 7 + w)
    for img in (rs.randint(0, 256, (h, w, 3)).astype(np.uint8), synth.synth_image(3, h, w)):
        got = ops.aug_image(2, img)
        want = po.cald_resize_image(img, 0.8)
        assert got.shape == want.shape and np.array_equal(got, want)
        got = ops.aug_image(3, img)
        want = po.cald_rotate_image(img, 5)[0]
        assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("hw", [(97, 133), (200, 300), (800, 1333)])
@pytest.mark.parametrize("factor", [1.5, 2, 5, 0.5])
def test_device_color_adjust_bit_exact(hw, factor):
    """cald_helper.ColorAdjust on the device vs the restatement that tests/test_pil_oracle.py pins to Pillow"""
    from cald_b200 import ops, synth
    from oracle import pil_oracle as po
    h, w = hw
    rs = np.random.RandomState(h + 3 * w)
    for img in (rs.randint(0, 256, (h, w, 3)).astype(np.uint8), synth.synth_image(5, h, w)):
        assert np.array_equal(ops.color_adjust(img, factor), po.cald_color_adjust(img, factor))
