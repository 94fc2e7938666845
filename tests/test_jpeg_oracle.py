"""Pins oracle/jpeg_oracle.py (the CPU restatement of libjpeg's baseline decode path) bit for bit against the installed
Pillow -- Image.open(...).convert('RGB') is how the reference reads its pools (detection/voc_utils.py:52-58,
detection/coco_utils.py:209-220)."""
import io

import numpy as np
import pytest
from PIL import Image

from oracle import jpeg_oracle as jo


def make_jpeg(h, w, seed, gray=False, **save_kw):
    from cald_b200 import synth
    img = synth.synth_image(1000 * h + w + seed, h, w)
    rs = np.random.RandomState(seed)
    img = np.clip(img.astype(int) + rs.randint(-25, 25, img.shape), 0, 255).astype(np.uint8)   # high frequencies too
    if gray:
        img = img.mean(-1).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", **save_kw)
    return buf.getvalue()


CASES = [
    dict(h=48, w=64, quality=90, subsampling=0),
    dict(h=48, w=64, quality=75, subsampling=1),
    dict(h=48, w=64, quality=75, subsampling=2),
    dict(h=37, w=51, quality=60, subsampling=2),       # sizes that are not MCU multiples: edge rules of the upsampler
    dict(h=37, w=51, quality=95, subsampling=1),
    dict(h=33, w=17, quality=85, subsampling=2),
    dict(h=8, w=8, quality=50, subsampling=2),          # one chroma sample per row
    dict(h=1, w=1, quality=50, subsampling=2),
    dict(h=50, w=70, quality=80, subsampling=2, restart_marker_blocks=3),
    dict(h=41, w=23, quality=30, subsampling=0, optimize=True),
    dict(h=64, w=64, quality=100, subsampling=2),
    dict(h=40, w=56, quality=80, gray=True),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join("%s%s" % (k[0], v) for k, v in c.items()))
def test_oracle_decodes_what_pillow_decodes(case):
    case = dict(case)
    data = make_jpeg(case.pop("h"), case.pop("w"), 3, **case)
    want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    got = jo.decode(data)
    assert got.shape == want.shape
    assert np.array_equal(got, want), np.abs(got.astype(int) - want.astype(int)).max()


def test_progressive_is_refused():
    data = make_jpeg(32, 32, 1, quality=80, progressive=True)
    with pytest.raises(jo.JpegError):
        jo.decode(data)
