"""Pool ingest: the device JPEG decoder against Pillow, bit for bit, and scoring straight from
files against scoring the decoded pixels."""
import io
import random

import numpy as np
import pytest
from PIL import Image

pytestmark = pytest.mark.gpu
AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']


def jpeg_bytes(img, **kw):
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", **kw)
    return buf.getvalue()


@pytest.fixture(scope="module")
def eng():
    from cald_b200 import synth
    from cald_b200.engine import Engine
    e = Engine(depth=50, num_classes=21, min_size=320, max_size=512)
    e.load_state_dict(synth.planted_frcnn_weights(50, 21, 0))
    return e


@pytest.mark.parametrize("walk", ["host", "device"])
def test_device_decode_is_pillow_bit_for_bit(eng, walk, monkeypatch):
    """Both placements of the entropy walk (engine.cu:UploadPipe.host_walk): host threads walking the scans into the
    page-locked staging buffer (default) and the one-lane-per-image device kernel (CALD_JPEG_WALK=device)."""
    from cald_b200 import synth
    monkeypatch.setenv("CALD_JPEG_WALK", walk)
    rs = np.random.RandomState(0)
    files = []
    for k, (h, w, kw) in enumerate([
            (375, 500, dict(quality=90, subsampling=2)), (500, 375, dict(quality=75, subsampling=2)),
            (333, 499, dict(quality=85, subsampling=1)), (200, 300, dict(quality=95, subsampling=0)),
            (37, 51, dict(quality=60, subsampling=2)), (8, 8, dict(quality=50, subsampling=2)),
            (1, 1, dict(quality=50, subsampling=2)), (241, 322, dict(quality=80, subsampling=2, restart_marker_blocks=5)),
            (120, 160, dict(quality=30, subsampling=2, optimize=True)), (480, 640, dict(quality=92, subsampling=2)),
            (800, 1333, dict(quality=88, subsampling=2))]):
        img = synth.synth_image(700 + k, h, w)
        img = np.clip(img.astype(int) + rs.randint(-20, 20, img.shape), 0, 255).astype(np.uint8)
        files.append(jpeg_bytes(img, **kw))
    files.append(jpeg_bytes(synth.synth_image(3, 90, 130).mean(-1).astype(np.uint8), quality=80))   # grayscale
    # more files than one decode chunk holds: the double-buffered pipeline is exercised
    files = files + files[:6]
    got = eng.decode_jpeg(files)
    for k, (f, g) in enumerate(zip(files, got)):
        want = np.asarray(Image.open(io.BytesIO(f)).convert("RGB"))
        assert g.shape == want.shape, k
        assert np.array_equal(g, want), (k, np.abs(g.astype(int) - want.astype(int)).max())


@pytest.mark.parametrize("walk", ["host", "device"])
def test_scoring_files_equals_scoring_their_pixels(eng, walk, monkeypatch):
    from cald_b200 import api, synth
    monkeypatch.setenv("CALD_JPEG_WALK", walk)
    from cald_b200.engine import expand_augs
    imgs = [synth.synth_image(i, 200, 300) for i in range(5)] + [synth.synth_image(9, 300, 200)]
    files = [jpeg_bytes(im, quality=90, subsampling=2) for im in imgs]
    pixels = [np.asarray(Image.open(io.BytesIO(f)).convert("RGB")) for f in files]
    random.seed(11)
    want, want_cls = api.score_images(eng, pixels, AUGS)
    tail_want = random.random()
    random.seed(11)
    views = expand_augs(AUGS)
    u = np.array([random.random() for _ in range(200 * len(files))])
    cons, cls, used, hs, ws = eng.score_jpeg(files, views, 1.3, u)
    assert list(hs) == [p.shape[0] for p in pixels] and list(ws) == [p.shape[1] for p in pixels]
    assert np.array_equal(cons, np.array(want)) and np.array_equal(cls, np.array(want_cls))
    random.seed(11)
    for _ in range(used):
        random.random()
    assert random.random() == tail_want


def test_unsupported_files_fail_loudly(eng):
    from cald_b200 import synth
    from cald_b200._lib import CaldError
    img = synth.synth_image(1, 64, 64)
    with pytest.raises(CaldError, match="progressive"):
        eng.decode_jpeg([jpeg_bytes(img, quality=80, progressive=True)])
    with pytest.raises(CaldError):
        eng.decode_jpeg([b"not a jpeg at all"])
    buf = io.BytesIO()
    Image.fromarray(img).convert("CMYK").save(buf, format="JPEG")
    with pytest.raises(CaldError):
        eng.decode_jpeg([buf.getvalue()])
    # a truncated scan decodes to something (libjpeg pads with zeros too) but must not crash or hang
    good = jpeg_bytes(img, quality=80)
    out = eng.decode_jpeg([good[:len(good) // 2]])
    assert out[0].shape == (64, 64, 3)


def test_get_uncertainty_files_equals_the_pil_loader_path(tmp_path):
    """The pool given as JPEG paths (decoded on the device) against the same files through PIL and the reference-style
    loader entry point: identical scores, class vectors and python RNG position."""
    import torch
    import cald_b200
    from cald_b200 import synth

    class ModelLike:
        def __init__(self, w):
            self.sd = {k: torch.from_numpy(v) for k, v in w.items()}
            self.transform = type("T", (), {"min_size": (320,), "max_size": 512})()

        def state_dict(self):
            return self.sd

    model = ModelLike(synth.planted_frcnn_weights(50, 21, 0))
    paths = []
    for i in range(7):
        p = tmp_path / ("img%d.jpg" % i)
        h, w = ((200, 300), (300, 200), (167, 250))[i % 3]
        Image.fromarray(synth.synth_image(40 + i, h, w)).save(str(p), format="JPEG", quality=88)
        paths.append(str(p))

    class Loader:
        def __iter__(self):
            for p in paths:
                yield (Image.open(p).convert("RGB"),), (None,)     # detection/voc_utils.py:52-58
    random.seed(3)
    want, want_cls = cald_b200.get_uncertainty(model, Loader(), AUGS, 21)
    tail_want = random.random()
    random.seed(3)
    got, got_cls = cald_b200.get_uncertainty_files(model, paths, AUGS, 21)
    assert random.random() == tail_want
    assert got == want and all(np.array_equal(a, b) for a, b in zip(got_cls, want_cls))
    cald_b200.close_engines()
