"""End-to-end parity of the scoring path: CUDA engine vs the CPU oracle (cald_train.get_uncertainty restated).

Tolerance: |score_engine - score_oracle| <= 1e-3 (BASELINE.json north_star), class vectors <= 1e-3, on images
where the oracle itself is stable.  The per-(image, augmentation) consistency values are compared too.
"""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']


@pytest.fixture(scope="module")
def setup():
    from cald_b200 import synth
    from cald_b200.engine import Engine
    from oracle import frcnn_oracle as fo
    w = synth.planted_frcnn_weights(50, 21, 0)
    eng = Engine(depth=50, num_classes=21, min_size=320, max_size=512, max_views_per_pass=8)
    eng.load_state_dict(w)
    return eng, w, fo.Cfg(50, 21, 320, 512), fo, synth


def test_scores_match_oracle(setup):
    eng, w, cfg, fo, synth = setup
    from cald_b200 import api
    from oracle import cald_oracle as co
    imgs = [synth.synth_image(i, 200, 300) for i in range(5)] + [synth.synth_image(20, 300, 200)]
    seeds = [1000 + i for i in range(len(imgs))]
    # oracle: reseed python's RNG per image; engine: same draws through the uniforms interface
    want_c, want_v = [], []
    traces = []
    for img, s in zip(imgs, seeds):
        random.seed(s)
        tr = {}
        c, v = co.score_image(lambda x: fo.forward(x, w, cfg), img, AUGS, 21, 1.3, trace=tr)
        want_c.append(c)
        want_v.append(v)
        traces.append(tr)
    got_c, got_v = [], []
    for img, s in zip(imgs, seeds):
        random.seed(s)
        c, v = api.score_images(eng, [img], AUGS)
        got_c.append(c[0])
        got_v.append(v[0])
    per_view = None
    err = np.abs(np.array(got_c) - np.array(want_c))
    print("scores engine", np.round(got_c, 5), "oracle", np.round(want_c, 5))
    # hard bound on every image (north_star: 1e-3); the split-half operands + truncation pre-compensation put the
    # engine within ~2x of the fp32 CPU run's own rounding noise (tools/stage_error.py), so 1e-4 holds with margin
    assert err.max() <= 1e-4, err
    for g, wv in zip(got_v, want_v):
        assert np.abs(g - wv).max() <= 1e-3, np.abs(g - wv)


def test_ragged_pool_of_extreme_shapes(setup):
    """Edge shapes in ONE call (one chunk, views of different sizes padded to a common pass shape): a 33x47 and a 32x32
    thumbnail (up-scaled 10x by the transform), 10:1 and 1:10 strips (max_size binds), odd sizes either side of the
    512 bound -- each against the oracle scoring it alone."""
    eng, w, cfg, fo, synth = setup
    from cald_b200 import api
    from oracle import cald_oracle as co
    shapes = [(33, 47), (64, 640), (640, 64), (257, 255), (32, 32), (511, 513)]
    imgs = [synth.synth_image(300 + k, h, wd) for k, (h, wd) in enumerate(shapes)]
    want_c, want_v = [], []
    for k, img in enumerate(imgs):
        random.seed(50 + k)
        c, v = co.score_image(lambda x: fo.forward(x, w, cfg), img, AUGS, 21, 1.3)
        want_c.append(float(c))
        want_v.append(np.asarray(v))
    got_c, got_v = [], []
    for k, img in enumerate(imgs):
        random.seed(50 + k)
        c, v = api.score_images(eng, [img], AUGS)
        got_c.append(c[0])
        got_v.append(v[0])
    err = np.abs(np.array(got_c) - np.array(want_c))
    print("edge shapes: engine", np.round(got_c, 5), "oracle", np.round(want_c, 5), "max err %.2e" % err.max())
    assert err.max() <= 1e-3, err
    for g, wv in zip(got_v, want_v):
        assert np.abs(g - wv).max() <= 1e-3
    # the same images in ONE call: an image's score does not depend on what shares its pass (python's RNG is consumed
    # image by image in the same order, so seeding once and replaying the per-image draws is not possible here --
    # compare against the engine's own one-by-one results under one seed instead)
    random.seed(77)
    one_by_one = [api.score_images(eng, [im], AUGS)[0][0] for im in imgs]
    random.seed(77)
    together, _ = api.score_images(eng, imgs, AUGS)
    assert np.abs(np.array(together) - np.array(one_by_one)).max() <= 1e-6


def test_rng_stream_is_consumed_like_the_reference(setup):
    eng, w, cfg, fo, synth = setup
    from cald_b200 import api
    from oracle import cald_oracle as co
    imgs = [synth.synth_image(i, 200, 300) for i in range(3)]
    random.seed(77)
    for img in imgs:
        co.score_image(lambda x: fo.forward(x, w, cfg), img, ['cut_out'], 21, 1.3)
    tail_oracle = random.random()
    random.seed(77)
    api.score_images(eng, imgs, ['cut_out'])
    tail_engine = random.random()
    assert tail_engine == tail_oracle


def test_batched_equals_single(setup):
    eng, w, cfg, fo, synth = setup
    from cald_b200 import api
    imgs = [synth.synth_image(i, 200, 300) for i in range(4)]
    random.seed(5)
    a, av = api.score_images(eng, imgs, ['flip', 'smaller_resize', 'rotation'])
    b = []
    for img in imgs:
        c, _ = api.score_images(eng, [img], ['flip', 'smaller_resize', 'rotation'])
        b.append(c[0])
    assert np.allclose(a, b, atol=1e-6)


def test_degenerate_calls(setup):
    """empty pool -> empty lists; no augmentation -> np.mean([]) = nan per image (cald_train.py:225) with the
    reference view's class vector"""
    eng, w, cfg, fo, synth = setup
    from cald_b200 import api
    assert api.score_images(eng, [], AUGS) == ([], [])
    img = synth.synth_image(0, 200, 300)
    cons, cls = api.score_images(eng, [img], [])
    assert len(cons) == 1 and np.isnan(cons[0])
    full, full_cls = api.score_images(eng, [img], ['flip'])
    assert cls[0].shape == full_cls[0].shape and (cls[0] >= 0).all()


def test_reloading_weights_frees_the_previous_copy(setup):
    """get_uncertainty re-reads the state_dict every cycle (the model was retrained in between): no device leak"""
    eng, w, cfg, fo, synth = setup
    from cald_b200 import api
    img = synth.synth_image(2, 200, 300)
    random.seed(1)
    before = api.score_images(eng, [img], ['flip'])[0][0]
    eng.load_state_dict(w)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(4):
        eng.load_state_dict(w)
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < (32 << 20), (free0, free1)   # one copy is ~170 MB
    random.seed(1)
    assert api.score_images(eng, [img], ['flip'])[0][0] == before
