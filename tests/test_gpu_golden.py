"""Engine vs fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py).

This is the parity claim of record: per-image consistency scores within 1e-3 of
cald_train.get_uncertainty's CPU fp32 output (BASELINE.json north_star), the class vectors within
1e-3, the same python-RNG stream position after the call, and -- for the selection stage -- the
identical index set when the engine's scores are pushed through select().
"""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']


@pytest.fixture(scope="module")
def eng():
    from cald_b200 import synth
    from cald_b200.engine import Engine
    e = Engine(depth=50, num_classes=21, min_size=320, max_size=512, max_views_per_pass=8)
    e.load_state_dict(synth.planted_frcnn_weights(50, 21, 0))
    return e


def test_detections_match_reference_fixture(eng):
    from cald_b200 import synth
    g = np.load(os.path.join(GOLD, "frcnn_r50_nc21_detect.npz"))
    imgs = [synth.synth_image(int(i), int(h), int(w)) for i, h, w in g["images"]]
    outs = eng.detect(imgs)
    for k, got in enumerate(outs):
        want_scores = g["%d_scores" % k]
        assert len(got["scores"]) == len(want_scores)      # every row of the detection list, not only its head
        n = len(want_scores)
        assert np.array_equal(got["labels"][:n], g["%d_labels" % k][:n])
        assert np.abs(got["scores"][:n] - want_scores[:n]).max() < 1e-3
        assert np.abs(got["boxes"][:n] - g["%d_boxes" % k][:n]).max() < 5e-2
        assert np.abs(got["props"][:n] - g["%d_props" % k][:n]).max() < 5e-2
        assert np.abs(got["scores_cls"][:n] - g["%d_scores_cls" % k][:n]).max() < 1e-3


def test_uncertainty_matches_reference_fixture(eng):
    from cald_b200 import api, synth
    g = np.load(os.path.join(GOLD, "frcnn_r50_nc21_uncertainty.npz"))
    imgs = [synth.synth_image(int(i), int(h), int(w)) for i, h, w in g["images"]]
    cons, cls = [], []
    for img, s in zip(imgs, g["seeds"]):
        random.seed(int(s))
        c, v = api.score_images(eng, [img], AUGS)
        cons.append(c[0])
        cls.append(v[0])
    err = np.abs(np.array(cons) - g["consistency"])
    print("engine", np.round(cons, 6), "reference", np.round(g["consistency"], 6), "err", err)
    assert err.max() <= 1e-3, err
    assert np.abs(np.array(cls) - g["cls"]).max() <= 1e-3


def test_rng_stream_position_matches_reference(eng):
    from cald_b200 import api, synth
    g = np.load(os.path.join(GOLD, "frcnn_r50_nc21_uncertainty.npz"))
    imgs = [synth.synth_image(int(i), int(h), int(w)) for i, h, w in g["images"][:3]]
    random.seed(int(g["cutout_only_seed"]))
    cons, _ = api.score_images(eng, imgs, ['cut_out'])
    assert random.random() == float(g["cutout_only_rng_tail"])
    assert np.abs(np.array(cons) - g["cutout_only_consistency"]).max() <= 1e-3


def test_topk_selection_identical(eng):
    """argsort of engine scores picks the same images as argsort of the reference's scores"""
    from cald_b200 import api, synth
    g = np.load(os.path.join(GOLD, "frcnn_r50_nc21_uncertainty.npz"))
    imgs = [synth.synth_image(int(i), int(h), int(w)) for i, h, w in g["images"]]
    cons = []
    for img, s in zip(imgs, g["seeds"]):
        random.seed(int(s))
        cons.append(api.score_images(eng, [img], AUGS)[0][0])
    assert np.array_equal(np.argsort(cons), np.argsort(g["consistency"]))   # the complete selection order
