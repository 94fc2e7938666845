"""The remaining augmentation kinds of the reference whitelist (cald_train.py:93-94) vs the oracle:
Gaussian / salt-pepper noise (torch CPU RNG), multi_cut_out (python RNG, 1..4 cuts), larger / multi resize,
ColorAdjust (Pillow-exact enhancement on the device) and ColorSwap (python RNG, drawn before the cutout draws)."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from cald_b200 import synth
    from cald_b200.engine import Engine
    from oracle import frcnn_oracle as fo
    w = synth.planted_frcnn_weights(50, 21, 0)
    eng = Engine(depth=50, num_classes=21, min_size=256, max_size=416, max_views_per_pass=8)
    eng.load_state_dict(w)
    wt = {k: torch.from_numpy(v) for k, v in w.items()}
    cfg = fo.Cfg(50, 21, 256, 416)
    return eng, (lambda x: fo.forward(x, wt, cfg)), synth


@pytest.mark.parametrize("augs", [['ga', 'sp'], ['multi_cut_out', 'cut_out'], ['larger_resize', 'multi_resize'],
                                  ['flip', 'ga', 'cut_out', 'smaller_resize', 'rotation', 'sp'],
                                  ['color_adjust', 'color_swap'], ['color_swap', 'cut_out', 'color_adjust', 'flip']])
def test_aug_kinds_match_oracle(setup, augs):
    eng, fwd, synth = setup
    from cald_b200 import api
    from oracle import cald_oracle as co
    imgs = [synth.synth_image(i, 160, 240) for i in (0, 3)]
    torch.manual_seed(11)
    random.seed(11)
    want, want_cls = [], []
    for im in imgs:
        c, v = co.score_image(fwd, im, augs, 21, 1.3)
        want.append(c)
        want_cls.append(v)
    tail_o = (random.random(), float(torch.rand(1)))
    torch.manual_seed(11)
    random.seed(11)
    got, got_cls = api.score_images(eng, imgs, augs)
    tail_e = (random.random(), float(torch.rand(1)))
    assert tail_e == tail_o  # both RNG streams left exactly where the reference leaves them
    err = np.abs(np.array(got) - np.array(want))
    print(augs, "engine", np.round(got, 5), "oracle", np.round(want, 5))
    assert err.max() <= 1e-3, err
    for g, wv in zip(got_cls, want_cls):
        assert np.abs(g - wv).max() <= 1e-3, np.abs(g - wv).max()


def test_multi_color_adjust_raises_like_the_reference(setup):
    eng, fwd, synth = setup
    from cald_b200 import api
    with pytest.raises(NameError):
        api.score_images(eng, [synth.synth_image(0, 160, 240)], ['multi_color_adjust'])
