"""Pins the CPU oracle (oracle/) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run where /root/reference is mounted)."""
import os
import random

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']


@pytest.fixture(scope="module")
def model():
    from cald_b200 import synth
    from oracle import frcnn_oracle as fo
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    w = {k: torch.from_numpy(v) for k, v in synth.planted_frcnn_weights(50, 21, 0).items()}
    cfg = fo.Cfg(50, 21, 320, 512)
    return (lambda x: fo.forward(x, w, cfg)), synth


def test_forward_matches_reference_detections(model):
    fwd, synth = model
    g = np.load(os.path.join(GOLD, "frcnn_r50_nc21_detect.npz"))
    from oracle.cald_oracle import to_tensor
    for k, (idx, h, w) in enumerate(g["images"][:3]):
        out = fwd(to_tensor(synth.synth_image(int(idx), int(h), int(w))))
        for key in ("boxes", "scores", "labels", "props", "prob_max", "scores_cls"):
            want = g["%d_%s" % (k, key)]
            got = out[key].numpy()
            assert got.shape == want.shape, (k, key)
            # same ATen kernels; thread-count dependent summation order allows ~1e-6
            assert np.abs(got.astype(np.float64) - want).max() <= 2e-5, (k, key)


def test_get_uncertainty_matches_reference(model):
    fwd, synth = model
    from oracle import cald_oracle as co
    g = np.load(os.path.join(GOLD, "frcnn_r50_nc21_uncertainty.npz"))
    imgs = [synth.synth_image(int(i), int(h), int(w)) for i, h, w in g["images"][:2]]
    cons, cls = co.get_uncertainty(fwd, imgs, AUGS, 21, 1.3, seeds=[int(s) for s in g["seeds"][:2]])
    assert np.abs(np.array(cons) - g["consistency"][:2]).max() <= 5e-5
    assert np.abs(np.array(cls) - g["cls"][:2]).max() <= 5e-5


def test_augmentations_match_cald_helper():
    from oracle import cald_oracle as co
    g = np.load(os.path.join(GOLD, "cald_helper_augs.npz"))
    img, boxes = g["image"], torch.from_numpy(g["boxes"])
    fi, fb = co.horizontal_flip(img, boxes)
    assert np.array_equal(fi.numpy(), g["flip_image"]) and np.array_equal(fb.numpy(), g["flip_boxes"])
    ri, rb = co.resize(img, boxes, 0.8)
    assert np.array_equal(ri.numpy(), g["resize_image"]) and np.array_equal(rb.numpy(), g["resize_boxes"])
    oi, ob = co.rotate(img, boxes, 5)
    assert np.array_equal(oi.numpy(), g["rotate_image"])
    assert np.abs(ob.numpy() - g["rotate_boxes"]).max() <= 1e-4
    random.seed(int(g["cutout_seed"]))
    ci = co.cutout(img, boxes, 2)
    assert np.array_equal(ci.numpy(), g["cutout_image"])


def test_selection_matches_reference():
    from oracle import cald_oracle as co
    g = np.load(os.path.join(GOLD, "selection.npz"))
    hist = []
    for row in g["labels"]:
        h = [0] * g["cls"].shape[1]
        for l in row[row >= 0]:
            h[l - 1] += 1
        hist.append(h)
    new = co.select(g["uncertainty"], list(g["cls"]), list(g["subset"]), hist, int(g["budget"]))
    assert [int(v) for v in new] == [int(v) for v in g["new_labeled"]]


def test_api_select_matches_reference():
    """the product-side mirror (cald_b200.api.select / cls_kldiv) against the same fixture"""
    from cald_b200 import api
    g = np.load(os.path.join(GOLD, "selection.npz"))

    class LL:
        def __iter__(self):
            for row in g["labels"]:
                yield (None,), ({"labels": torch.from_numpy(row[row >= 0])},)
    new = api.select(g["uncertainty"], list(g["cls"]), list(g["subset"]), LL(), int(g["budget"]))
    assert [int(v) for v in new] == [int(v) for v in g["new_labeled"]]
    nm = api.select(g["uncertainty"], list(g["cls"]), list(g["subset"]), LL(), int(g["budget"]), mutual=False)
    assert [int(v) for v in nm] == [int(g["subset"][i]) for i in np.argsort(g["uncertainty"])[:int(g["budget"])]]
