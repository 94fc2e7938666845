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


def test_retina_oracle_matches_reference_fixture():
    """oracle/retina_oracle.py against retinanet_cal.py's own output (tests/golden/make_golden_retina.py)"""
    from cald_b200 import synth
    from oracle import retina_oracle as ro
    from oracle import cald_oracle as co
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    w = {k: torch.from_numpy(v) for k, v in synth.planted_retinanet_weights(21, 0).items()}
    cfg = ro.Cfg(50, 21, 320, 512)
    g = np.load(os.path.join(GOLD, "retina_r50_nc21_detect.npz"))
    for k in (0, 1, 4):
        idx, h, wd = g["images"][k]
        out = ro.forward(co.to_tensor(synth.synth_image(int(idx), int(h), int(wd))), w, cfg)
        for key in ("boxes", "scores", "labels", "prob_max", "scores_cls"):
            want = g["%d_%s" % (k, key)]
            got = out[key].numpy()
            assert got.shape == want.shape, (k, key)
            if want.size:
                assert np.abs(got.astype(np.float64) - want).max() <= 2e-5, (k, key)
    u = np.load(os.path.join(GOLD, "retina_r50_nc21_uncertainty.npz"))
    imgs = [synth.synth_image(int(i), int(h), int(wd)) for i, h, wd in u["images"][:2]]
    cons, cls = co.get_uncertainty(lambda x: ro.forward(x, w, cfg), imgs, AUGS, 21, 1.3,
                                   seeds=[int(s) for s in u["seeds"][:2]])
    assert np.abs(np.array(cons) - u["consistency"][:2]).max() <= 1e-5
    assert np.abs(np.array(cls) - u["cls"][:2]).max() <= 1e-5


def test_retina_anchor_tables():
    """known answers for the RetinaNet anchor generator (retinanet_cal.py:347-351, tv:anchor_utils.py:58-75)"""
    from oracle import retina_oracle as ro
    assert ro.anchor_sizes()[0] == (32, 40, 50) and ro.anchor_sizes()[4] == (512, 645, 812)
    base = ro.cell_anchors((32, 40, 50)).numpy()
    assert base.shape == (9, 4)
    # ratio 0.5 (wide), scale 32: w = 32 / sqrt(.5) = 45.25, h = 22.63 -> +-23, +-11
    assert base[0].tolist() == [-23.0, -11.0, 23.0, 11.0]
    # ratio 1, scale 40 -> +-20
    assert base[4].tolist() == [-20.0, -20.0, 20.0, 20.0]
    a = ro.grid_anchors((64, 96), [(8, 12), (4, 6), (2, 3), (1, 2), (1, 1)])
    assert [len(x) for x in a] == [864, 216, 54, 18, 9]
    assert a[0][9].tolist() == [8 - 23.0, -11.0, 8 + 23.0, 11.0]   # second cell, stride 8


def test_baseline_scorers_oracle_matches_reference_fixture(model):
    """LT/C and LS+C restatements against lt_c_train / ls_c_train.get_uncertainty (make_golden_baselines.py)"""
    fwd, synth = model
    from oracle import cald_oracle as co
    g = np.load(os.path.join(GOLD, "baseline_scorers.npz"))
    imgs = [synth.synth_image(int(i), int(h), int(w)) for i, h, w in g["images"]]
    got = co.ltc_uncertainty(fwd, imgs[:3])
    assert np.abs(np.array(got) - g["ltc"][:3]).max() <= 1e-5
    torch.manual_seed(int(g["lsc_seed"]))
    got = co.lsc_stability(fwd, imgs[:1])
    assert abs(got[0] - g["lsc"][0]) <= 1e-4


def test_oracle_matches_pool_fixture_and_timing_mode_is_bit_identical():
    """The 100-image pool fixture (tests/golden/make_golden_pool.py, unmodified reference): the oracle reproduces the
    first images exactly, also in its timing mode (NMS / RoIAlign through torchvision's CPU kernels, what bench.py's
    CPU arm runs), and the recorded per-view detections are what the oracle's forward returns."""
    from cald_b200 import synth
    from oracle import cald_oracle as co
    from oracle import frcnn_oracle as fo
    g = np.load(os.path.join(GOLD, "pool_frcnn_r50_nc21.npz"))
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    w = {k: torch.from_numpy(v) for k, v in synth.planted_frcnn_weights(50, 21, 0).items()}
    cfg = fo.Cfg(50, 21, int(g["min_size"]), int(g["max_size"]))
    fwd = lambda x: fo.forward(x, w, cfg)  # noqa: E731
    idx, h, wd = (int(v) for v in g["images"][0])
    img = synth.synth_image(idx, h, wd)
    results = []
    for fast in (False, True):
        fo.USE_TORCHVISION_OPS = fast
        try:
            random.seed(int(g["seeds"][0]))
            tr = {}
            c, v = co.score_image(fwd, img, AUGS, 21, 1.3, trace=tr)
        finally:
            fo.USE_TORCHVISION_OPS = False
        results.append((float(c), np.asarray(v), tr))
    assert results[0][0] == results[1][0] and np.array_equal(results[0][1], results[1][1])
    assert abs(results[0][0] - float(g["consistency"][0])) <= 5e-6
    assert np.abs(results[0][1] - g["cls"][0]).max() <= 5e-6
    assert np.abs(np.array(results[0][2]["per_view"]) - g["per_view"][0]).max() <= 5e-6
    # recorded detections of the reference view == the oracle's reference forward
    a, b = int(g["det_offsets"][0]), int(g["det_offsets"][1])
    ref = results[0][2]["ref_full"]
    assert len(ref["scores"]) == b - a
    assert np.abs(ref["scores"].numpy() - g["det_scores"][a:b]).max() <= 2e-5
    assert np.array_equal(ref["labels"].numpy(), g["det_labels"][a:b].astype(np.int64))
