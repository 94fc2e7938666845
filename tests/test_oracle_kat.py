"""Known-answer and property tests of the oracle's small pieces (no reference needed)."""
import numpy as np
import pytest
import torch


def test_pil_restatement_is_bit_exact_with_pillow():
    from PIL import Image
    from oracle import pil_oracle as po
    rs = np.random.RandomState(0)
    for (h, w) in [(97, 133), (64, 96), (150, 100)]:
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        pil = Image.fromarray(img)
        assert np.array_equal(np.asarray(pil.resize((int(w * 0.8), int(h * 0.8)), Image.BILINEAR)),
                              po.cald_resize_image(img, 0.8))
        assert np.array_equal(np.asarray(pil.rotate(5, expand=True)), po.rotate_nearest_expand(img, 5))
        assert np.array_equal(np.asarray(pil.rotate(5, expand=True).resize((w, h))), po.cald_rotate_image(img, 5)[0])


def test_js_divergence_matches_scipy():
    import scipy.stats
    from oracle import cald_oracle as co
    rs = np.random.RandomState(1)
    for t in range(300):
        n = int(rs.choice([21, 91]))
        p = rs.dirichlet(np.ones(n) * 0.3).astype(np.float32)
        q = (rs.dirichlet(np.ones(n) * 0.3) if t % 2 else rs.uniform(0, 1, n)).astype(np.float32)
        m = (p + q) / 2
        js = 0.5 * scipy.stats.entropy(p, m) + 0.5 * scipy.stats.entropy(q, m)
        js = max(js, 0)
        assert co.js_divergence(p, q) == js


def test_subsample_quirks():
    from oracle import cald_oracle as co
    assert list(co.subsample_indices(40)) == list(range(40))
    idx = co.subsample_indices(41)
    assert len(idx) == 50 and idx[0] == 0 and idx[-1] == 40 and len(set(idx)) < 50  # duplicates when 40 < n < 50
    assert list(np.round([0.5, 1.5, 2.5])) == [0, 2, 2]  # banker's rounding is part of the contract
    idx = co.subsample_indices(100)
    assert len(set(idx)) == 50 and idx[-1] == 99


def test_class_vector_wraps_label_zero():
    from oracle import cald_oracle as co
    v = co.class_max_vector([0.5, 0.7, 0.2], [1, 0, 1], 21)
    assert v[0] == 0.5 and v[-1] == 0.7 and sum(1 for x in v if x) == 2


def test_pair_consistency_hand_computed():
    from oracle import cald_oracle as co
    # one reference box identical to the single detection, identical class vectors -> IoU 1, JS 0
    p = torch.tensor([[0.1, 0.9]])
    ref = {"scores_cls": p, "prob_max": torch.tensor([0.9])}
    det = {"boxes": torch.tensor([[0., 0., 10., 10.]]), "scores_cls": p, "prob_max": torch.tensor([0.9])}
    v = co.pair_consistency(ref, torch.tensor([[0., 0., 10., 10.]]), det, 1.3)
    assert abs(v - abs(1 + 0.5 * 1.8 - 1.3)) < 1e-6
    # disjoint boxes: IoU 0 -> |0.9 - 1.3| = 0.4
    v = co.pair_consistency(ref, torch.tensor([[20., 20., 30., 30.]]), det, 1.3)
    assert abs(v - 0.4) < 1e-6
    # empty augmented prediction contributes 0.0
    assert co.pair_consistency(ref, torch.tensor([[0., 0., 1., 1.]]), {"boxes": torch.zeros((0, 4))}, 1.3) == 0.0


def test_nms_matches_torchvision():
    import torchvision
    from oracle import frcnn_oracle as fo
    rs = np.random.RandomState(2)
    for _ in range(5):
        xy = rs.uniform(0, 100, (300, 2))
        wh = rs.uniform(5, 40, (300, 2))
        b = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        s = rs.uniform(0, 1, 300).astype(np.float32)
        for thr in (0.5, 0.7):
            want = torchvision.ops.nms(torch.from_numpy(b), torch.from_numpy(s), thr).numpy()
            assert np.array_equal(fo.nms_numpy(b, s, thr), want)


def test_roi_align_matches_torchvision():
    import torchvision
    from oracle import frcnn_oracle as fo
    rs = np.random.RandomState(3)
    feat = torch.from_numpy(rs.standard_normal((1, 16, 25, 42)).astype(np.float32))
    rois = torch.tensor([[3.0, 4.0, 80.0, 60.0], [0.0, 0.0, 160.0, 96.0], [100.2, 50.7, 101.0, 51.0],
                         [150.0, 90.0, 200.0, 120.0]])
    want = torchvision.ops.roi_align(feat, [rois], output_size=7, spatial_scale=0.25, sampling_ratio=2)
    got = fo.roi_align_level(feat[0], rois, 0.25)
    assert np.abs(got.numpy() - want.numpy()).max() < 1e-5


def test_cutout_consumes_four_draws_per_try():
    import random
    from oracle import cald_oracle as co

    class Counting:
        n = 0

        def uniform(self, a, b):
            self.n += 1
            return random.uniform(a, b)
    random.seed(0)
    rng = Counting()
    boxes = torch.tensor([[10., 10., 90., 90.]])
    rects = co.cutout_rects(100, 100, boxes, 2, rng)
    assert rng.n % 4 == 0 and rng.n >= 4 * len(rects) and len(rects) <= 2
