"""The drop-in entry points themselves (cald_b200.get_uncertainty with a reference-style model and loader of PIL images),
the host wrapper's RNG bookkeeping around empty reference predictions, weight reloads, and the RetinaNet class with more
candidates than the fast path's list holds."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']


class ModelLike:
    """What get_uncertainty needs of the reference's FRCNN_Feature / RetinaNet object: .state_dict() (torchvision keys)
    and .transform.min_size / .max_size (frcnn_la.py:224-235)."""

    class _T:
        pass

    def __init__(self, weights, min_size, max_size):
        self.sd = {k: torch.from_numpy(v.copy()) for k, v in weights.items()}
        self.transform = self._T()
        self.transform.min_size = (min_size,)
        self.transform.max_size = max_size

    def state_dict(self):
        return self.sd

    def eval(self):
        return self


class PILLoader:
    """DataLoader(dataset_aug, batch_size=1, collate_fn=utils.collate_fn) (cald_train.py:370-371): yields
    (tuple_of_PIL_images, tuple_of_targets)."""

    def __init__(self, images):
        self.images = images

    def __iter__(self):
        from PIL import Image
        for im in self.images:
            yield (Image.fromarray(im),), ({"boxes": None},)


def test_get_uncertainty_dropin_streams_the_pool_like_the_reference():
    """cald_b200.get_uncertainty(model, loader_of_PIL_images, augs, num_cls) over the 100-image cfg-1 pool with ONE seed
    for the whole pool, against what the unmodified reference returned for the same call: scores, class vectors and
    the position of python's RNG stream afterwards (tests/golden/make_golden_pool.py, run B)."""
    import cald_b200
    from cald_b200 import synth
    g = np.load(os.path.join(GOLD, "pool_frcnn_r50_nc21.npz"))
    imgs = [synth.synth_image(int(i), int(h), int(w)) for i, h, w in g["images"]]
    model = ModelLike(synth.planted_frcnn_weights(50, 21, 0), int(g["min_size"]), int(g["max_size"]))
    random.seed(int(g["stream_seed"]))
    cons, cls = cald_b200.get_uncertainty(model, PILLoader(imgs), AUGS, 21)
    tail = random.random()
    assert isinstance(cons, list) and isinstance(cons[0], float) and len(cons) == len(imgs)
    assert cls[0].shape == (20,) and cls[0].dtype == np.float64
    err = np.abs(np.array(cons) - g["stream_consistency"])
    print("stream run: > 1e-3 at", np.where(err > 1e-3)[0].tolist(), "max", err.max())
    # one cutout decision that flips moves the stream for every later image, so an identical tail means that all
    # 100 x (up to 50 tries) accept/reject decisions were the reference's
    assert tail == float(g["stream_rng_tail"])
    assert (err <= 1e-3).mean() >= 0.99, np.where(err > 1e-3)[0]
    assert np.abs(np.array(cls) - g["stream_cls"]).max(axis=1).mean() <= 1e-3
    cald_b200.close_engines()


def test_rng_streams_rewind_over_an_empty_reference_prediction():
    """The reference draws no swap / noise randomness for an image whose reference view has no detection
    (cald_train.py:118-121 precedes 127-157).  A flat grey image has none under the planted weights."""
    from cald_b200 import api, synth
    from cald_b200.engine import Engine
    from oracle import cald_oracle as co, frcnn_oracle as fo
    wnp = synth.planted_frcnn_weights(50, 21, 0)
    w = {k: torch.from_numpy(v) for k, v in wnp.items()}
    cfg = fo.Cfg(50, 21, 320, 512)
    eng = Engine(depth=50, num_classes=21, min_size=320, max_size=512)
    eng.load_state_dict(wnp)
    imgs = [synth.synth_image(0, 200, 300), np.full((200, 300, 3), 128, np.uint8), synth.synth_image(1, 200, 300)]
    for augs in (['flip', 'ga'], ['color_swap', 'sp']):
        random.seed(5)
        torch.manual_seed(5)
        want, want_cls = co.get_uncertainty(lambda x: fo.forward(x, w, cfg), imgs, augs, 21, 1.3)
        want_tail = (random.random(), float(torch.rand(1)))
        random.seed(5)
        torch.manual_seed(5)
        got, got_cls = api.score_images(eng, imgs, augs)
        tail = (random.random(), float(torch.rand(1)))
        assert want[1] == 0.0 and got[1] == 0.0 and not np.any(got_cls[1])
        assert tail == want_tail, (augs, tail, want_tail)
        assert np.abs(np.array(got) - np.array(want, dtype=np.float64)).max() <= 2e-3, (augs, got, want)
    eng.close()


def test_engine_for_reloads_changed_weights_and_shares_one_engine():
    from cald_b200 import api, synth
    api.close_engines()
    w = synth.planted_frcnn_weights(50, 21, 0)
    m1 = ModelLike(w, 320, 512)
    img = synth.synth_image(2, 200, 300)
    ev = api.EngineModel(m1)
    a = ev([torch.from_numpy(img).permute(2, 0, 1).float().div(255)])[0]
    # "retraining": in-place update of one tensor bumps its version counter
    m1.sd["roi_heads.box_predictor.cls_score.bias"][1:] += 1.5
    b = ev([img])[0]
    assert len(b["scores"]) != len(a["scores"]) or not np.allclose(a["scores"].numpy(), b["scores"].numpy())
    # a second model object with the same configuration shares the engine, and each sees its own weights
    m2 = ModelLike(w, 320, 512)
    e2 = api.engine_for(m2, 21)
    assert e2 is ev.engine
    c = api.EngineModel(m2)([img])[0]
    assert np.array_equal(c["scores"].numpy(), a["scores"].numpy())
    d = ev([img])[0]   # back to m1: its (modified) weights are loaded again
    assert np.array_equal(d["scores"].numpy(), b["scores"].numpy())
    api.close_engines()


@pytest.mark.parametrize("shift,min_over", [(1.0, 4096), (3.0, 8192)])
def test_retinanet_class_with_more_candidates_than_the_fast_list(shift, min_over):
    """A lightly trained RetinaNet puts far more than 4096 anchors of one class above the 0.05 threshold; the reference
    has no limit there (retinanet_cal.py:436-463).  The engine scans such classes in windows; the detections must be the
    oracle's, row for row (class-major order, 300 per class)."""
    from cald_b200 import synth
    from cald_b200.engine import Engine, ARCH_RETINANET
    from oracle import cald_oracle as co, retina_oracle as ro
    wnp = synth.planted_retinanet_weights(21, 0, cls_bias_shift=shift)
    w = {k: torch.from_numpy(v) for k, v in wnp.items()}
    cfg = ro.Cfg(50, 21, 320, 512)
    img = synth.synth_image(3, 200, 300)
    st = {}
    want = ro.forward(co.to_tensor(img), w, cfg, st)
    per_class = (torch.sigmoid(st["cls_logits"].reshape(-1, 21)) > 0.05).sum(0)
    assert int(per_class.max()) > min_over, per_class.max()
    eng = Engine(depth=50, num_classes=21, min_size=320, max_size=512, arch_id=ARCH_RETINANET)
    eng.load_state_dict(wnp)
    got = eng.detect([img])[0]
    wl, gl = want["labels"].numpy(), got["labels"]
    # same number of detections per class (up to a box or two on an NMS / threshold edge)
    cw, cg = np.bincount(wl, minlength=21), np.bincount(gl, minlength=21)
    assert np.abs(cw - cg).max() <= 2, (cw, cg)
    # row-for-row agreement on every class whose count matches
    matched = total = 0
    for c in range(21):
        if cw[c] != cg[c] or cw[c] == 0:
            continue
        a, b = want["scores"].numpy()[wl == c], got["scores"][gl == c]
        ba, bb = want["boxes"].numpy()[wl == c], got["boxes"][gl == c]
        ok = (np.abs(a - b) < 1e-3) & (np.abs(ba - bb).max(axis=1) < 0.1)
        matched += int(ok.sum())
        total += len(a)
    assert total >= 0.8 * len(wl) and matched >= 0.98 * total, (matched, total, len(wl))
    eng.close()
