"""The detection-only baseline scorers on the engine vs fixtures produced by the unmodified
lt_c_train.get_uncertainty / ls_c_train.get_uncertainty (tests/golden/make_golden_baselines.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _Model:
    """what api.engine_for needs from a torchvision detector: state_dict() and transform.min/max_size"""
    def __init__(self, w, mn, mx):
        self._w = {k: torch.from_numpy(v) for k, v in w.items()}
        self.transform = type("T", (), {"min_size": (mn,), "max_size": mx})()

    def state_dict(self):
        return self._w


@pytest.fixture(scope="module")
def setup():
    from cald_b200 import synth
    g = np.load(os.path.join(GOLD, "baseline_scorers.npz"))
    imgs = [synth.synth_image(int(i), int(h), int(w)) for i, h, w in g["images"]]
    model = _Model(synth.planted_frcnn_weights(50, 21, 0), int(g["min_size"]), int(g["max_size"]))
    return g, imgs, model


def test_lt_c_matches_reference_fixture(setup):
    g, imgs, model = setup
    from cald_b200 import api
    # the script's loader yields ToTensor()'d float images (lt_c_train.py:109-111)
    loader = [((torch.from_numpy(im).permute(2, 0, 1).float().div(255),), (None,)) for im in imgs]
    got = np.array(api.lt_c_uncertainty(model, loader))
    print("LT/C engine", got, "reference", g["ltc"])
    assert np.abs(got - g["ltc"]).max() <= 1e-3


def test_ls_c_matches_reference_fixture(setup):
    g, imgs, model = setup
    from cald_b200 import api
    loader = [((im,), (None,)) for im in imgs]
    torch.manual_seed(int(g["lsc_seed"]))
    got = np.array(api.ls_c_uncertainty(model, loader))
    print("LS+C engine", got, "reference", g["lsc"])
    err = np.abs(got - g["lsc"])
    assert err.max() <= 1e-3, err
    assert np.array_equal(np.argsort(got), np.argsort(g["lsc"]))  # same selection order


def test_lt_c_rejects_retinanet():
    from cald_b200 import synth
    from cald_b200._lib import CaldError
    from cald_b200.engine import Engine, ARCH_RETINANET
    eng = Engine(depth=50, num_classes=21, min_size=160, max_size=256, arch_id=ARCH_RETINANET)
    eng.load_state_dict(synth.planted_retinanet_weights(21, 0))
    with pytest.raises(CaldError):
        eng.score_ltc([synth.synth_image(0, 120, 160)])


def test_engine_model_serves_the_evaluation_loop(setup):
    """detection/engine.py's evaluators call model(list_of_float_tensors): same detections as the fixture"""
    g, imgs, model = setup
    from cald_b200 import EngineModel
    det = np.load(os.path.join(GOLD, "frcnn_r50_nc21_detect.npz"))
    m = EngineModel(model).eval()
    from cald_b200 import synth
    ims = [synth.synth_image(int(i), int(h), int(w)) for i, h, w in det["images"][:2]]
    outs = m([torch.from_numpy(im).permute(2, 0, 1).float().div(255) for im in ims])
    for k, o in enumerate(outs):
        assert set(o) >= {"boxes", "labels", "scores"} and o["labels"].dtype == torch.int64
        assert len(o["scores"]) == len(det["%d_scores" % k])
        n = len(o["scores"])
        assert np.array_equal(o["labels"].numpy()[:n], det["%d_labels" % k][:n])
        assert np.abs(o["scores"].numpy()[:n] - det["%d_scores" % k][:n]).max() < 1e-3
        assert np.abs(o["boxes"].numpy()[:n] - det["%d_boxes" % k][:n]).max() < 5e-2
