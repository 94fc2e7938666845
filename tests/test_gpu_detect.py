"""Stage-wise and end-to-end parity of the CUDA detector forward vs the CPU oracle.

Oracle = oracle/frcnn_oracle.py (bit-exact with the unmodified reference on CPU, see
tests/test_oracle_golden.py).  Dense stages are compared within fp32-faithful tolerances
(split-bf16 operands: ~2e-5 relative per layer); discrete stages are compared as sets.
"""
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
    eng = Engine(depth=50, num_classes=21, min_size=320, max_size=512, debug=True, max_views_per_pass=4)
    eng.load_state_dict(w)
    cfg = fo.Cfg(50, 21, 320, 512)
    return eng, w, cfg, fo, synth


def _nhwc(t):
    return t[0].permute(1, 2, 0).contiguous().numpy()


def test_dense_stages_match_oracle(setup):
    eng, w, cfg, fo, synth = setup
    img = synth.synth_image(5, 200, 300)
    st = {}
    fo.forward(torch.from_numpy(img).permute(2, 0, 1).float().div(255), w, cfg, st)
    eng.detect([img])
    inp = eng.debug_fetch("input").reshape(_nhwc(st["input"]).shape)
    assert np.abs(inp - _nhwc(st["input"])).max() < 2e-6
    for i, name in enumerate(("c2", "c3", "c4", "c5")):
        want = _nhwc(st["c"][i])
        got = eng.debug_fetch(name).reshape(want.shape)
        err = np.abs(got - want).max() / np.abs(want).max()
        assert err < 2e-5, (name, err)
    for i, name in enumerate(("p2", "p3", "p4", "p5", "p6")):
        want = _nhwc(st["p"][i])
        got = eng.debug_fetch(name).reshape(want.shape)
        err = np.abs(got - want).max() / np.abs(want).max()
        assert err < 2e-5, (name, err)
    for l in range(5):
        lg, dl = st["rpn"][l]
        h, wd = st["p"][l].shape[-2:]
        got = eng.debug_fetch("rpn%d" % l).reshape(h, wd, 16)
        assert np.abs(got[..., :3].reshape(-1) - lg.numpy()).max() < 1e-4
        assert np.abs(got[..., 3:15].reshape(-1, 4) - dl.numpy()).max() < 1e-4


def test_roialign_matches_oracle_on_engine_inputs(setup):
    """MultiScaleRoIAlign stage (a14) in isolation: the oracle's RoIAlign (tv:ops/roi_align.py restated) fed with the
    ENGINE's own pyramid and proposals must reproduce the engine's pooled features.  The pyramid values are split-bf16
    numbers (exact in fp32), so the only differences are the order of the fp32 operations and the 2^-17 rounding of
    the pooled output to split bf16."""
    eng, w, cfg, fo, synth = setup
    img = synth.synth_image(8, 200, 300)
    st = {}
    fo.forward(torch.from_numpy(img).permute(2, 0, 1).float().div(255), w, cfg, st)
    eng.detect([img])
    n = int(eng.debug_fetch("proposal_count")[0])
    props = eng.debug_fetch("proposals").reshape(-1, 4)
    cap = props.shape[0]
    assert 0 < n <= cap
    feats = []
    for i, name in enumerate(("p2", "p3", "p4", "p5")):
        c, h, wd = st["p"][i].shape[1:]
        feats.append(torch.from_numpy(eng.debug_fetch(name).reshape(h, wd, c)).permute(2, 0, 1)[None].contiguous())
    want = fo.multiscale_roi_align(feats, torch.from_numpy(props[:n].copy())).numpy()      # n x 256 x 7 x 7
    got = eng.debug_fetch("pooled").reshape(cap, 7, 7, 256)[:n].transpose(0, 3, 1, 2)
    scale = np.abs(want).max()
    assert scale > 0
    assert np.abs(got - want).max() <= 2e-5 * scale, np.abs(got - want).max() / scale
    # rows beyond the proposal count are zero
    assert not eng.debug_fetch("pooled").reshape(cap, -1)[n:].any()


def test_proposals_match_oracle(setup):
    eng, w, cfg, fo, synth = setup
    img = synth.synth_image(6, 200, 300)
    st = {}
    fo.forward(torch.from_numpy(img).permute(2, 0, 1).float().div(255), w, cfg, st)
    eng.detect([img])
    n = int(eng.debug_fetch("proposal_count")[0])
    got = eng.debug_fetch("proposals").reshape(-1, 4)[:n]
    want = st["proposals"].numpy()
    assert n == len(want)
    # the same proposals in the same order (score-sorted, tv:rpn.py:242-297), up to fp32-level coordinate noise
    assert np.abs(want - got).max() < 5e-2


def test_detections_match_oracle(setup):
    eng, w, cfg, fo, synth = setup
    imgs = [synth.synth_image(i, 200, 300) for i in range(4)] + [synth.synth_image(9, 300, 200)]
    outs = eng.detect(imgs)
    for img, got in zip(imgs, outs):
        want = fo.forward(torch.from_numpy(img).permute(2, 0, 1).float().div(255), w, cfg)
        nw = len(want["scores"])
        assert len(got["scores"]) == nw
        k = nw
        # the whole list must agree in order, label, score and box
        assert np.array_equal(got["labels"][:k], want["labels"].numpy()[:k])
        assert np.abs(got["scores"][:k] - want["scores"].numpy()[:k]).max() < 1e-3
        assert np.abs(got["boxes"][:k] - want["boxes"].numpy()[:k]).max() < 5e-2
        assert np.abs(got["scores_cls"][:k] - want["scores_cls"].numpy()[:k]).max() < 1e-3
        assert np.abs(got["prob_max"][:k] - want["prob_max"].numpy()[:k]).max() < 1e-3


def test_simt_and_tcgen05_paths_agree(setup):
    eng, w, cfg, fo, synth = setup
    from cald_b200.engine import Engine, CONV_SIMT
    img = synth.synth_image(7, 160, 224)
    e2 = Engine(depth=50, num_classes=21, min_size=160, max_size=256, conv_impl=CONV_SIMT, debug=True)
    e2.load_state_dict(w)
    e1 = Engine(depth=50, num_classes=21, min_size=160, max_size=256, debug=True)
    e1.load_state_dict(w)
    a = e1.detect([img])[0]
    b = e2.detect([img])[0]
    pa, pb = e1.debug_fetch("p2"), e2.debug_fetch("p2")
    assert np.abs(pa - pb).max() / np.abs(pb).max() < 1e-4
    k = min(len(a["scores"]), len(b["scores"]), 10)
    assert np.abs(a["scores"][:k] - b["scores"][:k]).max() < 1e-3
