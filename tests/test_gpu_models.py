"""Detector parity on the other model shapes the configs name: R101 backbone (cfg-4) and 91 classes (cfg-2/5)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("depth,nc", [(101, 21), (50, 91)])
def test_detect_matches_oracle(depth, nc):
    from cald_b200 import synth
    from cald_b200.engine import Engine
    from oracle import frcnn_oracle as fo
    w = synth.planted_frcnn_weights(depth, nc, 0)
    eng = Engine(depth=depth, num_classes=nc, min_size=320, max_size=512, debug=True)
    eng.load_state_dict(w)
    cfg = fo.Cfg(depth, nc, 320, 512)
    wt = {k: torch.from_numpy(v) for k, v in w.items()}
    for i in (0, 1):
        img = synth.synth_image(i, 200, 300)
        st = {}
        want = fo.forward(torch.from_numpy(img).permute(2, 0, 1).float().div(255), wt, cfg, st)
        got = eng.detect([img])[0]
        c5 = st["c"][3][0].permute(1, 2, 0).numpy()
        g5 = eng.debug_fetch("c5").reshape(c5.shape)
        assert np.abs(g5 - c5).max() / np.abs(c5).max() < 5e-4
        assert abs(len(got["scores"]) - len(want["scores"])) <= 1
        k = min(8, len(got["scores"]), len(want["scores"]))
        assert np.abs(got["scores"][:k] - want["scores"].numpy()[:k]).max() < 2e-3
        assert np.abs(got["boxes"][:k] - want["boxes"].numpy()[:k]).max() < 0.1
    eng.close()
