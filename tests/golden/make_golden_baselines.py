"""Golden fixtures for the two detection-only baseline scorers, produced by the UNMODIFIED reference functions
lt_c_train.get_uncertainty (lt_c_train.py:105-121) and ls_c_train.get_uncertainty (ls_c_train.py:108-155)
(build container only; needs /root/reference).  Also asserts that oracle/cald_oracle.py reproduces them exactly.
    python tests/golden/make_golden_baselines.py
"""
import os
import sys
import warnings

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_stubs  # noqa: E402
from oracle import frcnn_oracle as fo  # noqa: E402
from oracle import cald_oracle as co  # noqa: E402
from cald_b200 import synth  # noqa: E402

MIN_SIZE, MAX_SIZE, NC = 320, 512, 21
IMAGES = [(0, 200, 300), (1, 200, 300), (2, 200, 300), (20, 300, 200), (4, 167, 250)]


def main():
    torch.set_num_threads(8)
    ref_stubs.load()
    fr = ref_stubs.frcnn_module()
    ltc = ref_stubs.baseline_module("lt_c_train")
    lsc = ref_stubs.baseline_module("ls_c_train")
    w = synth.planted_frcnn_weights(50, NC, 0)
    m = fr.fasterrcnn_resnet50_fpn_feature(num_classes=NC, pretrained_backbone=False, min_size=MIN_SIZE,
                                           max_size=MAX_SIZE)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    m.eval()
    imgs = [synth.synth_image(i, h, wd) for i, h, wd in IMAGES]
    cfg = fo.Cfg(50, NC, MIN_SIZE, MAX_SIZE)
    wt = {k: torch.from_numpy(v) for k, v in w.items()}
    fwd = lambda x: fo.forward(x, wt, cfg)  # noqa: E731

    import torchvision.transforms.functional as F
    # LT/C: the script's loader yields ToTensor()'d images (lt_c_train.py:109-111)
    t_loader = [((F.to_tensor(Image.fromarray(im)),), (None,)) for im in imgs]
    u_ltc = ltc.get_uncertainty(m, t_loader)
    assert np.array_equal(np.array(u_ltc), np.array(co.ltc_uncertainty(fwd, imgs))), "LT/C oracle != reference"
    # LS+C: PIL loader, torch generator seeded once for the whole pool
    p_loader = [((Image.fromarray(im),), (None,)) for im in imgs]
    torch.manual_seed(77)
    u_lsc = lsc.get_uncertainty(m, p_loader)
    torch.manual_seed(77)
    mine = co.lsc_stability(fwd, imgs)
    assert np.array_equal(np.array(u_lsc, dtype=np.float64), np.array(mine, dtype=np.float64)), "LS+C oracle != reference"
    np.savez_compressed(os.path.join(HERE, "baseline_scorers.npz"), images=np.array(IMAGES), min_size=MIN_SIZE,
                        max_size=MAX_SIZE, ltc=np.array(u_ltc, dtype=np.float64),
                        lsc=np.array(u_lsc, dtype=np.float64), lsc_seed=77)
    print("LT/C", u_ltc)
    print("LS+C", [float(v) for v in u_lsc])
    print("baseline fixtures written; oracle == reference")


if __name__ == "__main__":
    main()
