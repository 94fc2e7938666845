"""Generate the RetinaNet golden fixtures with the UNMODIFIED reference (build container only).

Needs /root/reference (imported read-only through oracle/ref_stubs.py): runs
``detection.retinanet_cal.retinanet_resnet50_fpn_cal`` and ``cald_train.get_uncertainty`` on CPU with the
planted weights and writes tests/golden/retina_r50_nc21_*.npz.  Also checks that oracle/retina_oracle.py
reproduces the reference bit for bit on these inputs (the oracle's pin).  Re-run:
    python tests/golden/make_golden_retina.py
"""
import os
import random
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
from oracle import retina_oracle as ro  # noqa: E402
from oracle import cald_oracle as co  # noqa: E402
from cald_b200 import synth  # noqa: E402

AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']
MIN_SIZE, MAX_SIZE, NC = 320, 512, 21
IMAGES = [(0, 200, 300), (1, 200, 300), (2, 200, 300), (20, 300, 200), (4, 167, 250), (3, 200, 300)]  # (index, h, w)


def main():
    torch.set_num_threads(8)
    ct = ref_stubs.load(bp=1.3)
    rm = ref_stubs.retinanet_module()
    w = synth.planted_retinanet_weights(NC, 0)
    m = rm.retinanet_resnet50_fpn_cal(num_classes=NC, pretrained_backbone=False, min_size=MIN_SIZE, max_size=MAX_SIZE)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    m.eval()
    imgs = [synth.synth_image(i, h, wd) for i, h, wd in IMAGES]
    cfg = ro.Cfg(50, NC, MIN_SIZE, MAX_SIZE)

    import torchvision.transforms.functional as F
    det = {}
    for k, img in enumerate(imgs):
        with torch.no_grad():
            out = m([F.to_tensor(Image.fromarray(img))])[0]
        mine = ro.forward(co.to_tensor(img), w, cfg)
        for key in ("boxes", "scores", "labels", "prob_max", "scores_cls"):
            a, b = out[key].numpy(), mine[key].numpy()
            assert a.shape == b.shape, (k, key, a.shape, b.shape)
            d = float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) if a.size else 0.0
            assert d == 0.0, ("oracle differs from the reference", k, key, d)
            det["%d_%s" % (k, key)] = a
        sc = out["scores"].numpy()
        print("image %d: %d detections, %d classes, score margin to 0.05: %.2e" % (
            k, len(sc), len(np.unique(out["labels"].numpy())), float(np.abs(sc - 0.05).min()) if len(sc) else -1))
    np.savez_compressed(os.path.join(HERE, "retina_r50_nc21_detect.npz"), images=np.array(IMAGES),
                        min_size=MIN_SIZE, max_size=MAX_SIZE, **det)

    class Loader:
        def __iter__(self):
            for k, im in enumerate(imgs):
                random.seed(2000 + k)
                yield (Image.fromarray(im),), (None,)
    cons, cls = ct.get_uncertainty(m, Loader(), AUGS, NC)
    want_c, want_v = co.get_uncertainty(lambda x: ro.forward(x, w, cfg), imgs, AUGS, NC, 1.3,
                                        seeds=[2000 + k for k in range(len(imgs))])
    assert np.array_equal(np.array(cons, dtype=np.float64), np.array(want_c, dtype=np.float64)), (cons, want_c)
    assert np.array_equal(np.array(cls, dtype=np.float64), np.array(want_v, dtype=np.float64))
    np.savez_compressed(os.path.join(HERE, "retina_r50_nc21_uncertainty.npz"), images=np.array(IMAGES),
                        consistency=np.array(cons, dtype=np.float64), cls=np.array(cls, dtype=np.float64),
                        seeds=np.array([2000 + k for k in range(len(imgs))]))
    print("consistency", cons)
    print("retina fixtures written; oracle == reference on all of them")


if __name__ == "__main__":
    main()
