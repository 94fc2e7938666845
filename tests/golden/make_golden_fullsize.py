"""Full-size pool fixtures from the UNMODIFIED reference (build container only; needs /root/reference).

BASELINE.json configs[1] / configs[2] at their real shape: 800x1333 COCO-shaped images (every 5th one portrait,
1333x800), nc = 91, min/max size 800/1333, augmentations F,C,D,R, bp 1.3 -- Faster R-CNN R50-FPN
(frcnn_la.py via fasterrcnn_resnet50_fpn_feature) and RetinaNet R50-FPN (retinanet_cal.py:584-625), scored by
``cald_train.get_uncertainty`` itself.  The CPU reference needs ~20 s per image at this size, so the pools are small
(24 / 12 images); the 100-image pools of make_golden_pool.py carry the statistics, these pin the full-size shape.

Per pool: ``consistency`` / ``cls`` with python's RNG reseeded per image, ``stream_*`` = the same call with ONE seed for
the whole pool plus the next ``random.random()`` after it, and ``selected`` = the inline selection of
cald_train.py:439-448 at budget 4.

Re-run:  python tests/golden/make_golden_fullsize.py [frcnn|retina|all]
"""
import os
import random
import sys
import time
import warnings

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

from oracle import ref_stubs  # noqa: E402
from cald_b200 import synth  # noqa: E402
from make_golden_pool import labeled_loader  # noqa: E402

AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']
MIN_SIZE, MAX_SIZE, NC = 800, 1333, 91
BUDGET = 4
RETINA_SHIFT = -11.0   # classifier bias shift of the planted RetinaNet at nc = 91 (keeps the detection lists in the hundreds)
N = {"frcnn_r50": int(os.environ.get("CALD_FULL_N_FRCNN", "24")), "retina_r50": int(os.environ.get("CALD_FULL_N_RETINA", "12"))}


def pool_spec(n):
    return [(8000 + i, 1333, 800) if i % 5 == 4 else (8000 + i, 800, 1333) for i in range(n)]


def run_pool(tag, ct, model):
    spec = pool_spec(N[tag])
    imgs = [synth.synth_image(i, h, w) for i, h, w in spec]
    t0 = time.time()

    def loader(seeds):
        class L:
            def __iter__(self):
                for k, im in enumerate(imgs):
                    if seeds is not None:
                        random.seed(seeds[k])
                    print("  [%s] image %d  (%.0f s)" % (tag, k, time.time() - t0), flush=True)
                    yield (Image.fromarray(im),), (None,)
        return L()
    seeds = [9000 + k for k in range(len(imgs))]
    torch.set_num_threads(8)
    cons, cls = ct.get_uncertainty(model, loader(seeds), AUGS, NC)
    cons, cls = np.array(cons, dtype=np.float64), np.array(cls, dtype=np.float64)
    ll, label_rows = labeled_loader()
    subset = list(range(30000, 30000 + len(imgs)))
    arg = np.argsort(np.array(cons))
    cand = arg[:int(1.2 * BUDGET)]
    picked = ct.cls_kldiv(ll, [cls[i] for i in cand], BUDGET, 0)
    selected = np.array(list(torch.tensor(subset)[arg][picked].numpy()))
    random.seed(515151)
    cons_s, cls_s = ct.get_uncertainty(model, loader(None), AUGS, NC)
    tail = random.random()
    np.savez_compressed(os.path.join(os.environ.get("CALD_POOL_OUT", HERE), "fullsize_%s_nc91.npz" % tag),
                        images=np.array(spec), seeds=np.array(seeds), min_size=MIN_SIZE, max_size=MAX_SIZE,
                        consistency=cons, cls=cls, selected=selected, subset=np.array(subset), budget=BUDGET,
                        label_rows=label_rows, stream_seed=515151,
                        stream_consistency=np.array(cons_s, dtype=np.float64),
                        stream_cls=np.array(cls_s, dtype=np.float64), stream_rng_tail=tail, retina_shift=RETINA_SHIFT)
    print("  [%s] written after %.0f s; consistency min %.4f median %.4f max %.4f" % (
        tag, time.time() - t0, cons.min(), np.median(cons), cons.max()))


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    ct = ref_stubs.load(bp=1.3)
    if what in ("frcnn", "all"):
        fr = ref_stubs.frcnn_module()
        m = fr.fasterrcnn_resnet50_fpn_feature(num_classes=NC, pretrained_backbone=False, min_size=MIN_SIZE,
                                               max_size=MAX_SIZE)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.planted_frcnn_weights(50, NC, 0).items()}, strict=True)
        run_pool("frcnn_r50", ct, m.eval())
    if what in ("retina", "all"):
        rm = ref_stubs.retinanet_module()
        m = rm.retinanet_resnet50_fpn_cal(num_classes=NC, pretrained_backbone=False, min_size=MIN_SIZE, max_size=MAX_SIZE)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.planted_retinanet_weights(NC, 0, cls_bias_shift=RETINA_SHIFT).items()}, strict=True)
        run_pool("retina_r50", ct, m.eval())


if __name__ == "__main__":
    main()
