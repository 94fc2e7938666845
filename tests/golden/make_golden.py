"""Generate the golden fixtures with the UNMODIFIED reference (build container only).

Needs /root/reference (imported read-only through oracle/ref_stubs.py).  Writes small .npz files next to this
script; they are committed and are the pins of the oracle (tests/test_oracle_golden.py) and of the engine
(tests/test_gpu_golden.py).  Re-run:  python tests/golden/make_golden.py
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
from cald_b200 import synth  # noqa: E402

AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']
MIN_SIZE, MAX_SIZE, NC = 320, 512, 21
IMAGES = [(0, 200, 300), (1, 200, 300), (2, 200, 300), (20, 300, 200), (4, 167, 250)]  # (index, h, w)


def build_model(fr, w):
    m = fr.fasterrcnn_resnet50_fpn_feature(num_classes=NC, pretrained_backbone=False, min_size=MIN_SIZE,
                                           max_size=MAX_SIZE)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    return m.eval()


def main():
    torch.set_num_threads(8)
    ct = ref_stubs.load(bp=1.3)
    fr = ref_stubs.frcnn_module()
    hp = ref_stubs.helper_module()
    w = synth.planted_frcnn_weights(50, NC, 0)
    m = build_model(fr, w)
    imgs = [synth.synth_image(i, h, wd) for i, h, wd in IMAGES]

    # ---- (1) detector forward: task_model([F.to_tensor(img)])[0]
    det = {}
    import torchvision.transforms.functional as F
    for k, img in enumerate(imgs):
        with torch.no_grad():
            out = m([F.to_tensor(Image.fromarray(img))])[0]
        for key in ("boxes", "scores", "labels", "props", "prob_max", "scores_cls"):
            det["%d_%s" % (k, key)] = out[key].numpy()
    np.savez_compressed(os.path.join(HERE, "frcnn_r50_nc21_detect.npz"),
                        images=np.array(IMAGES), min_size=MIN_SIZE, max_size=MAX_SIZE, **det)

    # ---- (2) cald_train.get_uncertainty, python RNG reseeded per image
    class Loader:
        def __iter__(self):
            for k, im in enumerate(imgs):
                random.seed(1000 + k)
                yield (Image.fromarray(im),), (None,)
    cons, cls = ct.get_uncertainty(m, Loader(), AUGS, NC)
    # RNG stream position after a run WITHOUT per-image reseeding (pins the draw consumption)
    random.seed(4242)

    class Loader2:
        def __iter__(self):
            for im in imgs[:3]:
                yield (Image.fromarray(im),), (None,)
    cons2, _ = ct.get_uncertainty(m, Loader2(), ['cut_out'], NC)
    tail = random.random()
    np.savez_compressed(os.path.join(HERE, "frcnn_r50_nc21_uncertainty.npz"),
                        images=np.array(IMAGES), consistency=np.array(cons, dtype=np.float64),
                        cls=np.array(cls, dtype=np.float64), seeds=np.array([1000 + k for k in range(len(imgs))]),
                        cutout_only_consistency=np.array(cons2, dtype=np.float64), cutout_only_seed=4242,
                        cutout_only_rng_tail=tail)

    # ---- (3) augmentation helpers of cald/cald_helper.py on a small image with known boxes
    small = synth.synth_image(77, 97, 133)
    boxes = torch.tensor([[10.0, 12.0, 60.0, 70.0], [40.5, 5.25, 120.0, 90.0], [0.0, 0.0, 133.0, 97.0]])
    pil = Image.fromarray(small)
    fimg, fbox = hp.HorizontalFlip(pil, boxes)
    rimg, rbox = hp.resize(pil, boxes, 0.8)
    oimg, obox = hp.rotate(pil, boxes, 5)
    random.seed(9)
    cimg = hp.cutout(pil, boxes, None, 2)
    np.savez_compressed(os.path.join(HERE, "cald_helper_augs.npz"), image=small, boxes=boxes.numpy(),
                        flip_image=fimg.numpy(), flip_boxes=fbox.numpy(), resize_image=rimg.numpy(),
                        resize_boxes=rbox.numpy(), rotate_image=oimg.numpy(), rotate_boxes=obox.numpy(),
                        cutout_image=cimg.numpy(), cutout_seed=9)

    # ---- (4) cls_kldiv / selection (cald_train.py:234-271, 439-448)
    rs = np.random.RandomState(5)
    n, budget = 60, 10
    unc = rs.uniform(0, 1, n)
    cls_rows = [rs.uniform(0, 1, NC - 1) * (rs.uniform(0, 1, NC - 1) > 0.6) for _ in range(n)]
    cls_rows[3] = np.zeros(NC - 1)
    labeled = [[{"labels": torch.from_numpy(rs.randint(1, NC, rs.randint(1, 6)))}] for _ in range(25)]

    class LL:
        def __iter__(self):
            for t in labeled:
                yield (None,), tuple(t)
    arg = np.argsort(np.array(unc))
    cand = arg[:int(1.2 * budget)]
    picked = ct.cls_kldiv(LL(), [cls_rows[i] for i in cand], budget, 0)
    subset = list(range(500, 500 + n))
    new = list(torch.tensor(subset)[arg][picked].numpy())
    np.savez_compressed(os.path.join(HERE, "selection.npz"), uncertainty=unc, cls=np.array(cls_rows),
                        labels=np.array([np.pad(t[0]["labels"].numpy(), (0, 6 - len(t[0]["labels"])), constant_values=-1)
                                         for t in labeled]),
                        budget=budget, picked=np.array(picked), new_labeled=np.array(new), subset=np.array(subset))
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print("  ", f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
