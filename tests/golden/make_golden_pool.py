"""Pool-level golden fixtures from the UNMODIFIED reference (build container only; needs /root/reference).

BASELINE.json configs[0]: 100-image VOC2007-shaped pool (375x500 with a portrait minority), Faster R-CNN R50-FPN,
nc = 21, min/max size 600/1000 (cald_train.py:340), augmentations F,C,D,R, bp 1.3 -- and the same pool through
retinanet_resnet50_fpn_cal (retinanet_cal.py:584-625).  For every pool this writes

  * ``consistency`` / ``cls``: what ``cald_train.get_uncertainty`` returns (python RNG reseeded per image, so one
    image's cutout draws do not move the next image's),
  * ``stream_*``: the same call with ONE seed for the whole pool (the way the reference really runs) and the next
    ``random.random()`` after it -- pins the RNG stream position over 100 images,
  * ``per_view`` and the detections of every view (ragged arrays): recorded by a pass-through wrapper around the
    unmodified model object while ``get_uncertainty`` ran, so a GPU-side mismatch can be traced to the view and stage,
  * ``selected``: the inline selection of cald_train.py:439-448 at budget 10 on these scores,
  * ``noise_*``: the unmodified reference run again with a different intra-op thread count -- the reference's own
    run-to-run wobble on this pool, the floor any other implementation's parity has to be read against.

Re-run:  python tests/golden/make_golden_pool.py [frcnn|retina|all]
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
warnings.filterwarnings("ignore")

from oracle import ref_stubs  # noqa: E402
from oracle import cald_oracle as co  # noqa: E402
from cald_b200 import synth  # noqa: E402

AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']
MIN_SIZE, MAX_SIZE, NC = 600, 1000, 21
N_POOL = int(os.environ.get("CALD_POOL_N", "100"))  # 100 is the committed fixture; smaller values are for trying the script
BUDGET = 10
KEYS = ("boxes", "scores", "labels", "prob_max")


def pool_spec():
    """(synth index, h, w): VOC2007-like 375x500 landscape images, every 7th one portrait."""
    return [(5000 + i, 500, 375) if i % 7 == 6 else (5000 + i, 375, 500) for i in range(N_POOL)]


def pool_images():
    return [synth.synth_image(i, h, w) for i, h, w in pool_spec()]


class Recorder:
    """Pass-through around the unmodified model: get_uncertainty calls .eval() and model([image]) (cald_train.py:96-187)."""

    def __init__(self, model):
        self.model = model
        self.calls = []

    def eval(self):
        self.model.eval()
        return self

    def __call__(self, images):
        out = self.model(images)
        self.calls.append({k: out[0][k].detach().clone() for k in KEYS + ("scores_cls",)})
        return out


class Replay:
    """forward_fn for the oracle loop that returns the recorded outputs in call order."""

    def __init__(self, calls):
        self.calls, self.pos = calls, 0

    def __call__(self, _):
        out = self.calls[self.pos]
        self.pos += 1
        return out


def labeled_loader(seed=3, n=25):
    rs = np.random.RandomState(seed)
    labeled = [[{"labels": torch.from_numpy(rs.randint(1, NC, rs.randint(1, 6)))}] for _ in range(n)]

    class LL:
        def __iter__(self):
            for t in labeled:
                yield (None,), tuple(t)
    rows = np.array([np.pad(t[0]["labels"].numpy(), (0, 6 - len(t[0]["labels"])), constant_values=-1) for t in labeled])
    return LL(), rows


def run_pool(tag, ct, model, imgs):
    t0 = time.time()

    def loader(seeds):
        class L:
            def __iter__(self):
                for k, im in enumerate(imgs):
                    if seeds is not None:
                        random.seed(seeds[k])
                    if k % 10 == 0:
                        print("  [%s] image %d  (%.0f s)" % (tag, k, time.time() - t0), flush=True)
                    yield (Image.fromarray(im),), (None,)
        return L()
    seeds = [7000 + k for k in range(len(imgs))]

    # ---- (A) per-image seeds, detections of every view recorded
    torch.set_num_threads(8)
    rec = Recorder(model)
    cons, cls = ct.get_uncertainty(rec, loader(seeds), AUGS, NC)
    cons, cls = np.array(cons, dtype=np.float64), np.array(cls, dtype=np.float64)
    # the oracle loop on the recorded outputs must land on the same numbers; its trace gives the per-view values
    rp = Replay(rec.calls)
    per_view = np.zeros((len(imgs), len(AUGS)))
    n_views = np.zeros(len(imgs), dtype=np.int64)
    for k, im in enumerate(imgs):
        random.seed(seeds[k])
        tr = {}
        p0 = rp.pos
        c, v = co.score_image(rp, im, AUGS, NC, 1.3, trace=tr)
        n_views[k] = rp.pos - p0
        assert float(c) == cons[k] and np.array_equal(np.asarray(v, dtype=np.float64), cls[k]), (k, c, cons[k])
        if "per_view" in tr:
            per_view[k] = tr["per_view"]
    assert rp.pos == len(rec.calls)
    off = np.zeros(len(rec.calls) + 1, dtype=np.int64)
    for i, c in enumerate(rec.calls):
        off[i + 1] = off[i] + len(c["scores"])
    det = {"det_" + key: np.concatenate([c[key].numpy() for c in rec.calls]) for key in KEYS}
    det["det_labels"] = det["det_labels"].astype(np.int8)
    if off[-1] > 100000:   # RetinaNet keeps up to 300 detections per class: scores and labels only, to stay small
        det.pop("det_boxes")
        det.pop("det_prob_max")
    print("  [%s] run A done: %.0f s, %d forwards, %d detections" % (tag, time.time() - t0, len(rec.calls), off[-1]), flush=True)

    # ---- selection at budget 10 with the reference's own inline code (cald_train.py:439-447)
    ll, label_rows = labeled_loader()
    subset = list(range(20000, 20000 + len(imgs)))
    arg = np.argsort(np.array(cons))
    cand = arg[:int(1.2 * BUDGET)]
    picked = ct.cls_kldiv(ll, [cls[i] for i in cand], BUDGET, 0)
    selected = np.array(list(torch.tensor(subset)[arg][picked].numpy()))

    # ---- (B) one seed for the whole pool: RNG stream position after 100 images
    random.seed(424242)
    cons_s, cls_s = ct.get_uncertainty(model, loader(None), AUGS, NC)
    tail = random.random()
    print("  [%s] run B done: %.0f s" % (tag, time.time() - t0), flush=True)

    # ---- (D) the reference against itself with another thread count
    torch.set_num_threads(3)
    cons_n, cls_n = ct.get_uncertainty(model, loader(seeds), AUGS, NC)
    torch.set_num_threads(8)
    cons_n, cls_n = np.array(cons_n, dtype=np.float64), np.array(cls_n, dtype=np.float64)
    d = np.abs(cons_n - cons)
    print("  [%s] run D done: %.0f s; reference(8 threads) vs reference(3 threads): max %.3e, >1e-3: %d, >1e-6: %d" % (
        tag, time.time() - t0, d.max(), int((d > 1e-3).sum()), int((d > 1e-6).sum())), flush=True)

    np.savez_compressed(os.path.join(os.environ.get("CALD_POOL_OUT", HERE), "pool_%s_nc21.npz" % tag), images=np.array(pool_spec()), seeds=np.array(seeds),
                        min_size=MIN_SIZE, max_size=MAX_SIZE, consistency=cons, cls=cls, per_view=per_view,
                        n_views=n_views, det_offsets=off, selected=selected, subset=np.array(subset), budget=BUDGET,
                        label_rows=label_rows, stream_seed=424242, stream_consistency=np.array(cons_s, dtype=np.float64),
                        stream_cls=np.array(cls_s, dtype=np.float64), stream_rng_tail=tail,
                        noise_consistency=cons_n, noise_cls=cls_n, noise_threads=np.array([8, 3]), **det)
    print("  [%s] written; consistency min %.4f median %.4f max %.4f" % (tag, cons.min(), np.median(cons), cons.max()))


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    ct = ref_stubs.load(bp=1.3)
    imgs = pool_images()
    if what in ("frcnn", "all"):
        fr = ref_stubs.frcnn_module()
        m = fr.fasterrcnn_resnet50_fpn_feature(num_classes=NC, pretrained_backbone=False, min_size=MIN_SIZE,
                                               max_size=MAX_SIZE)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.planted_frcnn_weights(50, NC, 0).items()}, strict=True)
        run_pool("frcnn_r50", ct, m.eval(), imgs)
    if what in ("retina", "all"):
        rm = ref_stubs.retinanet_module()
        m = rm.retinanet_resnet50_fpn_cal(num_classes=NC, pretrained_backbone=False, min_size=MIN_SIZE, max_size=MAX_SIZE)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.planted_retinanet_weights(NC, 0).items()}, strict=True)
        run_pool("retina_r50", ct, m.eval(), imgs)


if __name__ == "__main__":
    main()
