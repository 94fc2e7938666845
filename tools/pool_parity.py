"""GPU box: score the 100-image fixture pools (tests/golden/pool_*_nc21.npz, written by the UNMODIFIED reference) with the
engine, print the per-image error distribution / selection equality / RNG stream position, and dump the engine's
per-view detections so that every outlier can be traced to the view and stage that flipped (tools/pool_diagnose.py, which
runs in the build container and costs no GPU time).

    python tools/pool_parity.py [frcnn|retina] [out.npz]
"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cald_b200 import api, synth  # noqa: E402
from cald_b200.engine import Engine, ARCH_FRCNN, ARCH_RETINANET  # noqa: E402

AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']
GOLD = os.path.join(ROOT, "tests", "golden")


def load_pool(kind):
    tag = "frcnn_r50" if kind == "frcnn" else "retina_r50"
    g = np.load(os.path.join(GOLD, "pool_%s_nc21.npz" % tag))
    imgs = [synth.synth_image(int(i), int(h), int(w)) for i, h, w in g["images"]]
    return g, imgs


def make_engine(kind, g, debug=True, **kw):
    if kind == "frcnn":
        w = synth.planted_frcnn_weights(50, 21, 0)
        arch = ARCH_FRCNN
    else:
        w = synth.planted_retinanet_weights(21, 0)
        arch = ARCH_RETINANET
    eng = Engine(depth=50, num_classes=21, min_size=int(g["min_size"]), max_size=int(g["max_size"]), debug=debug,
                 arch_id=arch, **kw)
    eng.load_state_dict(w)
    return eng


class LabeledLoader:
    def __init__(self, rows):
        self.rows = rows

    def __iter__(self):
        import torch
        for r in self.rows:
            yield (None,), ({"labels": torch.from_numpy(r[r >= 0])},)


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "frcnn"
    out = sys.argv[2] if len(sys.argv) > 2 else None
    g, imgs = load_pool(kind)
    eng = make_engine(kind, g, max_views_per_pass=40)
    n = len(imgs)
    A = len(AUGS)
    cons, cls, per_view = np.zeros(n), np.zeros((n, 20)), np.zeros((n, A))
    counts, dets = [], {"boxes": [], "scores": [], "labels": [], "prob_max": []}
    t0 = time.time()
    for k, im in enumerate(imgs):
        random.seed(int(g["seeds"][k]))
        c, v = api.score_images(eng, [im], AUGS)
        cons[k], cls[k] = c[0], v[0]
        pv = eng.last_per_view(1, A)
        per_view[k] = pv[0] if len(pv) else 0
        views = eng.debug_views(1, A)[0]
        for vw in views:
            counts.append(len(vw["scores"]))
            for key in dets:
                dets[key].append(vw[key].copy())
    dt = time.time() - t0
    err = np.abs(cons - g["consistency"])
    cerr = np.abs(cls - g["cls"]).max(axis=1)
    print("[%s] %d images in %.1f s (one image per call, debug copies on)" % (kind, n, dt))
    print("  |consistency - reference|: median %.2e  p90 %.2e  p99 %.2e  max %.2e;  > 1e-3: %d images %s" % (
        np.median(err), np.percentile(err, 90), np.percentile(err, 99), err.max(), int((err > 1e-3).sum()),
        np.where(err > 1e-3)[0].tolist()))
    print("  |class vector - reference| max per image: median %.2e  max %.2e;  > 1e-3: %d images %s" % (
        np.median(cerr), cerr.max(), int((cerr > 1e-3).sum()), np.where(cerr > 1e-3)[0].tolist()))
    pv_err = np.abs(per_view - g["per_view"])
    print("  per-view consistency: > 1e-3 in %d of %d views" % (int((pv_err > 1e-3).sum()), pv_err.size))
    # selection at the fixture's budget with the reference's inline code path (api.select == cald_train.py:439-447)
    sel = api.select(list(cons), [c for c in cls], list(g["subset"]), LabeledLoader(g["label_rows"]), int(g["budget"]))
    same = sorted(int(v) for v in sel) == sorted(int(v) for v in g["selected"])
    print("  selected set identical: %s  (engine %s | reference %s)" % (
        same, sorted(int(v) for v in sel), sorted(int(v) for v in g["selected"])))
    k = int(g["budget"])
    print("  top-%d by score identical: %s" % (k, set(np.argsort(cons)[:k]) == set(np.argsort(g["consistency"])[:k])))
    # ---- one seed for the whole pool, one call: RNG stream position after 100 images, batched scoring
    eng2 = make_engine(kind, g, debug=False, max_views_per_pass=64)
    random.seed(int(g["stream_seed"]))
    t0 = time.time()
    c2, v2 = api.score_images(eng2, imgs, AUGS)
    dt2 = time.time() - t0
    tail = random.random()
    e2 = np.abs(np.array(c2) - g["stream_consistency"])
    print("  one-seed pool run (batched, %.2f s): RNG tail equal: %s;  > 1e-3: %d images %s  max %.2e" % (
        dt2, tail == float(g["stream_rng_tail"]), int((e2 > 1e-3).sum()), np.where(e2 > 1e-3)[0].tolist(), e2.max()))
    print("  arena peak %.2f GB" % (eng2.arena_peak() / 2 ** 30))
    if out:
        off = np.concatenate([[0], np.cumsum(counts)])
        np.savez_compressed(out, consistency=cons, cls=cls, per_view=per_view, det_offsets=off,
                            stream_consistency=np.array(c2), stream_tail=tail,
                            **{"det_" + k2: (np.concatenate(v) if len(v) else np.zeros(0)) for k2, v in dets.items()})
        print("  wrote", out)


if __name__ == "__main__":
    main()
