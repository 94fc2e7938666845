"""Diagnostic (GPU box): which view and which stage is behind a full-size pool image whose score differs from the
reference fixture (tests/golden/fullsize_*_nc91.npz).  The oracle (bit-pinned to the unmodified reference) re-scores the
image on the host with a trace; the engine scores it with debug views; detections are compared view by view.

    python tools/diag_fullsize.py retina 2
"""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cald_b200 import api, synth  # noqa: E402
from cald_b200.engine import Engine, ARCH_FRCNN, ARCH_RETINANET  # noqa: E402
from oracle import cald_oracle as co  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "retina"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']
NC = 91
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                         "fullsize_%s_nc91.npz" % ("frcnn_r50" if kind == "frcnn" else "retina_r50")))
idx, h, w = (int(v) for v in g["images"][k])
img = synth.synth_image(idx, h, w)
if kind == "frcnn":
    from oracle import frcnn_oracle as mo
    wnp = synth.planted_frcnn_weights(50, NC, 0)
else:
    from oracle import retina_oracle as mo
    wnp = synth.planted_retinanet_weights(NC, 0, cls_bias_shift=float(g["retina_shift"]))
wt = {n: torch.from_numpy(v) for n, v in wnp.items()}
cfg = mo.Cfg(50, NC, 800, 1333)
torch.set_num_threads(max(1, min(32, len(os.sched_getaffinity(0)))))
random.seed(int(g["seeds"][k]))
tr = {}
c_or, v_or = co.score_image(lambda x: mo.forward(x, wt, cfg), img, AUGS, NC, 1.3, trace=tr)
print("image %d (%dx%d): reference fixture %.6f, oracle here %.6f" % (k, h, w, g["consistency"][k], c_or))
eng = Engine(depth=50, num_classes=NC, min_size=800, max_size=1333, debug=True,
             arch_id=ARCH_FRCNN if kind == "frcnn" else ARCH_RETINANET)
eng.load_state_dict(wnp)
random.seed(int(g["seeds"][k]))
c_en, v_en = api.score_images(eng, [img], AUGS)
pv = eng.last_per_view(1, len(AUGS))[0]
print("engine %.6f; per view (F, C, D, R): engine %s oracle %s" % (
    c_en[0], np.round(pv, 6), np.round(np.array(tr["per_view"], dtype=np.float64), 6)))
views = eng.debug_views(1, len(AUGS))[0]
want = [tr["ref_full"]] + list(tr["dets"])
for name, ge, wo in zip(["reference"] + AUGS, views, want):
    ws, wl = wo["scores"].numpy(), wo["labels"].numpy()
    gs, gl = ge["scores"], ge["labels"]
    line = "%-15s detections: oracle %d engine %d" % (name, len(ws), len(gs))
    if len(ws) == len(gs) and np.array_equal(wl, gl):
        line += "; same labels, max |score diff| %.2e, max |box diff| %.2e" % (
            np.abs(ws - gs).max() if len(ws) else 0.0, np.abs(wo["boxes"].numpy() - ge["boxes"]).max() if len(ws) else 0.0)
    else:
        # per class: the first class whose list differs, and the first differing row in it
        for c in range(NC):
            a, b = ws[wl == c], gs[gl == c]
            if len(a) != len(b) or (len(a) and np.abs(a - b).max() > 1e-4):
                j = next((i for i in range(min(len(a), len(b))) if abs(a[i] - b[i]) > 1e-4), min(len(a), len(b)))
                line += "; class %d: oracle %d rows engine %d, first difference at row %d (oracle %s engine %s)" % (
                    c, len(a), len(b), j, np.round(a[j:j + 2], 6), np.round(b[j:j + 2], 6))
                wb, gb = wo["boxes"].numpy()[wl == c], ge["boxes"][gl == c]
                if j < len(wb):
                    # the row the two lists disagree on: its best IoU with an earlier (better-scored) kept box of its class
                    def iou(p, q):
                        iw = max(0.0, min(p[2], q[2]) - max(p[0], q[0])); ih = max(0.0, min(p[3], q[3]) - max(p[1], q[1]))
                        inter = iw * ih
                        return inter / ((p[2] - p[0]) * (p[3] - p[1]) + (q[2] - q[0]) * (q[3] - q[1]) - inter)
                    extra = wb[j] if len(a) > len(b) or (j < len(gb) and a[j] > b[j]) else gb[j]
                    prev = wb[:j]
                    if len(prev):
                        best = max(iou(extra.astype(np.float64), p.astype(np.float64)) for p in prev)
                        line += "; best IoU of that box with an earlier kept box of the class: %.7f (NMS threshold 0.5)" % best
                break
    print(line)
