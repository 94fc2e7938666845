"""Diagnostic (GPU box): per-view detection parity of the engine vs the oracle for one image."""
import random
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cald_b200 import synth
from cald_b200.engine import Engine
from oracle import frcnn_oracle as fo, cald_oracle as co, pil_oracle as po

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 0
wnp = synth.planted_frcnn_weights(50, 21, 0)
w = {k: torch.from_numpy(v) for k, v in wnp.items()}
cfg = fo.Cfg(50, 21, 320, 512)
eng = Engine(depth=50, num_classes=21, min_size=320, max_size=512, debug=True)
eng.load_state_dict(wnp)
img = synth.synth_image(idx, 200, 300)
fwd = lambda x: fo.forward(x, w, cfg)
ref = fwd(co.to_tensor(img))
n = len(ref["scores"])
inds = torch.from_numpy(co.subsample_indices(n))
rb = ref["boxes"][inds]
random.seed(1000 + idx)
rects = co.cutout_rects(200, 300, rb, 2, random)
cut = img.copy()
for (l, t, r, b) in rects:
    cut[t:b, l:r] = 0
views = {"ref": img, "flip": np.ascontiguousarray(img[:, ::-1]), "cutout": cut,
         "resize": po.cald_resize_image(img, 0.8), "rotate": po.cald_rotate_image(img, 5)[0]}
for name, im in views.items():
    st = {}
    want = fo.forward(co.to_tensor(im), w, cfg, st)
    got = eng.detect([im])[0]
    nw, ng = len(want["scores"]), len(got["scores"])
    k = min(nw, ng)
    ws, gs = want["scores"].numpy(), got["scores"]
    same_lab = int((want["labels"].numpy()[:k] == got["labels"][:k]).sum())
    dsc = np.abs(ws[:k] - gs[:k])
    npc = int(eng.debug_fetch("proposal_count")[0])
    gp = eng.debug_fetch("proposals").reshape(-1, 4)[:npc]
    wp = st["proposals"].numpy()
    d = np.abs(wp[:, None, :] - gp[None, :, :]).max(-1)
    unmatched_w = int((d.min(1) > 0.05).sum())
    unmatched_g = int((d.min(0) > 0.05).sum())
    head = eng.debug_fetch("head")
    print("%-7s n_oracle=%d n_engine=%d labels_equal=%d/%d max|dscore|=%.2e (at %d) proposals: oracle %d engine %d unmatched %d/%d"
          % (name, nw, ng, same_lab, k, dsc.max() if k else 0, int(dsc.argmax()) if k else -1, len(wp), npc,
             unmatched_w, unmatched_g))
    for c in (1,):
        print("    class %d: oracle %s engine %s" % (c, np.round(ws[want["labels"].numpy() == c][:4], 4),
                                                    np.round(gs[got["labels"] == c][:4], 4)))

# ---- scoring path, one augmentation at a time: derive each view's class-max vector
from cald_b200 import api
names = ['flip', 'cut_out', 'smaller_resize', 'rotation']
refrow = np.array(co.class_max_vector(ref["scores"][inds], ref["labels"][inds], 21))
for a, vname in zip(names, ["flip", "cutout", "resize", "rotate"]):
    random.seed(1000 + idx)
    c, v = api.score_images(eng, [img], [a])
    augrow = v[0] * 2 - refrow
    want = fo.forward(co.to_tensor(views[vname]), w, cfg)
    wrow = np.array(co.class_max_vector(want["scores"], want["labels"], 21))
    bad = np.where(np.abs(augrow - wrow) > 1e-3)[0]
    print("%-14s consistency=%.6f  class-row max|diff|=%.2e  bad classes %s" % (a, c[0], np.abs(augrow - wrow).max(), bad.tolist()),
          [(int(b), round(float(augrow[b]), 4), round(float(wrow[b]), 4)) for b in bad])
