"""GPU box: error of the engine's dense stages against an fp64 run of the oracle forward, next to the error of the
fp32 CPU run (the reference's own arithmetic) against the same fp64 run.  Prints relative rms and the signed bias
(mean of the error projected on the sign of the exact value, relative to the mean magnitude) per stage.

    python tools/stage_error.py [h w min_size max_size]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cald_b200 import synth  # noqa: E402
from cald_b200.engine import Engine  # noqa: E402
from oracle import cald_oracle as co, frcnn_oracle as fo  # noqa: E402

h, w, mn, mx = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (375, 500, 600, 1000)
wnp = synth.planted_frcnn_weights(50, 21, 0)
w32 = {k: torch.from_numpy(v) for k, v in wnp.items()}
w64 = {k: v.double() for k, v in w32.items()}
cfg = fo.Cfg(50, 21, mn, mx)
img = synth.synth_image(5005, h, w)
st32, st64 = {}, {}
fo.forward(co.to_tensor(img), w32, cfg, st32)
fo.forward(co.to_tensor(img).double(), w64, cfg, st64)
eng = Engine(depth=50, num_classes=21, min_size=mn, max_size=mx, debug=True)
eng.load_state_dict(wnp)
eng.detect([img])


def nhwc(t):
    return t[0].permute(1, 2, 0).contiguous().numpy()


def stats(a, b):
    d = a.astype(np.float64) - b
    return np.sqrt((d ** 2).mean()) / np.sqrt((b ** 2).mean()), (d * np.sign(b)).mean() / np.abs(b).mean()


print("%-6s %28s %28s" % ("stage", "engine vs fp64 (rms, bias)", "CPU fp32 vs fp64 (rms, bias)"))
rows = [("c%d" % (i + 2), st32["c"][i], st64["c"][i]) for i in range(4)] + \
       [("p%d" % (i + 2), st32["p"][i], st64["p"][i]) for i in range(5)]
for name, a32, a64 in rows:
    want = nhwc(a64)
    got = eng.debug_fetch(name).reshape(want.shape)
    r1, b1 = stats(got, want)
    r2, b2 = stats(nhwc(a32), want)
    print("%-6s %14.2e %+13.2e %14.2e %+13.2e" % (name, r1, b1, r2, b2))
for l in range(5):
    lg64, dl64 = st64["rpn"][l]
    lg32, dl32 = st32["rpn"][l]
    hh, ww = st64["p"][l].shape[-2:]
    got = eng.debug_fetch("rpn%d" % l).reshape(hh, ww, 16)
    r1, b1 = stats(got[..., :3].reshape(-1), lg64.numpy().reshape(-1))
    r2, b2 = stats(lg32.numpy().reshape(-1), lg64.numpy().reshape(-1))
    d = np.abs(got[..., :3].reshape(-1) - lg64.numpy().reshape(-1))
    print("rpn%d   %14.2e %+13.2e %14.2e %+13.2e   objectness logits: max abs err engine %.2e, cpu32 %.2e" % (
        l, r1, b1, r2, b2, d.max(), np.abs(lg32.numpy().reshape(-1) - lg64.numpy().reshape(-1)).max()))
