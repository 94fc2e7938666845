"""One-CTA vs CTA-pair (cta_group::2) conv kernel on the tensor-bound layer shapes (run on the GPU box).
usage: CALD_OP_TIMING=1 python tools/pair_micro.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cald_b200 import ops

os.environ["CALD_OP_TIMING"] = "1"
SHAPES = [(8, 200, 336, 256, 256, 3), (16, 100, 168, 128, 128, 3), (16, 50, 84, 256, 256, 3),
          (1, 1, 268800, 1024, 256, 1), (1, 1, 64000, 1024, 1024, 1), (1, 1, 1075200, 512, 128, 1)]
for (n, h, w, cin, cout, k) in SHAPES:
    rs = np.random.RandomState(0)
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    wt = (rs.standard_normal((cout, cin, k, k)) * 0.02).astype(np.float32)
    outs = {}
    for pair in ("0", "1"):
        os.environ["CALD_CTA2"] = pair
        sys.stderr.write("%dx%dx%d k%d %d->%d  pair=%s  " % (n, h, w, k, cin, cout, pair)); sys.stderr.flush()
        outs[pair] = ops.conv2d(x, wt, None, relu=True, impl=0, prec=0)
    d = np.abs(outs["0"] - outs["1"]).max() / np.abs(outs["0"]).max()
    sys.stderr.write("   max rel diff pair vs single: %.2e\n" % d)
