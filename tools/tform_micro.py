"""Micro-benchmark (GPU box): the layer1 3x3 64 -> 64 conv on the pixel-major kernels (one-CTA BLOCK_N = 64, CTA pair)
and on the transposed-role kernel (igemm_t.cuh).   CALD_OP_TIMING=1 python tools/tform_micro.py [views]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cald_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rs = np.random.RandomState(0)
x = rs.standard_normal((n, 200, 336, 64)).astype(np.float32)
wt = (rs.standard_normal((64, 64, 3, 3)) * 0.04).astype(np.float32)
b = rs.standard_normal(64).astype(np.float32)
os.environ["CALD_OP_TIMING"] = "1"
outs = {}
for tag, env in (("one CTA, BLOCK_N = 64", dict(CALD_TFORM="0", CALD_CTA2="0")),
                 ("CTA pair, BLOCK_N = 64", dict(CALD_TFORM="0", CALD_CTA2="1")),
                 ("transposed roles, N = 256", dict(CALD_TFORM="1"))):
    for k in ("CALD_TFORM", "CALD_CTA2"):
        os.environ.pop(k, None)
    os.environ.update(env)
    sys.stderr.write("%dx200x336 k3 64->64  %-28s " % (n, tag))
    sys.stderr.flush()
    outs[tag] = ops.conv2d(x, wt, b, relu=True, prec=0, impl=0)
ref = outs["one CTA, BLOCK_N = 64"]
for tag, o in outs.items():
    print("%-28s max |diff to the one-CTA kernel| %.3e" % (tag, np.abs(o - ref).max()))
