"""Micro-benchmark of the tcgen05 conv kernel variants on one layer shape (run on the GPU box).
usage: CALD_OP_TIMING=1 python tools/conv_micro.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cald_b200 import ops

def run(tag, n, h, w, cin, cout, k, **kw):
    rs = np.random.RandomState(0)
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    wt = (rs.standard_normal((cout, cin, k, k)) * 0.02).astype(np.float32)
    sys.stderr.write("%-58s " % tag); sys.stderr.flush()
    ops.conv2d(x, wt, None, relu=True, impl=0, **kw)

for (n, h, w, cin, cout, k) in [(8, 200, 336, 256, 256, 3), (8, 200, 336, 64, 64, 3), (8, 200, 336, 256, 256, 1)]:
    shape = "%dx%dx%d k%d %d->%d" % (n, h, w, k, cin, cout)
    run(shape + " x3 BN128 chunk8", n, h, w, cin, cout, k, prec=0, block_n=0, kc=8)
    run(shape + " x3 BN128 nochunk", n, h, w, cin, cout, k, prec=0, block_n=128, kc=0)
    run(shape + " x3 BN64 nochunk", n, h, w, cin, cout, k, prec=0, block_n=64, kc=0)
    run(shape + " single-pass half BN128", n, h, w, cin, cout, k, prec=1, block_n=128, kc=0)
    if cout >= 256:
        run(shape + " single-pass half BN256", n, h, w, cin, cout, k, prec=1, block_n=256, kc=0)
    run(shape + " single-pass half BN64", n, h, w, cin, cout, k, prec=1, block_n=64, kc=0)
