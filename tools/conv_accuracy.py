"""Accuracy of the tcgen05 conv variants against an fp64 reference (run on the GPU box).

For every shape: relative rms and signed bias of (a) torch CPU fp32, (b) the engine's default launch, (c) forced
chunked accumulation with kc = 2 / 4 / 8 k-blocks per chunk, (d) the fp32-FMA SIMT checker on the same split operands
(the representation floor).  With CALD_RZ_BETA=0 the signed bias of (b) divided by the number of truncating accumulates
is the per-accumulate shrink `beta` that ConvParams::acc_gain compensates; the last column re-runs (b) with the beta
given on the command line.

    python tools/conv_accuracy.py [beta]
"""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cald_b200 import ops

BETA = sys.argv[1] if len(sys.argv) > 1 else None


def run(x, wt, **env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return ops.conv2d(x, wt, None, relu=False, **({"impl": 1} if env.pop("SIMT", None) else {}))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def case(n, h, w, cin, cout, k, relu_in=True, seed=1):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    if relu_in:
        x = np.maximum(x, 0)            # activations after a ReLU: positive mean, like the real layers
    wt = (rs.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    ref64 = F.conv2d(xt.double(), torch.from_numpy(wt).double(), padding=k // 2).permute(0, 2, 3, 1).numpy()
    ref32 = F.conv2d(xt, torch.from_numpy(wt), padding=k // 2).permute(0, 2, 3, 1).numpy()

    def rep(tag, y, adds=None):
        d = y.astype(np.float64) - ref64
        rms = np.sqrt((d ** 2).mean()) / np.sqrt((ref64 ** 2).mean())
        bias = (d * np.sign(ref64)).mean() / np.abs(ref64).mean()
        extra = "  bias / accumulate %.3e" % (bias / adds) if adds else ""
        print("  %-34s rms %.2e  bias %+.2e%s" % (tag, rms, bias, extra))
        return bias
    K = cin * k * k
    adds = K // 16
    print("conv %dx%dx%d k%d %d->%d  (K = %d, %d accumulates, relu_in=%s)" % (n, h, w, k, cin, cout, K, adds, relu_in))
    rep("torch CPU fp32", ref32)
    rep("simt fp32 fma on split operands", run(x, wt, SIMT=1))
    rep("engine default, beta 0", run(x, wt, CALD_RZ_BETA=0), adds if K // 64 <= 40 else 32)
    rep("unchunked, beta 0", run(x, wt, CALD_RZ_BETA=0, CALD_KC=0), adds)
    for kc in (2, 4, 8):
        rep("chunked kc=%d, beta 0" % kc, run(x, wt, CALD_RZ_BETA=0, CALD_KC=kc, CALD_CHUNK_ABOVE_KB=0), kc * 4)
    if BETA:
        rep("engine default, beta %s" % BETA, run(x, wt, CALD_RZ_BETA=BETA))
        rep("unchunked, beta %s" % BETA, run(x, wt, CALD_RZ_BETA=BETA, CALD_KC=0))


case(1, 24, 32, 64, 64, 3)
case(1, 24, 32, 256, 256, 3)
case(1, 24, 32, 256, 256, 3, relu_in=False)
case(1, 16, 16, 1024, 256, 1)
case(1, 16, 16, 2048, 512, 1)
case(1, 12, 12, 512, 512, 3)
case(1, 1, 256, 12544, 256, 1)
case(1, 24, 32, 64, 256, 1)
case(1, 24, 32, 256, 64, 1)
