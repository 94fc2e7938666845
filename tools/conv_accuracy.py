"""Accuracy of the tcgen05 conv kernel against an fp64 reference (run on the GPU box).

For every shape: relative rms and signed bias (mean of the error projected on the sign of the exact value, relative to
the mean magnitude) of
  (a) torch CPU fp32 -- the reference's own arithmetic,
  (b) fp32 FMAs on the engine's split-half operands (SIMT checker) -- the representation floor,
  (c) the tensor-core kernel without truncation compensation (CALD_RZ_C=0): its bias divided by the number of MMA
      accumulates is the per-accumulate shrink of the truncating fp32 accumulate,
  (d) the tensor-core kernel with the weight pre-compensation (RzPlan, conv_host.cuh) for a few values of c,
  (e) forced chunked accumulation (kc = 2 / 4 k-blocks per chunk) for comparison.

    python tools/conv_accuracy.py [c ...]          default c values: 2.8e-8 3.2e-8 3.6e-8
"""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cald_b200 import ops

CS = sys.argv[1:] or ["2.8e-8", "3.2e-8", "3.6e-8"]


def run(x, wt, simt=False, **env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return ops.conv2d(x, wt, None, relu=False, impl=1 if simt else 0)
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
        print("  %-36s rms %.2e  bias %+.2e%s" % (tag, rms, bias, extra))
    K = cin * k * k
    adds = K // 16
    print("conv %dx%dx%d k%d %d->%d  (K = %d, %d accumulates, relu_in=%s)" % (n, h, w, k, cin, cout, K, adds, relu_in))
    rep("torch CPU fp32", ref32)
    rep("fp32 FMA on split operands (SIMT)", run(x, wt, simt=True))
    rep("tensor core, c = 0", run(x, wt, CALD_RZ_C=0), adds if K // 64 <= 40 else 32)
    for c in CS:
        rep("tensor core, c = %s" % c, run(x, wt, CALD_RZ_C=c))
    for kc in (2, 4):
        rep("chunked kc=%d, c = %s" % (kc, CS[len(CS) // 2]), run(x, wt, CALD_RZ_C=CS[len(CS) // 2], CALD_KC=kc, CALD_CHUNK_ABOVE_KB=0))


case(1, 24, 32, 64, 64, 3)
case(1, 24, 32, 256, 256, 3)
case(1, 24, 32, 256, 256, 3, relu_in=False)
case(1, 16, 16, 1024, 256, 1)
case(1, 16, 16, 2048, 512, 1)
case(1, 12, 12, 512, 512, 3)
case(1, 1, 256, 12544, 256, 1)
case(1, 24, 32, 64, 256, 1)
case(1, 24, 32, 256, 64, 1)
