"""Accuracy of the tcgen05 conv variants against an fp64 reference (run on the GPU box)."""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cald_b200 import ops

def case(n, h, w, cin, cout, k, relu_in=True):
    rs = np.random.RandomState(1)
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    if relu_in:
        x = np.maximum(x, 0)            # activations after a ReLU: positive mean, like the real layers
    wt = (rs.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    ref64 = F.conv2d(xt.double(), torch.from_numpy(wt).double(), padding=k // 2).permute(0, 2, 3, 1).numpy()
    ref32 = F.conv2d(xt, torch.from_numpy(wt), padding=k // 2).permute(0, 2, 3, 1).numpy()
    scale = np.abs(ref64).max()
    def rep(tag, y):
        d = y.astype(np.float64) - ref64
        print("  %-28s max %.2e  rms %.2e  mean(signed, rel to |ref|) %.2e" % (
            tag, np.abs(d).max() / scale, np.sqrt((d ** 2).mean()) / scale,
            (d * np.sign(ref64)).mean() / np.abs(ref64).mean()))
    print("conv %dx%dx%d k%d %d->%d  (K = %d)" % (n, h, w, k, cin, cout, cin * k * k))
    rep("torch CPU fp32", ref32)
    for tag, kw in (("x3 default (xsep / chunk)", dict(prec=0)), ("x3 kc=0", dict(prec=0, kc=0)),
                    ("x3 BN256 3-MMA nochunk", dict(prec=0, block_n=256, kc=0)), ("simt fp32 fma", dict(prec=0, impl=1)),
                    ("bf16 single pass", dict(prec=1))):
        if "BN256" in tag and cout % 256:
            continue
        impl = kw.pop("impl", 0)
        rep(tag, ops.conv2d(x, wt, None, relu=False, impl=impl, **kw))

case(1, 24, 32, 256, 256, 3)
case(1, 24, 32, 64, 64, 3)
case(1, 16, 16, 2048, 512, 1)
case(1, 12, 12, 512, 512, 3)
case(1, 1, 256, 12544, 256, 1)
