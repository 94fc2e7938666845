"""Build container (CPU only, needs /root/reference): seconds per image of the UNMODIFIED reference
``cald_train.get_uncertainty`` next to the oracle port that bench.py times as ``cpu_baseline`` / ``--impl reference``
(kind "port") on the GPU box, where /root/reference does not exist.  Same images, weights, thread count; results must
also be identical.  Writes a markdown table (profiles/r02_cpu_port_vs_reference.md).

    python tools/cpu_port_vs_reference.py [out.md]
"""
import os
import random
import sys
import time
import warnings

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
from oracle import ref_stubs, cald_oracle as co, frcnn_oracle as fo  # noqa: E402
from cald_b200 import synth  # noqa: E402

AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']
CASES = [("cfg-1 shape: FRCNN R50-FPN nc=21, 500x375, 600/1000", 21, 375, 500, 600, 1000, 6),
         ("cfg-2 shape: FRCNN R50-FPN nc=91, 1333x800, 800/1333", 91, 800, 1333, 800, 1333, 3)]


def main():
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
    threads = len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    ct = ref_stubs.load(bp=1.3)
    fr = ref_stubs.frcnn_module()
    print("# CPU arm calibration: oracle port vs the unmodified reference\n", file=out)
    print("Host: %d threads, torch %s.  One warm-up image, then n timed images; python RNG seeded identically.\n" % (
        threads, torch.__version__), file=out)
    print("| workload | n | unmodified `cald_train.get_uncertainty` s/image | oracle port, numpy NMS / RoIAlign (parity "
          "tests) s/image | oracle port, torchvision NMS / RoIAlign kernels (bench.py CPU arm) s/image | max abs score "
          "difference vs the reference (both ports) |", file=out)
    print("|---|---|---|---|---|---|", file=out)
    for name, nc, h, w, mn, mx, n in CASES:
        wnp = synth.planted_frcnn_weights(50, nc, 0)
        wt = {k: torch.from_numpy(v) for k, v in wnp.items()}
        m = fr.fasterrcnn_resnet50_fpn_feature(num_classes=nc, pretrained_backbone=False, min_size=mn, max_size=mx)
        m.load_state_dict(wt, strict=True)
        m.eval()
        cfg = fo.Cfg(50, nc, mn, mx)
        imgs = [synth.synth_image(900 + i, h, w) for i in range(n + 1)]

        class L:
            def __init__(self, ims):
                self.ims = ims

            def __iter__(self):
                for im in self.ims:
                    yield (Image.fromarray(im),), (None,)
        random.seed(1)
        ct.get_uncertainty(m, L(imgs[:1]), AUGS, nc)
        t = time.time()
        ref, _ = ct.get_uncertainty(m, L(imgs[1:]), AUGS, nc)
        t_ref = (time.time() - t) / n
        fwd = lambda x: fo.forward(x, wt, cfg)  # noqa: E731
        res = {}
        for fast in (False, True):
            fo.USE_TORCHVISION_OPS = fast
            random.seed(1)
            co.score_image(fwd, imgs[0], AUGS, nc, 1.3)
            t = time.time()
            port = [co.score_image(fwd, im, AUGS, nc, 1.3)[0] for im in imgs[1:]]
            res[fast] = ((time.time() - t) / n,
                         float(np.abs(np.array(ref, dtype=np.float64) - np.array(port, dtype=np.float64)).max()))
        fo.USE_TORCHVISION_OPS = False
        print("| %s | %d | %.2f | %.2f (x%.2f) | %.2f (x%.2f) | %.1e / %.1e |" % (
            name, n, t_ref, res[False][0], res[False][0] / t_ref, res[True][0], res[True][0] / t_ref, res[False][1],
            res[True][1]), file=out)
        out.flush()


if __name__ == "__main__":
    main()
