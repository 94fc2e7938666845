"""Calibrate the planted-weight recipe (build container only; needs /root/reference).

Runs the UNMODIFIED reference detector with the un-centred planted weights on a
few synthetic calibration images, records the mean box-head feature (input of
``roi_heads.box_predictor.cls_score``) and stores ``-W_cls @ mean`` per class in
``cald_b200/planted_calib.json``.  ``synth.planted_frcnn_weights`` adds that
vector to the classifier bias so per-class logits are centred on real features.
Usage: python tools/calibrate_planted.py <depth> <num_classes> [height width min max]
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_stubs  # noqa: E402
from cald_b200 import synth  # noqa: E402

warnings.filterwarnings("ignore")


def main():
    depth, nc = int(sys.argv[1]), int(sys.argv[2])
    h, w, mn, mx = (int(v) for v in sys.argv[3:7]) if len(sys.argv) > 6 else (375, 500, 600, 1000)
    ref_stubs.load()
    fr = ref_stubs.frcnn_module()
    from torchvision.models.detection.backbone_utils import resnet_fpn_backbone
    wts = synth.planted_frcnn_weights(depth, nc, 0, calib=False)
    bb = resnet_fpn_backbone(backbone_name="resnet%d" % depth, weights=None)
    m = fr.FRCNN_Feature(bb, nc, min_size=mn, max_size=mx)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in wts.items()}, strict=True)
    m.eval()
    feats = []
    m.roi_heads.box_predictor.register_forward_hook(lambda mod, inp, out: feats.append(inp[0].flatten(1)))
    with torch.no_grad():
        for i in range(4):
            img = synth.synth_image(100000 + i, h, w)
            m([torch.from_numpy(img).permute(2, 0, 1).float().div(255)])
    f = torch.cat(feats)
    mu = f.mean(0).numpy().astype(np.float64)
    wc = wts["roi_heads.box_predictor.cls_score.weight"].astype(np.float64)
    center = -(wc @ mu)
    resid = (f.numpy() - mu) @ wc.T
    print("logit residual std per class:", resid.std(0).round(2))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cald_b200", "planted_calib.json")
    tab = json.load(open(path)) if os.path.exists(path) else {}
    tab[synth.calib_key(depth, nc, 0)] = [float(np.float32(v)) for v in center]
    json.dump(tab, open(path, "w"), indent=0, sort_keys=True)
    print("wrote", path, synth.calib_key(depth, nc, 0))


if __name__ == "__main__":
    main()
