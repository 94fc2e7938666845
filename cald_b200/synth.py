"""Deterministic synthetic inputs: planted detector weights and an image pool.

There is no network for checkpoints or datasets, so the benchmark and the
parity tests use (i) seeded weights of the exact architecture the reference
builds (detection/frcnn_la.py:278-289, retinanet_cal.py:584-625) and (ii) a pool
of low-frequency RGB u8 images of the configured shape.

"Planted" = random directions, but with magnitudes chosen so that the detector
behaves like a trained one numerically: O(1) activations through the residual
stack, a background-dominated classifier with a few confident foreground
detections per image.  A purely random init saturates every score and makes the
top-k / NMS / arg-max stages of the pipeline flip under 1e-7 perturbations
(SURVEY.md section 7), which would make parity untestable.
"""
import numpy as np

from . import arch


def _conv_w(rs, shape, gain=1.0):
    cout, cin, kh, kw = shape
    std = gain * np.sqrt(2.0 / (cin * kh * kw))
    return (rs.standard_normal(shape) * std).astype(np.float32)


def planted_frcnn_weights(depth=50, num_classes=21, seed=0, cls_gain=2.5, bg_bias=9.0,
                          box_gain=0.15, obj_gain=1.5, rpn_box_gain=0.15, fpn_inner_gain=0.3,
                          fpn_layer_gain=0.5, calib=None):
    """torchvision-keyed {name: np.float32 array} for FRCNN_Feature.

    ``calib``: None -> per-class logit centring looked up in planted_calib.json
    (written by tools/calibrate_planted.py); False -> no centring; array -> use it.
    """
    rs = np.random.RandomState(seed)
    shapes = arch.frcnn_params(depth, num_classes)
    w = {}
    for name, shp in shapes.items():
        leaf = name.rsplit(".", 1)[1]
        if len(shp) == 4:
            w[name] = _conv_w(rs, shp)
        elif leaf == "weight" and len(shp) == 2:
            std = np.sqrt(2.0 / shp[1])
            w[name] = (rs.standard_normal(shp) * std).astype(np.float32)
        elif leaf == "running_var":
            w[name] = rs.uniform(0.5, 1.5, shp).astype(np.float32)
        elif leaf == "running_mean":
            w[name] = (rs.standard_normal(shp) * 0.1).astype(np.float32)
        elif leaf == "bias":
            w[name] = (rs.standard_normal(shp) * 0.05).astype(np.float32)
        else:  # FrozenBN weight
            w[name] = rs.uniform(0.8, 1.2, shp).astype(np.float32)
    # keep the residual stack O(1): damp the last BN of every bottleneck
    # (R50 keeps the literal 0.35 its fixtures were generated with; deeper bodies get the gain that gives the same
    #  total variance growth (1 + g^2)^n_blocks over the residual stack)
    nblk = sum(arch.RESNET_BLOCKS[depth])
    g3 = 0.35 if depth == 50 else float(np.sqrt((1.0 + 0.35 ** 2) ** (16.0 / nblk) - 1.0))
    for name in shapes:
        if name.endswith(".bn3.weight"):
            w[name] *= np.float32(g3)
        if name.endswith(".downsample.1.weight"):
            w[name] *= np.float32(0.8)
    for name in shapes:
        if ".fpn.inner_blocks." in name and name.endswith("weight"):
            w[name] *= np.float32(fpn_inner_gain)
        if ".fpn.layer_blocks." in name and name.endswith("weight"):
            w[name] *= np.float32(fpn_layer_gain)
    # heads
    w["rpn.head.cls_logits.weight"] *= np.float32(obj_gain)
    w["rpn.head.bbox_pred.weight"] *= np.float32(rpn_box_gain)
    # remove the component along the (all-positive, post-ReLU) mean feature so the
    # winning class varies from proposal to proposal instead of being one global class
    for name in ("roi_heads.box_predictor.cls_score.weight", "rpn.head.cls_logits.weight"):
        ww = w[name].reshape(w[name].shape[0], -1)
        ww -= ww.mean(axis=1, keepdims=True)
    w["roi_heads.box_predictor.cls_score.weight"] *= np.float32(cls_gain)
    b = w["roi_heads.box_predictor.cls_score.bias"]
    b[0] += np.float32(bg_bias)
    w["roi_heads.box_predictor.bbox_pred.weight"] *= np.float32(box_gain)
    if calib is None:
        tab = _calib_table()
        key = calib_key(depth, num_classes, seed)
        if key not in tab:
            raise KeyError("no planted calibration for %s; run tools/calibrate_planted.py" % key)
        calib = tab[key]
    if calib is not False:
        # calibrated at cls_gain=1; the centring scales with the classifier gain
        b += np.asarray(calib, dtype=np.float32) * np.float32(cls_gain / _CALIB_GAIN)
    return w


def planted_retinanet_weights(num_classes=21, seed=0, depth=50, cls_gain=6.0, reg_gain=0.12, fpn_inner_gain=0.3,
                             fpn_layer_gain=0.5, tower_gain=1.0, tower_sparsity=3.0, cls_bias_shift=-7.0):
    """torchvision-keyed {name: np.float32 array} for retinanet_resnet50_fpn_cal (retinanet_cal.py:584-625).

    Same body recipe as ``planted_frcnn_weights``.  The classification logits keep the reference's prior bias
    -log(99) (retinanet_cal.py:90) lowered by ``cls_bias_shift`` and a zero-mean weight of large gain on top of a
    SPARSE last tower layer (its bias is pushed down by ``tower_sparsity``), which makes the logit distribution
    heavy-tailed like a trained detector's: nearly all of the 243k x K (anchor, class) scores sit far below the
    0.05 threshold and the few that cross it are spread over several logit units -- tens to hundreds of
    candidates per image instead of none (random init) or all of them (bumped bias); see SURVEY.md section 7.
    """
    rs = np.random.RandomState(seed + 7919)
    shapes = arch.retinanet_params(depth, num_classes)
    w = {}
    for name, shp in shapes.items():
        leaf = name.rsplit(".", 1)[1]
        if len(shp) == 4:
            w[name] = _conv_w(rs, shp)
        elif leaf == "running_var":
            w[name] = rs.uniform(0.5, 1.5, shp).astype(np.float32)
        elif leaf == "running_mean":
            w[name] = (rs.standard_normal(shp) * 0.1).astype(np.float32)
        elif leaf == "bias":
            w[name] = (rs.standard_normal(shp) * 0.05).astype(np.float32)
        else:  # FrozenBN weight
            w[name] = rs.uniform(0.8, 1.2, shp).astype(np.float32)
    nblk = sum(arch.RESNET_BLOCKS[depth])
    g3 = 0.35 if depth == 50 else float(np.sqrt((1.0 + 0.35 ** 2) ** (16.0 / nblk) - 1.0))
    for name in shapes:
        if name.endswith(".bn3.weight"):
            w[name] *= np.float32(g3)
        if name.endswith(".downsample.1.weight"):
            w[name] *= np.float32(0.8)
        if ".fpn.inner_blocks." in name and name.endswith("weight"):
            w[name] *= np.float32(fpn_inner_gain)
        if (".fpn.layer_blocks." in name or ".fpn.extra_blocks." in name) and name.endswith("weight"):
            w[name] *= np.float32(fpn_layer_gain)
        if ".conv." in name and name.startswith("head.") and name.endswith("weight"):
            w[name] *= np.float32(tower_gain)
    cw = w["head.classification_head.cls_logits.weight"]
    cw -= cw.reshape(cw.shape[0], -1).mean(axis=1).reshape(-1, 1, 1, 1)
    cw *= np.float32(cls_gain)
    w["head.classification_head.cls_logits.bias"] = np.full(cw.shape[0], -np.log(99.0) + cls_bias_shift,
                                                            dtype=np.float32) + \
        (rs.standard_normal(cw.shape[0]) * 0.05).astype(np.float32)
    w["head.classification_head.conv.6.bias"] -= np.float32(tower_sparsity)
    w["head.regression_head.bbox_reg.weight"] *= np.float32(reg_gain)
    return w


_CALIB_GAIN = 2.5


def calib_key(depth, num_classes, seed):
    return "frcnn_r%d_nc%d_seed%d" % (depth, num_classes, seed)


def _calib_table():
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "planted_calib.json")
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        return json.load(f)


def synth_image(index, height, width, seed=0, n_rect=6):
    """u8 HxWx3: bilinear-upsampled coarse noise plus a few textured rectangles."""
    rs = np.random.RandomState((seed * 1000003 + index) % (2 ** 31 - 1))
    gh, gw = 13, 17
    coarse = rs.uniform(30, 225, (gh, gw, 3))
    ys = np.linspace(0, gh - 1, height)
    xs = np.linspace(0, gw - 1, width)
    y0 = np.minimum(np.floor(ys).astype(int), gh - 2)
    x0 = np.minimum(np.floor(xs).astype(int), gw - 2)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    a = coarse[y0][:, x0]
    b = coarse[y0][:, x0 + 1]
    c = coarse[y0 + 1][:, x0]
    d = coarse[y0 + 1][:, x0 + 1]
    img = (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx)
    for _ in range(n_rect):
        rh = int(rs.uniform(0.08, 0.45) * height)
        rw = int(rs.uniform(0.08, 0.45) * width)
        ry = int(rs.uniform(0, height - rh))
        rx = int(rs.uniform(0, width - rw))
        col = rs.uniform(0, 255, 3)
        tex = rs.uniform(-25, 25, (rh // 8 + 1, rw // 8 + 1, 3))
        tex = np.repeat(np.repeat(tex, 8, 0), 8, 1)[:rh, :rw]
        img[ry:ry + rh, rx:rx + rw] = col + tex
    # fancy indexing above leaves `img` in a permuted memory order; hand out a C-contiguous HWC image
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def synth_pool(n, height, width, seed=0, portrait_every=0):
    """List of n u8 images; every ``portrait_every``-th is transposed-shape (W x H)."""
    out = []
    for i in range(n):
        if portrait_every and i % portrait_every == portrait_every - 1:
            out.append(synth_image(i, width, height, seed))
        else:
            out.append(synth_image(i, height, width, seed))
    return out
