"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of one RetinaNet forward.

Restates what ``task_model([image])`` computes for the reference's RetinaNet
(detection/retinanet_cal.py:36-62 heads, 135-151 / 225-241 head forwards, 347-351
anchors, 402-490 ``postprocess_detections``, 492-575 ``forward``, 584-625 factory)
on top of the un-vendored torchvision 0.26.0 pieces it imports (``tv:`` =
site-packages/torchvision: ops/feature_pyramid_network.py:224-250 LastLevelP6P7,
models/detection/anchor_utils.py:58-133, models/detection/transform.py:257-319).
It is the checker for the CUDA engine: nothing under ``cald_b200/`` may import it.

Pinned against the real reference by ``tests/golden/make_golden_retina.py`` (run
where /root/reference exists); fixtures in tests/golden/retina_*.npz.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import frcnn_oracle as fo

NUM_ANCHORS = 9


class Cfg:
    def __init__(self, depth=50, num_classes=21, min_size=600, max_size=1000, score_thresh=0.05,
                 nms_thresh=0.5, detections_per_img=300):
        self.__dict__.update(locals())
        del self.__dict__["self"]


def anchor_sizes():
    """retinanet_cal.py:347-348: x, int(x * 2^(1/3)), int(x * 2^(2/3)) for x in 32..512."""
    return tuple((x, int(x * 2 ** (1.0 / 3)), int(x * 2 ** (2.0 / 3))) for x in [32, 64, 128, 256, 512])


def cell_anchors(scales, ratios=(0.5, 1.0, 2.0)):
    """tv:anchor_utils.py:58-75: ratio-major, scale-minor; fp32; round-half-even."""
    sc = torch.as_tensor(scales, dtype=torch.float32)
    ar = torch.as_tensor(ratios, dtype=torch.float32)
    hr = torch.sqrt(ar)
    wr = 1 / hr
    ws = (wr[:, None] * sc[None, :]).view(-1)
    hs = (hr[:, None] * sc[None, :]).view(-1)
    return (torch.stack([-ws, -hs, ws, hs], dim=1) / 2).round()


def grid_anchors(padded_hw, feat_hw_list):
    """tv:anchor_utils.py:84-133 with the RetinaNet sizes: per level (H*W*9, 4), order (y, x, a)."""
    out = []
    for (gh, gw), scales in zip(feat_hw_list, anchor_sizes()):
        sh, sw = padded_hw[0] // gh, padded_hw[1] // gw
        base = cell_anchors(scales)
        sx = torch.arange(0, gw, dtype=torch.int32) * sw
        sy = torch.arange(0, gh, dtype=torch.int32) * sh
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        xx, yy = xx.reshape(-1), yy.reshape(-1)
        shifts = torch.stack((xx, yy, xx, yy), dim=1)
        out.append((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4))
    return out


def fpn_p3_p7(cs, w):
    """resnet_fpn_backbone(returned_layers=[2,3,4], extra_blocks=LastLevelP6P7(256,256)), retinanet_cal.py:618-619.

    cs = [C2, C3, C4, C5]; inner/layer blocks 0..2 sit on C3..C5; P6 = conv3x3/2(P5) (use_P5: 256 == 256),
    P7 = conv3x3/2(relu(P6))  (tv:ops/feature_pyramid_network.py:238-250).
    """
    feats = cs[1:]

    def inner(i, x):
        return F.conv2d(x, fo._t(w, "backbone.fpn.inner_blocks.%d.0.weight" % i),
                        fo._t(w, "backbone.fpn.inner_blocks.%d.0.bias" % i))

    def layer(i, x):
        return F.conv2d(x, fo._t(w, "backbone.fpn.layer_blocks.%d.0.weight" % i),
                        fo._t(w, "backbone.fpn.layer_blocks.%d.0.bias" % i), padding=1)

    last = inner(2, feats[2])
    res = [layer(2, last)]
    for i in (1, 0):
        lat = inner(i, feats[i])
        last = lat + F.interpolate(last, size=lat.shape[-2:], mode="nearest")
        res.insert(0, layer(i, last))
    p6 = F.conv2d(res[-1], fo._t(w, "backbone.fpn.extra_blocks.p6.weight"),
                  fo._t(w, "backbone.fpn.extra_blocks.p6.bias"), stride=2, padding=1)
    p7 = F.conv2d(F.relu(p6), fo._t(w, "backbone.fpn.extra_blocks.p7.weight"),
                  fo._t(w, "backbone.fpn.extra_blocks.p7.bias"), stride=2, padding=1)
    return res + [p6, p7]


def heads(feats, w, num_classes):
    """retinanet_cal.py:135-151 and 225-241: -> (cls_logits (sum HWA, K), bbox_regression (sum HWA, 4))."""
    cls_all, reg_all = [], []
    for f in feats:
        for head, last, outs, width in (("classification_head", "cls_logits", cls_all, num_classes),
                                        ("regression_head", "bbox_reg", reg_all, 4)):
            t = f
            for i in (0, 2, 4, 6):
                t = F.relu(F.conv2d(t, fo._t(w, "head.%s.conv.%d.weight" % (head, i)),
                                    fo._t(w, "head.%s.conv.%d.bias" % (head, i)), padding=1))
            o = F.conv2d(t, fo._t(w, "head.%s.%s.weight" % (head, last)), fo._t(w, "head.%s.%s.bias" % (head, last)),
                         padding=1)
            n, _, hh, ww = o.shape
            o = o.view(n, -1, width, hh, ww).permute(0, 3, 4, 1, 2).reshape(n, -1, width)
            outs.append(o[0])
    return torch.cat(cls_all), torch.cat(reg_all)


def postprocess_detections(cls_logits, bbox_reg, anchors, image_hw, cfg):
    """retinanet_cal.py:402-490 for ONE image (class loop; results concatenated in class order)."""
    scores = torch.sigmoid(cls_logits)
    boxes = fo.decode_boxes(bbox_reg, anchors, (1.0, 1.0, 1.0, 1.0))
    boxes = fo.clip_boxes(boxes, image_hw)
    k = scores.shape[1]
    ob, osc, olab, opm, ocls, oanc = [], [], [], [], [], []
    for c in range(k):
        inds = torch.where(scores[:, c] > cfg.score_thresh)[0]
        if inds.numel() == 0:
            continue
        b = boxes[inds]
        s = scores[inds, c]
        ws, hs = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
        keep = torch.where((ws >= 1e-2) & (hs >= 1e-2))[0]     # box_ops.remove_small_boxes(min_size=1e-2)
        inds, b, s = inds[keep], b[keep], s[keep]
        kk = torch.from_numpy(fo.nms_numpy(b.numpy(), s.numpy(), cfg.nms_thresh))[:cfg.detections_per_img]
        inds = inds[kk]
        ob.append(b[kk]); osc.append(s[kk]); oanc.append(inds)
        olab.append(torch.full((len(kk),), c, dtype=torch.int64))
        ocls.append(scores[inds])
        opm.append(scores[inds].max(dim=1)[0] if len(kk) else scores.new_zeros((0,)))
    if not ob:
        z = scores.new_zeros
        return {"boxes": z((0, 4)), "scores": z((0,)), "labels": torch.zeros((0,), dtype=torch.int64),
                "scores_cls": z((0, k)), "prob_max": z((0,)), "anchor_idx": torch.zeros((0,), dtype=torch.int64)}
    return {"boxes": torch.cat(ob), "scores": torch.cat(osc), "labels": torch.cat(olab),
            "scores_cls": torch.cat(ocls), "prob_max": torch.cat(opm), "anchor_idx": torch.cat(oanc)}


def resize_boxes_tv(boxes, from_hw, to_hw):
    """tv:models/detection/transform.py:306-319: the ratios are fp32 tensors (fp32 division)."""
    rh = torch.tensor(to_hw[0], dtype=torch.float32) / torch.tensor(from_hw[0], dtype=torch.float32)
    rw = torch.tensor(to_hw[1], dtype=torch.float32) / torch.tensor(from_hw[1], dtype=torch.float32)
    x1, y1, x2, y2 = boxes.unbind(1)
    return torch.stack((x1 * rw, y1 * rh, x2 * rw, y2 * rh), dim=1)


def forward(img_chw, w, cfg, stages=None):
    """One ``task_model([img])[0]`` of retinanet_resnet50_fpn_cal (plus 'anchor_idx' for the parity tests)."""
    with torch.no_grad():
        oh, ow = img_chw.shape[-2:]
        x, image_hw = fo.transform(img_chw, cfg)
        cs = fo.resnet_body(x, w, cfg.depth)
        feats = fpn_p3_p7(cs, w)
        logits, reg = heads(feats, w, cfg.num_classes)
        anchors = torch.cat(grid_anchors(tuple(x.shape[-2:]), [tuple(f.shape[-2:]) for f in feats]))
        det = postprocess_detections(logits, reg, anchors, image_hw, cfg)
        det["boxes"] = resize_boxes_tv(det["boxes"], image_hw, (oh, ow))
        if stages is not None:
            stages.update(dict(input=x, image_hw=image_hw, c=cs, p=feats, cls_logits=logits, bbox_regression=reg,
                               anchors=anchors))
        return det
