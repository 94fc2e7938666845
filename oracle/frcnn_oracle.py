"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of one Faster R-CNN forward.

Restates, stage by stage, what ``task_model([image])`` computes on the CALD
scoring path (reference: detection/frcnn_la.py:237-275 and the un-vendored
torchvision 0.26.0 it subclasses; ``tv:`` = site-packages/torchvision).  It is
the checker for the CUDA engine: nothing under ``cald_b200/`` may import it.

Pinned against the real reference by ``tests/golden/make_golden.py`` (run where
/root/reference exists): every stage tensor of this restatement is compared with
forward hooks on the unmodified ``FRCNN_Feature`` and the resulting detections
are committed as fixtures (tests/golden/*.npz, tests/test_oracle_golden.py).

Dense arithmetic uses torch CPU fp32 functional ops (conv2d / linear /
interpolate / max_pool2d), i.e. the same ATen kernels the reference's CPU run
uses; everything discrete (anchors, top-k, NMS, level mapping, RoIAlign
sampling, post-processing) is restated in numpy / explicit torch indexing.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from cald_b200 import arch

IMAGE_MEAN = (0.485, 0.456, 0.406)  # frcnn_la.py:230-233
IMAGE_STD = (0.229, 0.224, 0.225)
BBOX_XFORM_CLIP = math.log(1000.0 / 16)  # tv:models/detection/_utils.py:148


class Cfg:
    def __init__(self, depth=50, num_classes=21, min_size=600, max_size=1000,
                 rpn_pre_nms_top_n=1000, rpn_post_nms_top_n=1000, rpn_nms_thresh=0.7,
                 rpn_min_size=1e-3, box_score_thresh=0.05, box_nms_thresh=0.5,
                 box_detections_per_img=100):
        self.__dict__.update(locals())
        del self.__dict__["self"]


def _t(w, name):
    v = w[name]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))


# ---------------------------------------------------------------- transform
def resize_scale(h, w, min_size, max_size):
    """tv:models/detection/transform.py:57-62 (eager branch): python doubles."""
    return min(float(min_size) / float(min(h, w)), float(max_size) / float(max(h, w)))


def resized_hw(h, w, min_size, max_size):
    """F.interpolate(recompute_scale_factor=True): floor(size * scale) in double."""
    s = resize_scale(h, w, min_size, max_size)
    return int(math.floor(float(h) * s)), int(math.floor(float(w) * s))


def transform(img_chw, cfg):
    """normalize -> bilinear resize -> zero-pad to /32.  tv:transform.py:119-255.

    img_chw: float32 tensor 3xHxW in [0,1] (any float really; noise augs exceed it).
    Returns (padded 1x3xHpxWp, (h_resized, w_resized)).
    """
    mean = torch.tensor(IMAGE_MEAN, dtype=torch.float32)[:, None, None]
    std = torch.tensor(IMAGE_STD, dtype=torch.float32)[:, None, None]
    x = (img_chw - mean) / std
    h, w = x.shape[-2:]
    s = resize_scale(h, w, cfg.min_size, cfg.max_size)
    x = F.interpolate(x[None], size=None, scale_factor=s, mode="bilinear",
                      recompute_scale_factor=True, align_corners=False)[0]
    rh, rw = x.shape[-2:]
    ph = int(math.ceil(float(rh) / 32.0) * 32)
    pw = int(math.ceil(float(rw) / 32.0) * 32)
    out = x.new_zeros((1, 3, ph, pw))
    out[0, :, :rh, :rw] = x
    return out, (rh, rw)


# ---------------------------------------------------------------- backbone
def frozen_bn(x, w, prefix):
    """tv:ops/misc.py:54-63, eps = 1e-5."""
    scale = _t(w, prefix + ".weight") * (_t(w, prefix + ".running_var") + 1e-5).rsqrt()
    bias = _t(w, prefix + ".bias") - _t(w, prefix + ".running_mean") * scale
    return x * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)


def resnet_body(x, w, depth):
    """tv:models/resnet.py Bottleneck (stride on conv2), returns [C2, C3, C4, C5]."""
    x = F.conv2d(x, _t(w, "backbone.body.conv1.weight"), None, stride=2, padding=3)
    x = F.relu(frozen_bn(x, w, "backbone.body.bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = []
    for li, nblk in enumerate(arch.RESNET_BLOCKS[depth]):
        for b in range(nblk):
            pre = "backbone.body.layer%d.%d" % (li + 1, b)
            stride = 2 if (b == 0 and li > 0) else 1
            idt = x
            o = F.relu(frozen_bn(F.conv2d(x, _t(w, pre + ".conv1.weight")), w, pre + ".bn1"))
            o = F.conv2d(o, _t(w, pre + ".conv2.weight"), None, stride=stride, padding=1)
            o = F.relu(frozen_bn(o, w, pre + ".bn2"))
            o = frozen_bn(F.conv2d(o, _t(w, pre + ".conv3.weight")), w, pre + ".bn3")
            if b == 0:
                idt = frozen_bn(F.conv2d(x, _t(w, pre + ".downsample.0.weight"), None, stride=stride),
                                w, pre + ".downsample.1")
            x = F.relu(o + idt)
        outs.append(x)
    return outs


def fpn(cs, w):
    """tv:ops/feature_pyramid_network.py:172-221 with LastLevelMaxPool -> [P2..P5, pool]."""
    n = len(cs)

    def inner(i, x):
        return F.conv2d(x, _t(w, "backbone.fpn.inner_blocks.%d.0.weight" % i),
                        _t(w, "backbone.fpn.inner_blocks.%d.0.bias" % i))

    def layer(i, x):
        return F.conv2d(x, _t(w, "backbone.fpn.layer_blocks.%d.0.weight" % i),
                        _t(w, "backbone.fpn.layer_blocks.%d.0.bias" % i), padding=1)

    last = inner(n - 1, cs[-1])
    res = [layer(n - 1, last)]
    for i in range(n - 2, -1, -1):
        lat = inner(i, cs[i])
        td = F.interpolate(last, size=lat.shape[-2:], mode="nearest")
        last = lat + td
        res.insert(0, layer(i, last))
    res.append(F.max_pool2d(res[-1], kernel_size=1, stride=2, padding=0))
    return res


# ---------------------------------------------------------------- RPN
def rpn_head(feats, w):
    """tv:models/detection/rpn.py:71-78.  Returns per level (logits HxWx3, deltas HxWx3x4)."""
    out = []
    for f in feats:
        t = F.relu(F.conv2d(f, _t(w, "rpn.head.conv.0.0.weight"), _t(w, "rpn.head.conv.0.0.bias"), padding=1))
        lg = F.conv2d(t, _t(w, "rpn.head.cls_logits.weight"), _t(w, "rpn.head.cls_logits.bias"))
        dl = F.conv2d(t, _t(w, "rpn.head.bbox_pred.weight"), _t(w, "rpn.head.bbox_pred.bias"))
        a = lg.shape[1]
        hh, ww = lg.shape[-2:]
        # permute_and_flatten (rpn.py:88-110): (N, A*C, H, W) -> (N, H, W, A, C)
        lg = lg[0].permute(1, 2, 0).reshape(-1)
        dl = dl[0].reshape(a, 4, hh, ww).permute(2, 3, 0, 1).reshape(-1, 4)
        out.append((lg, dl))
    return out


ANCHOR_SIZES = (32, 64, 128, 256, 512)
ANCHOR_RATIOS = (0.5, 1.0, 2.0)


def cell_anchors(size, ratios=ANCHOR_RATIOS):
    """tv:anchor_utils.py:58-75 (fp32 sqrt, round-half-even)."""
    scales = torch.tensor([size], dtype=torch.float32)
    ar = torch.tensor(ratios, dtype=torch.float32)
    hr = torch.sqrt(ar)
    wr = 1 / hr
    ws = (wr[:, None] * scales[None, :]).view(-1)
    hs = (hr[:, None] * scales[None, :]).view(-1)
    return (torch.stack([-ws, -hs, ws, hs], dim=1) / 2).round()


def grid_anchors(padded_hw, feat_hw_list):
    """tv:anchor_utils.py:84-133: stride = padded // grid, order (y, x, a)."""
    out = []
    for (gh, gw), size in zip(feat_hw_list, ANCHOR_SIZES):
        sh, sw = padded_hw[0] // gh, padded_hw[1] // gw
        base = cell_anchors(size)
        sx = torch.arange(0, gw, dtype=torch.int32) * sw
        sy = torch.arange(0, gh, dtype=torch.int32) * sh
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        xx, yy = xx.reshape(-1), yy.reshape(-1)
        shifts = torch.stack((xx, yy, xx, yy), dim=1)
        out.append((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4))
    return out


def decode_boxes(deltas, boxes, weights):
    """tv:_utils.py:183-224.  deltas (N, 4k), boxes (N, 4) -> (N, 4k)."""
    wx, wy, ww, wh = weights
    widths = boxes[:, 2] - boxes[:, 0]
    heights = boxes[:, 3] - boxes[:, 1]
    cx = boxes[:, 0] + 0.5 * widths
    cy = boxes[:, 1] + 0.5 * heights
    dx = deltas[:, 0::4] / wx
    dy = deltas[:, 1::4] / wy
    dw = torch.clamp(deltas[:, 2::4] / ww, max=BBOX_XFORM_CLIP)
    dh = torch.clamp(deltas[:, 3::4] / wh, max=BBOX_XFORM_CLIP)
    pcx = dx * widths[:, None] + cx[:, None]
    pcy = dy * heights[:, None] + cy[:, None]
    pw = torch.exp(dw) * widths[:, None]
    ph = torch.exp(dh) * heights[:, None]
    hw_, hh_ = 0.5 * pw, 0.5 * ph
    return torch.stack((pcx - hw_, pcy - hh_, pcx + hw_, pcy + hh_), dim=2).flatten(1)


def clip_boxes(boxes, hw):
    """tv:ops/boxes.py clip_boxes_to_image: x in [0, W], y in [0, H]."""
    h, w = hw
    b = boxes.clone()
    b[..., 0::2] = b[..., 0::2].clamp(min=0, max=w)
    b[..., 1::2] = b[..., 1::2].clamp(min=0, max=h)
    return b


# Timing mode (bench.py's CPU arm only): run NMS and RoIAlign through torchvision's own CPU kernels -- the very
# kernels the reference calls (tv:ops/boxes.py nms, tv:ops/roi_align.py) -- instead of the numpy restatements below,
# which are 1.3-1.8x slower per image than the reference's loop (profiles/r02_cpu_port_vs_reference.md) and would
# flatter the GPU/CPU ratio.  Results are bit-identical either way (tests/test_oracle_kat.py); the parity tests keep
# the independent numpy restatements.
USE_TORCHVISION_OPS = False


def nms_numpy(boxes, scores, thresh):
    """Greedy NMS, same arithmetic as torchvision's CPU kernel (fp32; > thresh suppresses).

    Candidates are visited by descending score with ties broken by ascending index
    (stable); returns kept indices in that order.
    """
    if USE_TORCHVISION_OPS:
        import torchvision
        return torchvision.ops.nms(torch.as_tensor(np.asarray(boxes, dtype=np.float32)),
                                   torch.as_tensor(np.asarray(scores, dtype=np.float32)), float(thresh)).numpy()
    boxes = np.asarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-scores, kind="stable")
    x1, y1, x2, y2 = (boxes[:, i] for i in range(4))
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thresh = np.float32(thresh)
    for oi in range(n):
        i = order[oi]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[oi + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        ww = np.maximum(np.float32(0), xx2 - xx1)
        hh = np.maximum(np.float32(0), yy2 - yy1)
        inter = ww * hh
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thresh]] = True
    return np.asarray(keep, dtype=np.int64)


def rpn_proposals(head_out, padded_hw, feat_hw_list, image_hw, cfg):
    """tv:rpn.py:242-297, 336-372 for ONE image -> (proposals (P,4), scores (P,))."""
    anchors = grid_anchors(padded_hw, feat_hw_list)
    lv_boxes, lv_scores = [], []
    for (lg, dl), anc in zip(head_out, anchors):
        k = min(cfg.rpn_pre_nms_top_n, lg.shape[0])
        # topk by logit; explicit tie rule = lower index first
        order = np.argsort(-lg.numpy(), kind="stable")[:k]
        idx = torch.from_numpy(order)
        props = decode_boxes(dl[idx], anc[idx], (1.0, 1.0, 1.0, 1.0))
        sc = torch.sigmoid(lg[idx])
        props = clip_boxes(props, image_hw)
        ws, hs = props[:, 2] - props[:, 0], props[:, 3] - props[:, 1]
        keep = (ws >= cfg.rpn_min_size) & (hs >= cfg.rpn_min_size) & (sc >= 0.0)
        props, sc = props[keep], sc[keep]
        kept = nms_numpy(props.numpy(), sc.numpy(), cfg.rpn_nms_thresh)
        lv_boxes.append(props[kept])
        lv_scores.append(sc[kept])
    boxes = torch.cat(lv_boxes)
    scores = torch.cat(lv_scores)
    order = np.argsort(-scores.numpy(), kind="stable")[:cfg.rpn_post_nms_top_n]
    return boxes[order], scores[order]


# ---------------------------------------------------------------- RoI heads
def map_levels(boxes, k_min=2, k_max=5):
    """tv:ops/poolers.py:73-84."""
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    s = torch.sqrt(area)
    lv = torch.floor(4 + torch.log2(s / 224) + torch.tensor(1e-6, dtype=s.dtype))
    return (torch.clamp(lv, min=k_min, max=k_max).to(torch.int64) - k_min)


def roi_align_level(feat, rois, scale, out=7, sr=2):
    """RoIAlign aligned=False, following tv:ops/roi_align.py:115-200 / the C++ CPU kernel.

    feat: C x H x W; rois: K x 4 (image coords).  Returns K x C x out x out.
    """
    c, h, w = feat.shape
    k = rois.shape[0]
    if k == 0:
        return feat.new_zeros((0, c, out, out))
    if USE_TORCHVISION_OPS:
        import torchvision
        return torchvision.ops.roi_align(feat[None], [rois], (out, out), spatial_scale=scale, sampling_ratio=sr,
                                         aligned=False)
    x1 = rois[:, 0] * scale
    y1 = rois[:, 1] * scale
    x2 = rois[:, 2] * scale
    y2 = rois[:, 3] * scale
    rw = torch.clamp(x2 - x1, min=1.0)
    rh = torch.clamp(y2 - y1, min=1.0)
    bw = rw / out
    bh = rh / out
    p = torch.arange(out, dtype=torch.float32)
    i = torch.arange(sr, dtype=torch.float32)
    # y[k, ph, iy] = y1 + ph*bh + (iy + .5) * bh / sr
    ys = y1[:, None, None] + p[None, :, None] * bh[:, None, None] + (i[None, None, :] + 0.5) * (bh / sr)[:, None, None]
    xs = x1[:, None, None] + p[None, :, None] * bw[:, None, None] + (i[None, None, :] + 0.5) * (bw / sr)[:, None, None]

    def prep(v, size):
        bad = (v < -1.0) | (v > size)
        v = v.clamp(min=0)
        lo = v.to(torch.int64)
        top = lo >= size - 1
        lo = torch.where(top, torch.full_like(lo, size - 1), lo)
        hi = torch.where(top, lo, lo + 1)
        v = torch.where(top, lo.to(v.dtype), v)
        l = v - lo.to(v.dtype)
        return lo, hi, l, 1.0 - l, bad

    ylo, yhi, ly, hy, ybad = prep(ys, h)
    xlo, xhi, lx, hx, xbad = prep(xs, w)
    # gather: feat[:, y, x] for all (k, ph, iy, pw, ix)
    def g(yi, xi):
        return feat[:, yi[:, :, :, None, None], xi[:, None, None, :, :]]  # C,K,PH,IY,PW,IX
    w1 = (hy[:, :, :, None, None] * hx[:, None, None, :, :])
    w2 = (hy[:, :, :, None, None] * lx[:, None, None, :, :])
    w3 = (ly[:, :, :, None, None] * hx[:, None, None, :, :])
    w4 = (ly[:, :, :, None, None] * lx[:, None, None, :, :])
    val = w1 * g(ylo, xlo) + w2 * g(ylo, xhi) + w3 * g(yhi, xlo) + w4 * g(yhi, xhi)
    bad = ybad[:, :, :, None, None] | xbad[:, None, None, :, :]
    val = torch.where(bad[None], torch.zeros((), dtype=val.dtype), val)
    # sum order of the CPU kernel: iy outer, ix inner
    acc = None
    for a in range(sr):
        for b in range(sr):
            t = val[:, :, :, a, :, b]
            acc = t if acc is None else acc + t
    acc = acc / float(sr * sr)
    return acc.permute(1, 0, 2, 3).contiguous()


def multiscale_roi_align(feats, rois):
    """tv:ops/poolers.py:200-222,289-321 on P2..P5 (the 'pool' level is not used)."""
    lv = map_levels(rois)
    res = feats[0].new_zeros((rois.shape[0], feats[0].shape[1], 7, 7))
    for l in range(4):
        idx = torch.where(lv == l)[0]
        if idx.numel():
            res[idx] = roi_align_level(feats[l][0], rois[idx], 1.0 / (4 * 2 ** l))
    return res


def box_head(x, w):
    """TwoMLPHead + FastRCNNPredictor, tv:faster_rcnn.py:286-307,347-372."""
    x = x.flatten(1)
    x = F.relu(F.linear(x, _t(w, "roi_heads.box_head.fc6.weight"), _t(w, "roi_heads.box_head.fc6.bias")))
    x = F.relu(F.linear(x, _t(w, "roi_heads.box_head.fc7.weight"), _t(w, "roi_heads.box_head.fc7.bias")))
    lg = F.linear(x, _t(w, "roi_heads.box_predictor.cls_score.weight"), _t(w, "roi_heads.box_predictor.cls_score.bias"))
    dl = F.linear(x, _t(w, "roi_heads.box_predictor.bbox_pred.weight"), _t(w, "roi_heads.box_predictor.bbox_pred.bias"))
    return lg, dl


def postprocess_detections(logits, deltas, proposals, image_hw, cfg):
    """detection/frcnn_la.py:32-87 for ONE image."""
    nc = logits.shape[1]
    boxes = decode_boxes(deltas, proposals, (10.0, 10.0, 5.0, 5.0)).reshape(-1, nc, 4)
    scores = F.softmax(logits, -1)
    boxes = clip_boxes(boxes, image_hw)
    n = scores.shape[0]
    prob_max = scores[:, 1:].max(1)[0]
    cand_p, cand_c = torch.where(scores[:, 1:] > cfg.box_score_thresh)  # row-major == reference flatten order
    cand_c = cand_c + 1
    cb = boxes[cand_p, cand_c]
    cs = scores[cand_p, cand_c]
    keep_all = []
    for c in torch.unique(cand_c).tolist():
        ii = torch.where(cand_c == c)[0]
        kk = nms_numpy(cb[ii].numpy(), cs[ii].numpy(), cfg.box_nms_thresh)
        keep_all.append(ii[torch.from_numpy(kk)])
    if keep_all:
        keep = torch.sort(torch.cat(keep_all))[0]
        order = np.argsort(-cs[keep].numpy(), kind="stable")
        keep = keep[torch.from_numpy(order)][:cfg.box_detections_per_img]
    else:
        keep = torch.zeros((0,), dtype=torch.int64)
    p = cand_p[keep]
    return {
        "boxes": cb[keep], "scores": cs[keep], "labels": cand_c[keep],
        "props": proposals[p], "prob_max": prob_max[p], "scores_cls": scores[p],
    }


def resize_boxes(boxes, from_hw, to_hw):
    """detection/frcnn_la.py:307-315: python-float ratios times fp32 tensor."""
    rh = float(to_hw[0]) / float(from_hw[0])
    rw = float(to_hw[1]) / float(from_hw[1])
    x1, y1, x2, y2 = boxes.unbind(1)
    return torch.stack((x1 * rw, y1 * rh, x2 * rw, y2 * rh), dim=1)


# ---------------------------------------------------------------- whole forward
def forward(img_chw, w, cfg, stages=None):
    """One ``task_model([img])[0]`` (without the 'features' entry).

    img_chw: torch float32 3xHxW.  ``stages`` (optional dict) receives every
    intermediate tensor for the stage-wise parity tests.
    """
    with torch.no_grad():
        oh, ow = img_chw.shape[-2:]
        x, image_hw = transform(img_chw, cfg)
        cs = resnet_body(x, w, cfg.depth)
        feats = fpn(cs, w)
        head = rpn_head(feats, w)
        feat_hw = [tuple(f.shape[-2:]) for f in feats]
        props, pscores = rpn_proposals(head, tuple(x.shape[-2:]), feat_hw, image_hw, cfg)
        pooled = multiscale_roi_align(feats[:4], props)
        logits, deltas = box_head(pooled, w)
        det = postprocess_detections(logits, deltas, props, image_hw, cfg)
        det["boxes"] = resize_boxes(det["boxes"], image_hw, (oh, ow))
        det["props"] = resize_boxes(det["props"], image_hw, (oh, ow))
        if stages is not None:
            stages.update(dict(input=x, image_hw=image_hw, c=cs, p=feats, rpn=head, proposals=props,
                               proposal_scores=pscores, pooled=pooled, logits=logits, deltas=deltas))
        return det
