"""TEST INFRASTRUCTURE ONLY -- restatement of CALD's scoring loop and augmentations.

Follows cald_train.py:91-231 (``get_uncertainty``), cald_train.py:234-271
(``cls_kldiv``), cald_train.py:439-457 (selection) and cald/cald_helper.py:23-243
(augmentations) in numpy / torch-CPU.  The detector is passed in as a callable so
the same loop can sit on top of the oracle forward (oracle/frcnn_oracle.py), the
real reference model, or recorded detections.  Nothing under ``cald_b200/`` may
import this module.

Pinned by tests/golden (generated with the unmodified reference, see
tests/golden/make_golden.py) and by known-answer tests in tests/test_oracle_*.py.
"""
import math
import random

import numpy as np
import torch

from . import pil_oracle

AUG_NAMES = ['flip', 'multi_ga', 'color_adjust', 'color_swap', 'multi_color_adjust', 'multi_sp', 'cut_out',
             'multi_cut_out', 'multi_resize', 'larger_resize', 'smaller_resize', 'rotation', 'ga', 'sp']


def to_tensor(img_u8):
    """torchvision F.to_tensor on an HxWx3 u8 array: CHW float32, value / 255."""
    return torch.from_numpy(np.ascontiguousarray(img_u8)).permute(2, 0, 1).to(torch.float32).div(255)


def subsample_indices(n):
    """cald_train.py:110-111: > 40 detections -> 50 linspace picks (banker's rounding, duplicates)."""
    if n > 40:
        return np.round(np.linspace(0, n - 1, 50)).astype(int)
    return np.arange(n)


def class_max_vector(scores, labels, num_cls):
    """cald_train.py:114-116: cls_corr[l-1] = max score (python negative index wraps for l = 0)."""
    v = [0.0] * (num_cls - 1)
    for s, l in zip(np.asarray(scores).tolist(), np.asarray(labels).tolist()):
        v[l - 1] = max(v[l - 1], s)
    return v


# ------------------------------------------------------------------ augmentations
def horizontal_flip(img_u8, boxes):
    """cald_helper.py:23-30."""
    t = to_tensor(img_u8)
    width = t.shape[-1]
    b = boxes.clone()
    b[:, [0, 2]] = width - boxes[:, [2, 0]]
    return t.flip(-1), b


def cutout_rects(height, width, boxes, cut_num=2, rng=random, remove_thres=0.4, min_thres=0.1):
    """The accept/reject loop of cald_helper.py:88-132; returns accepted (l, t, r, b) ints.

    ``rng`` must expose ``uniform`` (python's global ``random`` in the reference);
    four draws per try, in the order h, w, left, top.
    """
    rects = []
    boxes = boxes.to(torch.float32)
    for _ in range(50):
        ch = rng.uniform(0.05 * height, 0.2 * height)
        cw = rng.uniform(0.05 * width, 0.2 * width)
        left = rng.uniform(0, width - cw)
        right = left + cw
        top = rng.uniform(0, height - ch)
        bottom = top + ch
        c = torch.tensor([int(left), int(top), int(right), int(bottom)], dtype=torch.float32)
        mx = torch.min(c[2:], boxes[:, 2:])
        mn = torch.max(c[:2], boxes[:, :2])
        inter = torch.clamp(mx - mn, min=0)
        ov = inter[:, 0] * inter[:, 1]
        area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
        r = (ov / area).max().item()
        if r > remove_thres or r < min_thres:
            continue
        rects.append((int(left), int(top), int(right), int(bottom)))
        if len(rects) >= cut_num:
            break
    return rects


def cutout(img_u8, boxes, cut_num=2, rng=random):
    t = to_tensor(img_u8)
    for (l, tp, r, b) in cutout_rects(t.shape[1], t.shape[2], boxes, cut_num, rng):
        t[:, tp:b, l:r] = 0
    return t


def resize(img_u8, boxes, ratio):
    """cald_helper.py:47-53."""
    return to_tensor(pil_oracle.cald_resize_image(img_u8, ratio)), boxes * ratio


def rotate_boxes(boxes, w, h, angle_deg, rot_w, rot_h):
    """Box half of cald_helper.rotate (cald_helper.py:156-222).

    (rot_w, rot_h) = size of the expanded rotated PIL image.
    """
    cx, cy = w / 2, h / 2
    ang = np.radians(angle_deg)
    alpha, beta = np.cos(ang), np.sin(ang)
    m = torch.tensor([[alpha, beta, (1 - alpha) * cx - beta * cy],
                      [-beta, alpha, beta * cx + (1 - alpha) * cy]])
    bw = (boxes[:, 2] - boxes[:, 0]).reshape(-1, 1)
    bh = (boxes[:, 3] - boxes[:, 1]).reshape(-1, 1)
    x1 = boxes[:, 0].reshape(-1, 1)
    y1 = boxes[:, 1].reshape(-1, 1)
    x4 = boxes[:, 2].reshape(-1, 1)
    y4 = boxes[:, 3].reshape(-1, 1)
    corners = torch.stack((x1, y1, x1 + bw, y1, x1, y1 + bh, x4, y4), dim=1).reshape(-1, 2)
    corners = torch.cat((corners, torch.ones(corners.shape[0], 1)), dim=1)
    cos = np.abs(m[0, 0])
    sin = np.abs(m[0, 1])
    nW = int((h * sin) + (w * cos))
    nH = int((h * cos) + (w * sin))
    m[0, 2] += (nW / 2) - cx
    m[1, 2] += (nH / 2) - cy
    rc = torch.mm(m.float(), corners.t()).t().reshape(-1, 8)
    xs = rc[:, [0, 2, 4, 6]]
    ys = rc[:, [1, 3, 5, 7]]
    nb = torch.cat((xs.min(1)[0].reshape(-1, 1), ys.min(1)[0].reshape(-1, 1),
                    xs.max(1)[0].reshape(-1, 1), ys.max(1)[0].reshape(-1, 1)), dim=1)
    sx, sy = rot_w / w, rot_h / h
    nb /= torch.Tensor([sx, sy, sx, sy])
    nb[:, 0] = torch.clamp(nb[:, 0], 0, w)
    nb[:, 1] = torch.clamp(nb[:, 1], 0, h)
    nb[:, 2] = torch.clamp(nb[:, 2], 0, w)
    nb[:, 3] = torch.clamp(nb[:, 3], 0, h)
    return nb


def rotate(img_u8, boxes, angle_deg=5):
    h, w = img_u8.shape[:2]
    out, rw, rh = pil_oracle.cald_rotate_image(img_u8, angle_deg)
    return to_tensor(out), rotate_boxes(boxes, w, h, angle_deg, rw, rh)


def gaussian_noise(img_u8, std):
    """cald_helper.py:72-75 (draws from torch's global CPU generator)."""
    t = to_tensor(img_u8)
    return t + torch.randn(t.size()) * std / 255.0


def salt_pepper(img_u8, prob):
    """cald_helper.py:78-85."""
    t = to_tensor(img_u8)
    noise = torch.rand(t.size())
    salt, pepper = torch.max(t), torch.min(t)
    t[noise < prob / 2] = salt
    t[noise > 1 - prob / 2] = pepper
    return t


COLOR_PERMS = ((0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0))


def color_swap(img_u8, rng=random):
    """cald_helper.py:56-62: one random.randint draw, then a channel permutation of the tensor."""
    swap = COLOR_PERMS[rng.randint(0, len(COLOR_PERMS) - 1)]
    return to_tensor(img_u8)[list(swap), :, :]


def color_adjust(img_u8, factor):
    """cald_helper.py:65-69: PIL brightness -> contrast -> saturation (oracle/pil_oracle.py), then to_tensor."""
    return to_tensor(pil_oracle.cald_color_adjust(img_u8, factor))


# ------------------------------------------------------------------ reduction
def js_divergence(p, q):
    """cald_train.py:211-216: scipy.stats.entropy semantics on float32 vectors.

    entropy(pk, qk) normalises BOTH arguments to sum 1, then sums rel_entr.
    """
    p = np.asarray(p, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    m = (p + q) / 2

    def ent(a, b):
        a = 1.0 * a / np.sum(a, axis=0, keepdims=True)
        b = 1.0 * b / np.sum(b, axis=0, keepdims=True)
        # scipy.special.rel_entr's float32 loop evaluates x*log(x/y) in double and rounds once
        a64, b64 = a.astype(np.float64), b.astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            v = np.where(a64 > 0, a64 * np.log(a64 / b64), np.where(a64 == 0, 0.0, np.inf)).astype(np.float32)
        return np.sum(v, axis=0)

    js = 0.5 * ent(p, m) + 0.5 * ent(q, m)
    if js < 0:
        js = 0
    return js


def pair_consistency(ref, aug_boxes, det, bp):
    """cald_train.py:189-223 for one augmented view -> consistency_img (python float).

    ref: dict with (already sub-sampled) prob_max, scores_cls; aug_boxes: the
    reference boxes mapped into the view; det: the view's detections.
    """
    boxes = det["boxes"]
    if len(boxes) == 0:
        return 0.0
    cons = 1.0
    for ab, rsc, rpm in zip(aug_boxes, ref["scores_cls"], ref["prob_max"]):
        width = torch.min(ab[2], boxes[:, 2]) - torch.max(ab[0], boxes[:, 0])
        height = torch.min(ab[3], boxes[:, 3]) - torch.max(ab[1], boxes[:, 1])
        a_area = (ab[2] - ab[0]) * (ab[3] - ab[1])
        b_area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
        inter = width * height
        iou = inter / (a_area + b_area - inter)
        iou[width < 0] = 0.0
        iou[height < 0] = 0.0
        j = torch.argmax(iou)
        js = js_divergence(rsc.numpy(), det["scores_cls"][j].numpy())
        v = torch.abs(torch.max(iou) + 0.5 * (1 - js) * (rpm + det["prob_max"][j]) - bp).item()
        cons = min(cons, v)
    return cons


def make_views(img_u8, ref_boxes, augs, rng=random):
    """cald_train.py:123-183 for the augmentations that are runnable (the reference's
    'multi_color_adjust' raises NameError, cald_train.py:148).  Returns [(image CHW, boxes)]."""
    views = []
    if 'flip' in augs:
        views.append(horizontal_flip(img_u8, ref_boxes))
    if 'ga' in augs:
        views.append((gaussian_noise(img_u8, 16), ref_boxes))
    if 'multi_ga' in augs:
        for i in range(1, 7):
            views.append((gaussian_noise(img_u8, i * 8), ref_boxes))
    if 'color_adjust' in augs:
        views.append((color_adjust(img_u8, 1.5), ref_boxes))
    if 'color_swap' in augs:
        views.append((color_swap(img_u8, rng), ref_boxes))
    if 'multi_color_adjust' in augs:
        raise NameError("name 'reference_boxes' is not defined")  # cald_train.py:148, as in the reference
    if 'sp' in augs:
        views.append((salt_pepper(img_u8, 0.1), ref_boxes))
    if 'multi_sp' in augs:
        for i in range(1, 7):
            views.append((salt_pepper(img_u8, i * 0.05), ref_boxes))
    if 'cut_out' in augs:
        views.append((cutout(img_u8, ref_boxes, 2, rng), ref_boxes))
    if 'multi_cut_out' in augs:
        for i in range(1, 5):
            views.append((cutout(img_u8, ref_boxes, i, rng), ref_boxes))
    if 'multi_resize' in augs:
        for i in range(7, 10):
            views.append(resize(img_u8, ref_boxes, i * 0.1))
    if 'larger_resize' in augs:
        views.append(resize(img_u8, ref_boxes, 1.2))
    if 'smaller_resize' in augs:
        views.append(resize(img_u8, ref_boxes, 0.8))
    if 'rotation' in augs:
        views.append(rotate(img_u8, ref_boxes, 5))
    return views


def score_image(forward_fn, img_u8, augs, num_cls, bp=1.3, rng=random, trace=None):
    """One iteration of the loop at cald_train.py:101-228 -> (consistency, class vector)."""
    out = forward_fn(to_tensor(img_u8))
    n = len(out["scores"])
    inds = torch.from_numpy(subsample_indices(n)) if n > 40 else None
    ref = {k: (out[k][inds] if inds is not None else out[k])
           for k in ("boxes", "prob_max", "scores_cls", "labels", "scores")}
    cls_rows = [class_max_vector(ref["scores"], ref["labels"], num_cls)]
    if out["boxes"].shape[0] == 0:
        return 0.0, np.mean(cls_rows, axis=0)
    views = make_views(img_u8, ref["boxes"], augs, rng)
    dets = [forward_fn(v) for v, _ in views]
    cons = []
    for det, (_, ab) in zip(dets, views):
        cls_rows.append(class_max_vector(det["scores"], det["labels"], num_cls))
        cons.append(np.mean(pair_consistency(ref, ab, det, bp)))
    if trace is not None:
        trace.update(ref=ref, views=views, dets=dets, per_view=cons, ref_full=out)
    return np.mean(cons), np.mean(np.array(cls_rows), axis=0)


def get_uncertainty(forward_fn, images_u8, augs, num_cls, bp=1.3, seeds=None):
    """Restated cald_train.get_uncertainty.  ``seeds[i]`` (optional) reseeds python's
    ``random`` before image i so that the cutout draws are reproducible per image."""
    cons_all, cls_all = [], []
    for i, img in enumerate(images_u8):
        if seeds is not None:
            random.seed(seeds[i])
        c, v = score_image(forward_fn, img, augs, num_cls, bp)
        cons_all.append(c)
        cls_all.append(v)
    return cons_all, cls_all


# ------------------------------------------------------------------ baseline scorers
def ltc_uncertainty(forward_fn, images_u8):
    """lt_c_train.get_uncertainty (lt_c_train.py:89-121), including its own calcu_iou."""
    def calcu_iou(a, b):
        width = min(a[2], b[2]) - max(a[0], b[0]) + 1
        height = min(a[3], b[3]) - max(a[1], b[1]) + 1
        if width <= 0 or height <= 0:
            return 0
        aa = (a[2] - a[0]) * (a[3] - a[1] + 1)
        ba = (b[2] - b[0]) * (b[3] - b[1] + 1)
        inter = width * height
        return inter / (aa + ba - inter)
    out = []
    for img in images_u8:
        det = forward_fn(to_tensor(img))
        unc = 1.0
        for box, prop, pm in zip(det["boxes"], det["props"], det["prob_max"]):
            unc = min(unc, torch.abs(calcu_iou(box, prop) + pm - 1).item())
        out.append(unc)
    return out


def lsc_stability(forward_fn, images_u8):
    """ls_c_train.get_uncertainty (ls_c_train.py:108-155): six Gaussian views drawn from torch's global generator."""
    out = []
    for img in images_u8:
        det = forward_fn(to_tensor(img))
        boxes, pm = det["boxes"], det["prob_max"]
        if boxes.shape[0] == 0:
            out.append(0.0)
            continue
        if len(boxes) > 30:
            inds = torch.topk(pm, 30)[1]
            boxes, pm = boxes[inds], pm[inds]
        stab = [0.0] * len(boxes)
        u = torch.max(1 - pm).item()
        views = [gaussian_noise(img, i * 8) for i in range(1, 7)]
        for v in views:
            b = forward_fn(v)["boxes"]
            if len(b) == 0:
                continue
            for i, ab in enumerate(boxes):
                width = torch.min(ab[2], b[:, 2]) - torch.max(ab[0], b[:, 0])
                height = torch.min(ab[3], b[:, 3]) - torch.max(ab[1], b[:, 1])
                aa = (ab[2] - ab[0]) * (ab[3] - ab[1])
                ba = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
                inter = width * height
                iou = inter / (aa + ba - inter)
                iou[width < 0] = 0.0
                iou[height < 0] = 0.0
                stab[i] += torch.max(iou).item()
        stab = np.array(stab) / 6.0
        p = pm.numpy()
        out.append(np.sum(p * stab) / np.sum(p) - u)
    return out


# ------------------------------------------------------------------ selection
def cls_kldiv(label_hist_rows, cls_corrs, budget, uniform=False):
    """cald_train.py:234-271 given the per-labeled-image class histograms (rows)."""
    cls_inds = []
    for a in list(np.where(np.sum(cls_corrs, axis=1) == 0)[0]):
        cls_inds.append(a)
    kld = torch.nn.KLDivLoss(reduction='none')
    while len(cls_inds) < budget:
        _c = torch.tensor(np.array(cls_corrs))
        _r = torch.tensor(np.mean(np.array(label_hist_rows), axis=0)).unsqueeze(0)
        if uniform:
            p = torch.nn.functional.softmax(_r + _c, -1)
            q = torch.nn.functional.softmax(torch.ones(_r.shape) / len(_r), -1)
            lm = ((p + q) / 2).log()
            js = torch.sum(kld(lm, p), dim=1) / 2 + torch.sum(kld(lm, q), dim=1) / 2
            js[cls_inds] = 100
            cls_inds.append(torch.argmin(js).item())
        else:
            p = torch.nn.functional.softmax(_r, -1)
            q = torch.nn.functional.softmax(_c, -1)
            lm = ((p + q) / 2).log()
            js = torch.sum(kld(lm, p), dim=1) / 2 + torch.sum(kld(lm, q), dim=1) / 2
            js[cls_inds] = -1
            cls_inds.append(torch.argmax(js).item())
    return cls_inds


def select(uncertainty, cls_corrs, subset, label_hist_rows, budget, mr=1.2, mutual=True, uniform=False):
    """cald_train.py:439-448 (mutual) / 452-455 (--no-mutual) -> newly labeled dataset indices."""
    arg = np.argsort(np.array(uncertainty))
    if not mutual:
        return list(torch.tensor(subset)[arg][:budget].numpy())
    cand = arg[:int(mr * budget)]
    picked = cls_kldiv(label_hist_rows, [cls_corrs[i] for i in cand], budget, uniform)
    return list(torch.tensor(subset)[arg][picked].numpy())
