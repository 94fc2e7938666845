"""Host-side mirror of the reference's scoring API.

``get_uncertainty(task_model, unlabeled_loader, augs, num_cls)`` has the signature and the
return value of cald_train.py:91-231 and ``select(...)`` / ``cls_kldiv(...)`` reproduce the
selection code at cald_train.py:234-271 and 439-457, so the active-learning cycle can call
them unchanged.  Underneath, every image is scored by the CUDA engine through the C ABI
(include/cald_b200.h); there is no torch op and no CPU fallback on the scoring path.  torch
is only used to read ``task_model.state_dict()`` and (multi-GPU) for the final all-gather.
"""
import random

import numpy as np

from . import engine as _eng

_AUG_WHITELIST = ['flip', 'multi_ga', 'color_adjust', 'color_swap', 'multi_color_adjust', 'multi_sp', 'cut_out',
                  'multi_cut_out', 'multi_resize', 'larger_resize', 'smaller_resize', 'rotation', 'ga', 'sp']

# the reference reads a module-global ``args`` (cald_train.py:220, 254); mirror it as module state
bp = 1.3
uniform = False
mr = 1.2

_engine_cache = {}


def _aug_kinds(augs):
    """Reference names -> (kind, param) views in reference order (cald_train.py:123-183)."""
    for a in augs:
        if a not in _AUG_WHITELIST:
            print('{} is not in the pre-set augmentations!'.format(a))  # cald_train.py:95
    if 'multi_color_adjust' in augs:
        # cald_train.py:145-148 appends `reference_boxes`, a name that does not exist: the reference itself raises here
        raise NameError("name 'reference_boxes' is not defined (cald_train.py:148: 'multi_color_adjust' is "
                        "unreachable in the reference)")
    return _eng.expand_augs(augs)


def _model_config(task_model, num_cls):
    """Read depth / sizes from a torchvision-style detector (FRCNN_Feature, frcnn_la.py:146-235)."""
    t = getattr(task_model, "transform", None)
    min_size = int(t.min_size[-1]) if t is not None else 800
    max_size = int(t.max_size) if t is not None else 1333
    sd = task_model.state_dict()
    depth = 101 if any(k.startswith("backbone.body.layer3.22.") for k in sd) else 50
    # retinanet_resnet50_fpn_cal (retinanet_cal.py:584-625) is recognised by its head keys
    arch_id = _eng.ARCH_RETINANET if "head.classification_head.cls_logits.weight" in sd else _eng.ARCH_FRCNN
    return depth, min_size, max_size, sd, arch_id


def engine_for(task_model, num_cls, device=0, **kw):
    """Build (or reuse) the engine for this model object; weights are re-read on every call of
    get_uncertainty because the AL cycle retrains the model in between (cald_train.py:409-411)."""
    depth, mn, mx, sd, arch_id = _model_config(task_model, num_cls)
    key = (arch_id, depth, num_cls, mn, mx, device, tuple(sorted(kw.items())))
    eng = _engine_cache.get(key)
    if eng is None:
        eng = _eng.Engine(depth=depth, num_classes=num_cls, min_size=mn, max_size=mx, device=device,
                          arch_id=arch_id, **kw)
        _engine_cache[key] = eng
    eng.load_state_dict(sd)
    return eng


def _to_u8(image):
    """PIL RGB image (what dataset_aug yields, cald_train.py:289) or HxWx3 u8 array -> contiguous u8."""
    a = np.asarray(image)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("expected an RGB PIL image / HxWx3 uint8 array")
    return np.ascontiguousarray(a)


def _draw_noise(images, views):
    """torch CPU-generator draws in the order the reference makes them (cald_helper.py:74, 80): per image, per
    noise view: torch.randn(image.size()) for GaussianNoise, torch.rand(image.size()) for SaltPepperNoise."""
    import torch
    planes = []
    for im in images:
        size = (3, im.shape[0], im.shape[1])
        for kind, _ in views:
            if kind == _eng.AUG_GAUSS:
                planes.append(torch.randn(size).numpy())
            elif kind == _eng.AUG_SALTPEPPER:
                planes.append(torch.rand(size).numpy())
    return planes


def score_images(eng, images, augs, chunk=64):
    """Score u8 images with an existing engine.  Consumes python's global ``random`` stream exactly as
    cald_helper.cutout would (4 uniforms per try, data-dependent number of tries) and torch's global CPU
    generator exactly as GaussianNoise / SaltPepperNoise would."""
    views = _aug_kinds(augs)
    n_cut = sum(1 for k, _ in views if k == _eng.AUG_CUTOUT)
    n_swap = sum(1 for k, _ in views if k == _eng.AUG_COLOR_SWAP)
    has_noise = any(k in _eng.NOISE_KINDS for k, _ in views)
    if n_swap and n_cut:
        # ColorSwap's random.randint precedes the same image's cutout draws in python's RNG stream and the number of
        # cutout draws is data dependent: keep the stream exact by scoring one image per call
        chunk = 1
    cons_all, cls_all = [], []
    for pos in range(0, len(images), chunk):
        batch = images[pos:pos + chunk]
        # cald_helper.ColorSwap: perms[random.randint(0, len(perms) - 1)], one draw per (image, swap view)
        swaps = [random.randint(0, 5) for _ in range(n_swap * len(batch))] if n_swap else None
        u = None
        if n_cut:
            state = random.getstate()
            u = np.array([random.random() for _ in range(200 * n_cut * len(batch))], dtype=np.float64)
        noise = _draw_noise(batch, views) if has_noise else None
        cons, cls, used = eng.score(batch, views, bp, u, noise, swaps)
        if n_cut:
            random.setstate(state)
            for _ in range(used):
                random.random()
        cons_all.extend(float(c) for c in cons)
        cls_all.extend(np.array(r, dtype=np.float64) for r in cls)
    return cons_all, cls_all


def get_uncertainty(task_model, unlabeled_loader, augs, num_cls, device=0, **engine_kw):
    """Drop-in for cald_train.get_uncertainty (cald_train.py:91-231).

    unlabeled_loader yields ``(tuple_of_PIL_images, tuple_of_targets)`` with batch size 1
    (cald_train.py:370-371).  Returns ``(consistency_all, cls_all)``: list of float and list
    of float64 arrays of shape (num_cls - 1,), in loader order.
    """
    eng = engine_for(task_model, num_cls, device, **engine_kw)
    images = []
    for imgs, _ in unlabeled_loader:
        for image in imgs:
            images.append(_to_u8(image))
    return score_images(eng, images, augs)


def _loader_images(unlabeled_loader):
    """Images of a reference-style loader (tuple_of_images, tuple_of_targets) as u8 HWC arrays.  The baseline scripts
    feed ToTensor()'d float tensors (lt_c_train.py:109-111); those came from u8 pixels, so x * 255 is exact."""
    out = []
    for imgs, _ in unlabeled_loader:
        for image in imgs:
            if hasattr(image, "detach"):  # torch CHW float in [0, 1]
                a = image.detach().cpu().numpy()
                out.append(np.ascontiguousarray(np.rint(a * 255.0).astype(np.uint8).transpose(1, 2, 0)))
            else:
                out.append(_to_u8(image))
    return out


def lt_c_uncertainty(task_model, unlabeled_loader, device=0, chunk=64, **engine_kw):
    """Drop-in for lt_c_train.get_uncertainty (lt_c_train.py:105-121): list of float, loader order."""
    eng = engine_for(task_model, _num_classes(task_model), device, **engine_kw)
    images = _loader_images(unlabeled_loader)
    out = []
    for pos in range(0, len(images), chunk):
        out.extend(float(v) for v in eng.score_ltc(images[pos:pos + chunk]))
    return out


def ls_c_uncertainty(task_model, unlabeled_loader, aves=None, device=0, chunk=16, **engine_kw):
    """Drop-in for ls_c_train.get_uncertainty (ls_c_train.py:108-155): list of float, loader order.  Draws the six
    torch.randn planes per image from torch's global CPU generator in the reference's order."""
    import torch
    eng = engine_for(task_model, _num_classes(task_model), device, **engine_kw)
    images = _loader_images(unlabeled_loader)
    out = []
    for pos in range(0, len(images), chunk):
        batch = images[pos:pos + chunk]
        noise = []
        for im in batch:
            # NOTE: the reference draws nothing for an image without reference detections (ls_c_train.py:118-120);
            # the engine cannot know that before the forward, so the torch stream differs after such an image.
            noise += [torch.randn((3, im.shape[0], im.shape[1])).numpy() for _ in range(6)]
        out.extend(float(v) for v in eng.score_lsc(batch, noise))
    return out


class EngineModel:
    """The detections-only mode of the engine behind the call shape the reference's evaluation loops use
    (detection/engine.py:85-158 voc_evaluate, 177-256 coco_evaluate: ``model.eval(); outputs = model(images)`` with
    ``images`` a list of CHW float tensors in [0, 1]).  Returns the reference's output dicts (frcnn_la.py:131-141 /
    retinanet_cal.py:479-485, without 'features') as CPU torch tensors -- SURVEY.md 8(f) item 3."""

    def __init__(self, task_model, device=0, **engine_kw):
        self.engine = engine_for(task_model, _num_classes(task_model), device, **engine_kw)

    def eval(self):
        return self

    def __call__(self, images):
        import torch
        u8 = []
        for im in images:
            if hasattr(im, "detach"):
                a = im.detach().cpu().numpy()
                u8.append(np.ascontiguousarray(np.rint(a * 255.0).astype(np.uint8).transpose(1, 2, 0)))
            else:
                u8.append(_to_u8(im))
        return [{k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in d.items()} for d in self.engine.detect(u8)]


def _num_classes(task_model):
    sd = task_model.state_dict()
    if "roi_heads.box_predictor.cls_score.weight" in sd:
        return int(sd["roi_heads.box_predictor.cls_score.weight"].shape[0])
    return int(sd["head.classification_head.cls_logits.weight"].shape[0]) // 9


def cls_kldiv(labeled_loader, cls_corrs, budget, cycle=0):
    """The class-balance stage of cald_train.py:234-271, same picks, without its O(budget) recomputation.

    The reference recomputes JS(softmax(mean label histogram) || softmax(class vector)) inside its while loop, but
    the histogram update is commented out there (cald_train.py:270), so the divergences never change: they are
    computed once here (same torch ops, float64) and the loop only masks what was already picked.
    """
    import torch
    n_cls = cls_corrs[0].shape[0]
    hist = []
    for _, targets in labeled_loader:
        for target in targets:
            row = [0] * n_cls
            for l in target['labels']:
                row[l - 1] += 1
            hist.append(row)
    picked = [int(a) for a in np.where(np.sum(cls_corrs, axis=1) == 0)[0]]  # all-zero class vectors go first (246-247)
    if len(picked) >= budget:
        return picked
    corr = torch.tensor(np.array(cls_corrs))
    mean_hist = torch.tensor(np.mean(np.array(hist), axis=0)).unsqueeze(0)
    kld = torch.nn.KLDivLoss(reduction='none')
    if uniform:   # args.uniform (cald_train.py:254-261): closest to the uniform distribution first
        p = torch.nn.functional.softmax(mean_hist + corr, -1)
        q = torch.nn.functional.softmax(torch.ones(mean_hist.shape) / len(mean_hist), -1)
        fill, pick = 100, torch.argmin
    else:         # default (262-269): most different from the labeled set's class distribution first
        p = torch.nn.functional.softmax(mean_hist, -1)
        q = torch.nn.functional.softmax(corr, -1)
        fill, pick = -1, torch.argmax
    log_mean = ((p + q) / 2).log()
    js = torch.sum(kld(log_mean, p), dim=1) / 2 + torch.sum(kld(log_mean, q), dim=1) / 2
    while len(picked) < budget:
        js[picked] = fill
        picked.append(int(pick(js).item()))
    return picked


def select(uncertainty, cls_corrs, subset, labeled_loader, budget_num, cycle=0, mutual=True):
    """The inline selection of cald_train.py:439-448 (mutual) / 452-455 (--no-mutual).

    Returns the dataset indices to move from the unlabeled to the labeled set.
    """
    import torch
    arg = np.argsort(np.array(uncertainty))
    if not mutual:
        return list(torch.tensor(subset)[arg][:budget_num].numpy())
    cand = arg[:int(mr * budget_num)]
    picked = cls_kldiv(labeled_loader, [cls_corrs[i] for i in cand], budget_num, cycle)
    return list(torch.tensor(subset)[arg][picked].numpy())
