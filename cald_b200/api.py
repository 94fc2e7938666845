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


def _weights_fingerprint(sd):
    """Cheap identity of a state_dict's CONTENT: torch bumps ``tensor._version`` on every in-place update (optimizer
    steps, load_state_dict), so (storage address, version) of every tensor changes whenever the weights do.  Anything
    that is not a torch tensor (numpy weights in the tests) has no version counter: None = always reload."""
    fp = []
    for k, v in sd.items():
        if not hasattr(v, "_version") or not hasattr(v, "data_ptr"):
            return None
        fp.append((k, v.data_ptr(), v._version, tuple(v.shape)))
    return hash(tuple(fp))


def engine_for(task_model, num_cls, device=0, **kw):
    """The engine for this detector configuration (one per configuration and GPU, built on first use) holding THIS
    model's current weights.  The AL cycle retrains the model between scoring calls (cald_train.py:409-411) and one
    configuration may serve several model objects, so the weights are re-read whenever the state_dict's fingerprint
    differs from what the engine holds -- never silently stale, not re-uploaded when nothing changed."""
    depth, mn, mx, sd, arch_id = _model_config(task_model, num_cls)
    key = (arch_id, depth, num_cls, mn, mx, device, tuple(sorted(kw.items())))
    eng = _engine_cache.get(key)
    if eng is None:
        eng = _eng.Engine(depth=depth, num_classes=num_cls, min_size=mn, max_size=mx, device=device,
                          arch_id=arch_id, **kw)
        _engine_cache[key] = eng
    fp = _weights_fingerprint(sd)
    if fp is None or getattr(eng, "_weights_fp", None) != fp:
        eng.load_state_dict(sd)
        eng._weights_fp = fp
    return eng


def close_engines():
    """Destroy every cached engine (each holds a device arena of up to 64 GB and a weight replica)."""
    for eng in _engine_cache.values():
        eng.close()
    _engine_cache.clear()


def _to_u8(image):
    """PIL RGB image (what dataset_aug yields, cald_train.py:289) or HxWx3 u8 array -> contiguous u8."""
    a = np.asarray(image)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("expected an RGB PIL image / HxWx3 uint8 array")
    return np.ascontiguousarray(a)


def _draw_noise_one(im, views):
    """torch CPU-generator draws for ONE image in the order the reference makes them (cald_helper.py:74, 80): per noise
    view torch.randn(image.size()) for GaussianNoise, torch.rand(image.size()) for SaltPepperNoise."""
    import torch
    size = (3, im.shape[0], im.shape[1])
    planes = []
    for kind, _ in views:
        if kind == _eng.AUG_GAUSS:
            planes.append(torch.randn(size).numpy())
        elif kind == _eng.AUG_SALTPEPPER:
            planes.append(torch.rand(size).numpy())
    return planes


def score_images(eng, images, augs, chunk=None):
    """Score u8 images with an existing engine.  Consumes python's global ``random`` stream exactly as
    cald_helper.cutout / ColorSwap would (4 uniforms per cutout try, data-dependent number of tries; one randint per
    swap view) and torch's global CPU generator exactly as GaussianNoise / SaltPepperNoise would.

    The reference draws NOTHING for an image whose reference prediction is empty (it ``continue``s at
    cald_train.py:118-121, before any augmentation is built).  Cutout draws are data dependent anyway and are resolved
    on the device; the swap / noise draws are made optimistically here, before the forward pass, and when a chunk turns
    out to contain an empty image both generators are rewound to where they stood before that image's draws and the
    rest of the chunk is scored again -- so the streams always end where the reference leaves them."""
    import torch
    views = _aug_kinds(augs)
    n_cut = sum(1 for k, _ in views if k == _eng.AUG_CUTOUT)
    n_swap = sum(1 for k, _ in views if k == _eng.AUG_COLOR_SWAP)
    has_noise = any(k in _eng.NOISE_KINDS for k, _ in views)
    if chunk is None:
        chunk = 2 * eng.images_per_chunk(len(views))  # several engine passes per call: the upload pipeline overlaps them
    if (n_swap or has_noise) and n_cut:
        # swap / noise draws of an image precede the same image's data-dependent cutout draws in the reference's
        # streams: keep both exact by scoring one image per call
        chunk = 1
    speculative = bool(n_swap or has_noise) and len(views) > 0
    cons_all, cls_all = [], []
    pos = 0
    while pos < len(images):
        batch = images[pos:pos + chunk]
        marks, swaps, noise = [], [], []
        for im in batch:
            if speculative:
                marks.append((random.getstate() if n_swap else None, torch.get_rng_state() if has_noise else None))
            # cald_helper.ColorSwap: perms[random.randint(0, len(perms) - 1)], one draw per swap view
            swaps += [random.randint(0, 5) for _ in range(n_swap)]
            if has_noise:
                noise += _draw_noise_one(im, views)
        u = None
        if n_cut:
            state = random.getstate()
            u = np.array([random.random() for _ in range(200 * n_cut * len(batch))], dtype=np.float64)
        cons, cls, used = eng.score(batch, views, bp, u, noise if has_noise else None, swaps if n_swap else None)
        if n_cut:
            random.setstate(state)
            for _ in range(used):
                random.random()
        good = len(batch)
        if speculative:
            ref_counts = eng.last_ref_counts(len(batch))
            for i in range(len(batch)):
                if ref_counts[i] == 0:
                    py_state, torch_state = marks[i]
                    if py_state is not None:
                        random.setstate(py_state)
                    if torch_state is not None:
                        torch.set_rng_state(torch_state)
                    good = i + 1  # the empty image's own result (0.0, zeros) does not depend on the draws
                    break
        cons_all.extend(float(c) for c in cons[:good])
        cls_all.extend(np.array(r, dtype=np.float64) for r in cls[:good])
        pos += good
    return cons_all, cls_all


def get_uncertainty_files(task_model, paths, augs, num_cls, device=0, **engine_kw):
    """get_uncertainty over a pool given as JPEG FILE PATHS (or bytes objects) instead of a DataLoader of PIL images:
    the pool ingest of SURVEY.md 8(f).  Replaces ``Image.open(path).convert('RGB')`` on DataLoader workers
    (detection/voc_utils.py:52-58, coco_utils.py:209-220): the files are read on the host, decoded on the device (bit
    for bit Pillow's pixels) and scored, in loader (= list) order.  Same return value as get_uncertainty.  Consumes
    python's ``random`` stream like the reference (cutout, ColorSwap); the noise augmentations need a PIL-side image
    size before drawing and are only available through get_uncertainty."""
    eng = engine_for(task_model, num_cls, device, **engine_kw)
    views = _aug_kinds(augs)
    if any(k in _eng.NOISE_KINDS for k, _ in views):
        raise ValueError("noise augmentations ('ga', 'sp', ...) are not available on the file path; use get_uncertainty")
    n_cut = sum(1 for k, _ in views if k == _eng.AUG_CUTOUT)
    n_swap = sum(1 for k, _ in views if k == _eng.AUG_COLOR_SWAP)
    # files are small (~150 KB): hand the engine eight chunks per call, so that only one chunk in eight has its
    # decode latency (the entropy walk of its images + the device decode) exposed instead of hidden behind the
    # previous chunk's forward passes
    group = 1 if (n_swap and n_cut) else 8 * eng.images_per_chunk(max(1, len(views)))

    def read(p):
        if isinstance(p, (bytes, bytearray, memoryview)):
            return bytes(p)
        with open(p, "rb") as f:
            return f.read()
    cons_all, cls_all = [], []
    batches = ([read(p) for p in paths[i:i + group]] for i in range(0, len(paths), group))
    for files in _prefetched(batches, 2):
        pos = 0
        while pos < len(files):
            part = files[pos:]
            marks = [None] * len(part)
            swaps = []
            for i in range(len(part)):
                if n_swap:
                    marks[i] = random.getstate()
                    swaps += [random.randint(0, 5) for _ in range(n_swap)]
            u = None
            if n_cut:
                state = random.getstate()
                u = np.array([random.random() for _ in range(200 * n_cut * len(part))], dtype=np.float64)
            cons, cls, used, _, _ = eng.score_jpeg(part, views, bp, u, swaps if n_swap else None)
            if n_cut:
                random.setstate(state)
                for _ in range(used):
                    random.random()
            good = len(part)
            if n_swap:   # rewind over an image without reference detections (see score_images)
                ref_counts = eng.last_ref_counts(len(part))
                for i in range(len(part)):
                    if ref_counts[i] == 0:
                        random.setstate(marks[i])
                        good = i + 1
                        break
            cons_all.extend(float(c) for c in cons[:good])
            cls_all.extend(np.array(r, dtype=np.float64) for r in cls[:good])
            pos += good
    return cons_all, cls_all


def _prefetched(iterable, depth):
    """Iterate ``iterable`` in a background thread, ``depth`` items ahead (the C call releases the GIL, so the
    DataLoader's collation / PIL decode of the next images overlaps the current scoring pass)."""
    import queue
    import threading
    q = queue.Queue(maxsize=max(1, depth))
    end = object()

    def work():
        try:
            for item in iterable:
                q.put(item)
            q.put(end)
        except BaseException as ex:  # re-raised in the consumer
            q.put(ex)
    threading.Thread(target=work, daemon=True).start()
    while True:
        item = q.get()
        if item is end:
            return
        if isinstance(item, BaseException):
            raise item
        yield item


def get_uncertainty(task_model, unlabeled_loader, augs, num_cls, device=0, **engine_kw):
    """Drop-in for cald_train.get_uncertainty (cald_train.py:91-231).

    unlabeled_loader yields ``(tuple_of_PIL_images, tuple_of_targets)`` with batch size 1
    (cald_train.py:370-371).  Returns ``(consistency_all, cls_all)``: list of float and list
    of float64 arrays of shape (num_cls - 1,), in loader order.

    The loader is consumed as a stream, like the reference consumes it (cald_train.py:101): a group of a few engine
    passes' worth of images is held in host memory at a time, never the pool, and the next group is pulled from the
    loader while the current one is on the GPU.
    """
    eng = engine_for(task_model, num_cls, device, **engine_kw)
    group = 2 * eng.images_per_chunk(max(1, len(_eng.expand_augs(augs))))
    cons_all, cls_all, pending = [], [], []

    def flush():
        c, v = score_images(eng, pending, augs)
        cons_all.extend(c)
        cls_all.extend(v)
        del pending[:]
    for imgs, _ in _prefetched(unlabeled_loader, 2 * group):
        for image in imgs:
            pending.append(_to_u8(image))
        if len(pending) >= group:
            flush()
    if pending:
        flush()
    return cons_all, cls_all


def _image_u8(image):
    """One loader image as a u8 HWC array.  The baseline scripts feed ToTensor()'d float tensors
    (lt_c_train.py:109-111); those came from u8 pixels, so x * 255 is exact."""
    if hasattr(image, "detach"):  # torch CHW float in [0, 1]
        a = image.detach().cpu().numpy()
        return np.ascontiguousarray(np.rint(a * 255.0).astype(np.uint8).transpose(1, 2, 0))
    return _to_u8(image)


def _loader_groups(unlabeled_loader, group):
    """Stream a reference-style loader (tuple_of_images, tuple_of_targets) as lists of ``group`` u8 images."""
    pending = []
    for imgs, _ in _prefetched(unlabeled_loader, 2 * group):
        for image in imgs:
            pending.append(_image_u8(image))
        if len(pending) >= group:
            yield pending
            pending = []
    if pending:
        yield pending


def lt_c_uncertainty(task_model, unlabeled_loader, device=0, chunk=64, **engine_kw):
    """Drop-in for lt_c_train.get_uncertainty (lt_c_train.py:105-121): list of float, loader order."""
    eng = engine_for(task_model, _num_classes(task_model), device, **engine_kw)
    out = []
    for batch in _loader_groups(unlabeled_loader, chunk):
        out.extend(float(v) for v in eng.score_ltc(batch))
    return out


def ls_c_uncertainty(task_model, unlabeled_loader, aves=None, device=0, chunk=16, **engine_kw):
    """Drop-in for ls_c_train.get_uncertainty (ls_c_train.py:108-155): list of float, loader order.  Draws the six
    torch.randn planes per image from torch's global CPU generator in the reference's order."""
    import torch
    eng = engine_for(task_model, _num_classes(task_model), device, **engine_kw)
    out = []
    for batch in _loader_groups(unlabeled_loader, chunk):
        pos = 0
        while pos < len(batch):
            part = batch[pos:]
            marks, noise = [], []
            for im in part:
                marks.append(torch.get_rng_state())
                noise += [torch.randn((3, im.shape[0], im.shape[1])).numpy() for _ in range(6)]
            vals = eng.score_lsc(part, noise)
            good = len(part)
            # the reference draws nothing for an image without reference detections (ls_c_train.py:118-120, before
            # the noise loop): rewind the generator to before the first such image's draws and score the rest again
            ref_counts = eng.last_ref_counts(len(part))
            for i in range(len(part)):
                if ref_counts[i] == 0:
                    torch.set_rng_state(marks[i])
                    good = i + 1
                    break
            out.extend(float(v) for v in vals[:good])
            pos += good
    return out


class EngineModel:
    """The detections-only mode of the engine behind the call shape the reference's evaluation loops use
    (detection/engine.py:85-158 voc_evaluate, 177-256 coco_evaluate: ``model.eval(); outputs = model(images)`` with
    ``images`` a list of CHW float tensors in [0, 1]).  Returns the reference's output dicts (frcnn_la.py:131-141 /
    retinanet_cal.py:479-485, without 'features') as CPU torch tensors -- SURVEY.md 8(f) item 3."""

    def __init__(self, task_model, device=0, **engine_kw):
        self.task_model, self.device, self.engine_kw = task_model, device, engine_kw
        self.engine = engine_for(task_model, _num_classes(task_model), device, **engine_kw)

    def eval(self):
        return self

    def __call__(self, images):
        import torch
        # engines are shared per configuration and the model is retrained between evaluations: engine_for re-reads
        # the weights whenever they differ from what the engine holds (fingerprint of the state_dict)
        self.engine = engine_for(self.task_model, _num_classes(self.task_model), self.device, **self.engine_kw)
        u8 = [_image_u8(im) for im in images]
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


def _label_histogram_mean(labeled_loader, n_cls):
    """np.mean of the per-image label histograms of the labeled set (cald_train.py:237-242, 253)."""
    hist = []
    for _, targets in labeled_loader:
        for target in targets:
            row = [0] * n_cls
            for l in target['labels']:
                row[l - 1] += 1
            hist.append(row)
    return np.mean(np.array(hist), axis=0)


def select(uncertainty, cls_corrs, subset, labeled_loader, budget_num, cycle=0, mutual=True, engine=None):
    """The inline selection of cald_train.py:439-448 (mutual) / 452-455 (--no-mutual).

    Returns the dataset indices to move from the unlabeled to the labeled set.  With ``engine`` the argsort, the
    candidate cut and cls_kldiv run on that engine's GPU (cald_select) -- same picks, same order.
    """
    import torch
    if engine is not None and mutual:
        cls = np.asarray(cls_corrs, dtype=np.float64)
        pos = engine.select(uncertainty, cls, _label_histogram_mean(labeled_loader, cls.shape[1]), budget_num,
                            int(mr * budget_num), uniform)
        return list(torch.tensor(subset)[torch.from_numpy(pos.astype(np.int64))].numpy())
    arg = np.argsort(np.array(uncertainty))
    if not mutual:
        return list(torch.tensor(subset)[arg][:budget_num].numpy())
    cand = arg[:int(mr * budget_num)]
    picked = cls_kldiv(labeled_loader, [cls_corrs[i] for i in cand], budget_num, cycle)
    return list(torch.tensor(subset)[arg][picked].numpy())
