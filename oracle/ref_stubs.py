"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference on CPU.

This module exists so that ``tests/golden/make_golden.py`` (run in the build
container, where ``/root/reference`` is mounted) can execute the reference's own
``cald_train.get_uncertainty`` / ``cald.cald_helper`` / ``detection.frcnn_la`` and
write golden vectors.  It never travels into the product path and nothing under
``cald_b200/`` imports it.  On the GPU box ``/root/reference`` does not exist and
``available()`` returns False.

Stub list follows SURVEY.md section 8(c): modules the reference imports at top
level that are absent (or renamed) on the installed stack are replaced by empty
modules; ``Tensor.cuda`` becomes the identity and ``torch.cuda.synchronize`` a
no-op so the hard-coded ``.cuda()`` calls (cald_train.py:107,125-183;
cald_helper.py:116,181,193,218) run on CPU.
"""
import os
import sys
import types
import argparse

REF_ROOT = os.environ.get("CALD_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "cald_train.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = {}


def load(bp=1.3, uniform=False, mr=1.2):
    """Return the reference's ``cald_train`` module (imported once)."""
    if "cald_train" in _loaded:
        ct = _loaded["cald_train"]
        ct.args = argparse.Namespace(bp=bp, uniform=uniform, mr=mr)
        return ct
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import torch
    import torchvision

    # torchvision.models.utils (frcnn_la.py:20, retinanet_cal.py:10)
    _mod("torchvision.models.utils", load_state_dict_from_url=torch.hub.load_state_dict_from_url)
    # matplotlib (cald_train.py:11, cald_helper.py:268)
    if "matplotlib" not in sys.modules:
        mpl = _mod("matplotlib")
        mpl.pyplot = _mod("matplotlib.pyplot")
    # pycocotools (coco_utils.py:9-10, coco_eval.py:10-12)
    pc = _mod("pycocotools")
    pc.mask = _mod("pycocotools.mask")
    pc.coco = _mod("pycocotools.coco", COCO=object)
    pc.cocoeval = _mod("pycocotools.cocoeval", COCOeval=object)
    _mod("terminaltables", AsciiTable=object)
    mm = _mod("mmcv")
    mm.utils = _mod("mmcv.utils", print_log=print)
    _mod("torch._six", string_classes=(str,))
    try:
        import cv2  # noqa: F401
    except Exception:
        _mod("cv2")
    # mobilenetv3.py:9-10 imports symbols that no longer exist
    import torchvision.models.mobilenet as mb
    for n in ("ConvBNReLU", "_make_divisible", "model_urls"):
        if not hasattr(mb, n):
            setattr(mb, n, object if n != "model_urls" else {})
    # CPU execution of hard-coded .cuda()
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import cald_train  # noqa: E402
    cald_train.args = argparse.Namespace(bp=bp, uniform=uniform, mr=mr)
    _loaded["cald_train"] = cald_train
    return cald_train


def frcnn_module():
    load()
    import detection.frcnn_la as m
    return m


def retinanet_module():
    load()
    import detection.retinanet_cal as m
    return m


def baseline_module(name):
    """'lt_c_train' or 'ls_c_train', unmodified.  ls_c_train.py:52 imports ``cal4od.cal4od_helper``, a module that does
    not exist in the reference tree; it is aliased to cald.cald_helper, the file the script was
    evidently written against (it only needs GaussianNoise)."""
    load()
    import importlib
    import cald.cald_helper as hp
    pkg = _mod("cal4od")
    pkg.cal4od_helper = hp
    sys.modules["cal4od.cal4od_helper"] = hp
    return importlib.import_module(name)


def helper_module():
    load()
    import cald.cald_helper as m
    return m
