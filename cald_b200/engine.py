"""ctypes binding of the public C ABI (include/cald_b200.h) -- one Engine per GPU."""
import ctypes
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_longlong, c_size_t, c_uint8, c_void_p

import numpy as np

from . import arch
from ._lib import CaldError, lib

ARCH_FRCNN, ARCH_RETINANET = 0, 1
PREC_F16X3, PREC_F16 = 0, 1   # split-half x3 (fp32-faithful, the product mode) / single-pass half
PREC_BF16X3, PREC_BF16 = PREC_F16X3, PREC_F16  # round-1 names
CONV_TCGEN05, CONV_SIMT = 0, 1
AUG_FLIP, AUG_CUTOUT, AUG_RESIZE, AUG_ROTATION, AUG_GAUSS, AUG_SALTPEPPER = 0, 1, 2, 3, 4, 5
AUG_COLOR_ADJUST, AUG_COLOR_SWAP = 6, 7
AUG_SMALLER_RESIZE = AUG_RESIZE
NOISE_KINDS = (AUG_GAUSS, AUG_SALTPEPPER)


class Aug(ctypes.Structure):
    """cald_aug: one augmented view = (kind, the reference's per-call argument)."""
    _fields_ = [("kind", c_int), ("param", c_double)]


def expand_augs(names):
    """Reference augmentation names (cald_train.py:93-94) -> list of (kind, param) views in the exact order
    get_uncertainty appends them (cald_train.py:123-183), with the reference's literal arguments."""
    v = []
    if 'flip' in names:
        v.append((AUG_FLIP, 0.0))
    if 'ga' in names:
        v.append((AUG_GAUSS, 16.0))
    if 'multi_ga' in names:
        v += [(AUG_GAUSS, float(i * 8)) for i in range(1, 7)]
    if 'color_adjust' in names:
        v.append((AUG_COLOR_ADJUST, 1.5))
    if 'color_swap' in names:
        v.append((AUG_COLOR_SWAP, 0.0))
    if 'sp' in names:
        v.append((AUG_SALTPEPPER, 0.1))
    if 'multi_sp' in names:
        v += [(AUG_SALTPEPPER, i * 0.05) for i in range(1, 7)]
    if 'cut_out' in names:
        v.append((AUG_CUTOUT, 2.0))
    if 'multi_cut_out' in names:
        v += [(AUG_CUTOUT, float(i)) for i in range(1, 5)]
    if 'multi_resize' in names:
        v += [(AUG_RESIZE, i * 0.1) for i in range(7, 10)]
    if 'larger_resize' in names:
        v.append((AUG_RESIZE, 1.2))
    if 'smaller_resize' in names:
        v.append((AUG_RESIZE, 0.8))
    if 'rotation' in names:
        v.append((AUG_ROTATION, 5.0))
    return v


SUPPORTED_AUGS = ('flip', 'ga', 'multi_ga', 'color_adjust', 'color_swap', 'sp', 'multi_sp', 'cut_out',
                  'multi_cut_out', 'multi_resize', 'larger_resize', 'smaller_resize', 'rotation')


class Config(ctypes.Structure):
    _fields_ = [("arch", c_int), ("depth", c_int), ("num_classes", c_int), ("min_size", c_int),
                ("max_size", c_int), ("rpn_pre_nms_top_n", c_int), ("rpn_post_nms_top_n", c_int),
                ("rpn_nms_thresh", c_float), ("box_score_thresh", c_float), ("box_nms_thresh", c_float),
                ("box_detections_per_img", c_int), ("retina_max_detections", c_int), ("device", c_int), ("precision", c_int),
                ("conv_impl", c_int), ("max_views_per_pass", c_int), ("workspace_bytes", c_size_t),
                ("debug", c_int)]


def _as_u8(im):
    # np.ascontiguousarray was observed to copy arrays that view page-locked torch memory (31 ms per 8 images):
    # pass anything that already is a contiguous u8 ndarray through untouched
    if isinstance(im, np.ndarray) and im.dtype == np.uint8 and im.flags["C_CONTIGUOUS"]:
        return im
    return np.ascontiguousarray(im, dtype=np.uint8)


def _u8_list(images):
    imgs = [_as_u8(im) for im in images]
    for im in imgs:
        if im.ndim != 3 or im.shape[2] != 3:
            raise ValueError("images must be HxWx3 uint8 (RGB)")
    n = len(imgs)
    ptrs = (POINTER(c_uint8) * n)(*[im.ctypes.data_as(POINTER(c_uint8)) for im in imgs])
    hs = (c_int * n)(*[im.shape[0] for im in imgs])
    ws = (c_int * n)(*[im.shape[1] for im in imgs])
    return imgs, ptrs, hs, ws


class Engine:
    """The scoring engine for one detector on one GPU."""

    def __init__(self, depth=50, num_classes=21, min_size=600, max_size=1000, device=0, precision=PREC_F16X3,
                 conv_impl=CONV_TCGEN05, max_views_per_pass=0, workspace_bytes=0, debug=False,
                 arch_id=ARCH_FRCNN, **overrides):
        L = lib()
        L.cald_create.argtypes = [POINTER(Config), POINTER(c_void_p)]
        L.cald_last_error.restype = c_char_p
        L.cald_last_error.argtypes = [c_void_p]
        L.cald_destroy.argtypes = [c_void_p]
        L.cald_views_per_pass.argtypes = [c_void_p]
        cfg = Config()
        L.cald_config_default(ctypes.byref(cfg), arch_id, depth, num_classes, min_size, max_size)
        cfg.device, cfg.precision, cfg.conv_impl = device, precision, conv_impl
        cfg.max_views_per_pass, cfg.workspace_bytes, cfg.debug = max_views_per_pass, workspace_bytes, int(debug)
        for k, v in overrides.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        self.num_classes = num_classes
        self.depth = depth
        self.arch_id = arch_id
        self._h = c_void_p()
        self._L = L
        if L.cald_create(ctypes.byref(cfg), ctypes.byref(self._h)) != 0:
            raise CaldError(L.cald_last_error(None).decode())

    def _check(self, rc):
        if rc != 0:
            raise CaldError(self._L.cald_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.cald_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict):
        """state_dict: torchvision-keyed mapping name -> tensor / ndarray (checkpoint['model'])."""
        names, arrays = [], []
        for k, v in state_dict.items():
            if hasattr(v, "detach"):
                v = v.detach().cpu().numpy()
            a = np.ascontiguousarray(v, dtype=np.float32)
            if a.ndim > 4:
                raise ValueError("tensor %s has rank %d" % (k, a.ndim))
            names.append(arch.canonical_key(k).encode())
            arrays.append(a)
        n = len(names)
        c_names = (c_char_p * n)(*names)
        c_data = (POINTER(c_float) * n)(*[a.ctypes.data_as(POINTER(c_float)) for a in arrays])
        c_ndim = (c_int * n)(*[a.ndim for a in arrays])
        shp = np.ones((n, 4), dtype=np.int64)
        for i, a in enumerate(arrays):
            shp[i, :a.ndim] = a.shape
        self._L.cald_load_weights.argtypes = [c_void_p, c_int, POINTER(c_char_p), POINTER(POINTER(c_float)),
                                              POINTER(c_int), POINTER(c_int64)]
        self._check(self._L.cald_load_weights(self._h, n, c_names, c_data, c_ndim,
                                              shp.ctypes.data_as(POINTER(c_int64))))

    # ------------------------------------------------------------------ scoring
    @staticmethod
    def _aug_array(views):
        views = [(v, 0.0) if isinstance(v, int) else v for v in views]
        return (Aug * max(1, len(views)))(*[Aug(int(k), float(p)) for k, p in views]), len(views)

    def score(self, images, views, bp=1.3, uniforms=None, noise=None, swap_perms=None):
        """views: list of (kind, param) (see expand_augs); noise: list of float32 [3,H,W] planes, one per
        (image, noise view) in image-major order; swap_perms: int per (image, color_swap view).
        -> (consistency float64[n], cls float64[n, C-1], consumed)."""
        imgs, ptrs, hs, ws = _u8_list(images)
        n = len(imgs)
        a, na = self._aug_array(views)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        nu = 0 if u is None else u.size
        nz_keep, nz_ptrs = [], None
        if noise:
            n_noise = sum(1 for v in views if (v if isinstance(v, int) else v[0]) in NOISE_KINDS)
            if len(noise) != n * n_noise:
                raise ValueError("expected %d noise planes (%d images x %d noise views), got %d"
                                 % (n * n_noise, n, n_noise, len(noise)))
            nz_keep = [np.ascontiguousarray(z, dtype=np.float32) for z in noise]
            nz_ptrs = (POINTER(c_float) * len(nz_keep))(*[z.ctypes.data_as(POINTER(c_float)) for z in nz_keep])
        sp = None if swap_perms is None else np.ascontiguousarray(swap_perms, dtype=np.int32)
        consumed = c_int(0)
        cons = np.zeros(n, dtype=np.float64)
        cls = np.zeros((n, self.num_classes - 1), dtype=np.float64)
        self._L.cald_score.argtypes = [c_void_p, c_int, POINTER(POINTER(c_uint8)), POINTER(c_int), POINTER(c_int),
                                       c_int, POINTER(Aug), c_double, POINTER(c_double), c_int, POINTER(c_int),
                                       POINTER(POINTER(c_float)), POINTER(c_int), POINTER(c_double),
                                       POINTER(c_double)]
        self._check(self._L.cald_score(self._h, n, ptrs, hs, ws, na, a, float(bp),
                                       None if u is None else u.ctypes.data_as(POINTER(c_double)), nu,
                                       ctypes.byref(consumed), nz_ptrs,
                                       None if sp is None else sp.ctypes.data_as(POINTER(c_int)),
                                       cons.ctypes.data_as(POINTER(c_double)),
                                       cls.ctypes.data_as(POINTER(c_double))))
        return cons, cls, consumed.value

    # ------------------------------------------------------------------ pool ingest (JPEG files)
    @staticmethod
    def _file_list(files):
        keep = [np.frombuffer(f, dtype=np.uint8) if isinstance(f, (bytes, bytearray, memoryview)) else
                np.ascontiguousarray(f, dtype=np.uint8) for f in files]
        n = len(keep)
        ptrs = (POINTER(c_uint8) * n)(*[k.ctypes.data_as(POINTER(c_uint8)) for k in keep])
        sizes = (c_size_t * n)(*[k.size for k in keep])
        return keep, ptrs, sizes

    def jpeg_info(self, data):
        """(height, width, components) of one JPEG file (bytes); host-only header walk."""
        keep, ptrs, sizes = self._file_list([data])
        h, w, c = c_int(0), c_int(0), c_int(0)
        self._L.cald_jpeg_info.argtypes = [POINTER(c_uint8), c_size_t, POINTER(c_int), POINTER(c_int), POINTER(c_int)]
        if self._L.cald_jpeg_info(ptrs[0], sizes[0], ctypes.byref(h), ctypes.byref(w), ctypes.byref(c)) != 0:
            raise CaldError(self._L.cald_last_error(None).decode())
        return h.value, w.value, c.value

    def decode_jpeg(self, files):
        """Decode baseline JPEG files (bytes objects) ON THE DEVICE -> list of HxWx3 u8 arrays, bit-identical to
        np.asarray(PIL.Image.open(f).convert('RGB'))."""
        keep, ptrs, sizes = self._file_list(files)
        n = len(keep)
        outs = []
        for f in files:
            h, w, _ = self.jpeg_info(f)
            outs.append(np.zeros((h, w, 3), dtype=np.uint8))
        optrs = (POINTER(c_uint8) * n)(*[o.ctypes.data_as(POINTER(c_uint8)) for o in outs])
        self._L.cald_jpeg_decode.argtypes = [c_void_p, c_int, POINTER(POINTER(c_uint8)), POINTER(c_size_t),
                                             POINTER(POINTER(c_uint8))]
        self._check(self._L.cald_jpeg_decode(self._h, n, ptrs, sizes, optrs))
        return outs

    def score_jpeg(self, files, views, bp=1.3, uniforms=None, swap_perms=None):
        """score() over JPEG FILES (bytes objects): decode and scoring both on the device.
        -> (consistency float64[n], cls float64[n, C-1], consumed, heights, widths)."""
        keep, ptrs, sizes = self._file_list(files)
        n = len(keep)
        a, na = self._aug_array(views)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        sp = None if swap_perms is None else np.ascontiguousarray(swap_perms, dtype=np.int32)
        consumed = c_int(0)
        cons = np.zeros(n, dtype=np.float64)
        cls = np.zeros((n, self.num_classes - 1), dtype=np.float64)
        hs, ws = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        self._L.cald_score_jpeg.argtypes = [c_void_p, c_int, POINTER(POINTER(c_uint8)), POINTER(c_size_t), c_int,
                                            POINTER(Aug), c_double, POINTER(c_double), c_int, POINTER(c_int),
                                            POINTER(c_int), POINTER(c_double), POINTER(c_double), POINTER(c_int),
                                            POINTER(c_int)]
        self._check(self._L.cald_score_jpeg(self._h, n, ptrs, sizes, na, a, float(bp),
                                            None if u is None else u.ctypes.data_as(POINTER(c_double)),
                                            0 if u is None else u.size, ctypes.byref(consumed),
                                            None if sp is None else sp.ctypes.data_as(POINTER(c_int)),
                                            cons.ctypes.data_as(POINTER(c_double)), cls.ctypes.data_as(POINTER(c_double)),
                                            hs.ctypes.data_as(POINTER(c_int)), ws.ctypes.data_as(POINTER(c_int))))
        return cons, cls, consumed.value, hs, ws

    def select(self, uncertainty, cls_corrs, mean_hist, budget, n_cand, uniform=False):
        """On-device argsort -> candidates -> cls_kldiv (cald_train.py:234-271, 439-448) -> pool positions to label."""
        u = np.ascontiguousarray(uncertainty, dtype=np.float64)
        c = np.ascontiguousarray(cls_corrs, dtype=np.float64)
        h = np.ascontiguousarray(mean_hist, dtype=np.float64)
        n, c1 = c.shape
        out = np.zeros(max(1, min(n, n_cand)), dtype=np.int32)
        k = c_int(0)
        dp = POINTER(c_double)
        self._L.cald_select.argtypes = [c_void_p, c_int, dp, dp, c_int, dp, c_int, c_int, c_int, POINTER(c_int), c_int,
                                        POINTER(c_int)]
        self._check(self._L.cald_select(self._h, n, u.ctypes.data_as(dp), c.ctypes.data_as(dp), c1, h.ctypes.data_as(dp),
                                        int(budget), int(n_cand), int(bool(uniform)), out.ctypes.data_as(POINTER(c_int)),
                                        out.size, ctypes.byref(k)))
        return out[:k.value]

    def score_device(self, d_ptrs, hs, ws, views, bp=1.3, uniforms=None):
        """Like score() but images are already resident in HBM (list of device pointers)."""
        n = len(d_ptrs)
        ptrs = (POINTER(c_uint8) * n)(*[ctypes.cast(int(p), POINTER(c_uint8)) for p in d_ptrs])
        chs = (c_int * n)(*hs)
        cws = (c_int * n)(*ws)
        a, na = self._aug_array(views)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        consumed = c_int(0)
        cons = np.zeros(n, dtype=np.float64)
        cls = np.zeros((n, self.num_classes - 1), dtype=np.float64)
        self._L.cald_score_device.argtypes = [c_void_p, c_int, POINTER(POINTER(c_uint8)), POINTER(c_int),
                                              POINTER(c_int), c_int, POINTER(Aug), c_double, POINTER(c_double),
                                              c_int, POINTER(c_int), POINTER(POINTER(c_float)), POINTER(c_int),
                                              POINTER(c_double), POINTER(c_double)]
        self._check(self._L.cald_score_device(self._h, n, ptrs, chs, cws, na, a, float(bp),
                                              None if u is None else u.ctypes.data_as(POINTER(c_double)),
                                              0 if u is None else u.size, ctypes.byref(consumed), None, None,
                                              cons.ctypes.data_as(POINTER(c_double)),
                                              cls.ctypes.data_as(POINTER(c_double))))
        return cons, cls, consumed.value

    def score_lsc(self, images, noise):
        """LS+C stability (ls_c_train.py:108-155).  noise: 6 float32 [3,H,W] torch.randn planes per image,
        image-major (std 8, 16, ..., 48 is applied by the engine).  -> float64[n]."""
        imgs, ptrs, hs, ws = _u8_list(images)
        n = len(imgs)
        nz_keep = [np.ascontiguousarray(z, dtype=np.float32) for z in noise]
        if len(nz_keep) != 6 * n:
            raise ValueError("LS+C needs 6 noise planes per image")
        nz_ptrs = (POINTER(c_float) * len(nz_keep))(*[z.ctypes.data_as(POINTER(c_float)) for z in nz_keep])
        out = np.zeros(n, dtype=np.float64)
        self._L.cald_score_lsc.argtypes = [c_void_p, c_int, POINTER(POINTER(c_uint8)), POINTER(c_int), POINTER(c_int),
                                           POINTER(POINTER(c_float)), POINTER(c_double)]
        self._check(self._L.cald_score_lsc(self._h, n, ptrs, hs, ws, nz_ptrs, out.ctypes.data_as(POINTER(c_double))))
        return out

    def score_ltc(self, images):
        """LT/C uncertainty (lt_c_train.py:105-121), Faster R-CNN only.  -> float64[n]."""
        imgs, ptrs, hs, ws = _u8_list(images)
        n = len(imgs)
        out = np.zeros(n, dtype=np.float64)
        self._L.cald_score_ltc.argtypes = [c_void_p, c_int, POINTER(POINTER(c_uint8)), POINTER(c_int), POINTER(c_int),
                                           POINTER(c_double)]
        self._check(self._L.cald_score_ltc(self._h, n, ptrs, hs, ws, out.ctypes.data_as(POINTER(c_double))))
        return out

    def last_ref_counts(self, n_images):
        """Detections of each image's reference view in the last score() / score_lsc() call (0 = the reference skips
        the image before drawing any augmentation randomness)."""
        out = np.zeros(n_images, dtype=np.int32)
        self._L.cald_last_ref_counts.argtypes = [c_void_p, POINTER(c_int), c_int]
        k = self._L.cald_last_ref_counts(self._h, out.ctypes.data_as(POINTER(c_int)), n_images)
        return out[:max(k, 0)]

    def last_per_view(self, n_images, n_augs):
        out = np.zeros(n_images * n_augs, dtype=np.float32)
        self._L.cald_last_per_view.argtypes = [c_void_p, POINTER(c_float), c_int]
        k = self._L.cald_last_per_view(self._h, out.ctypes.data_as(POINTER(c_float)), out.size)
        return out[:k].reshape(-1, n_augs) if n_augs else out[:0]

    def detect(self, images):
        """List of per-image dicts with the reference's output schema (frcnn_la.py:131-141; RetinaNet:
        retinanet_cal.py:479-485 -- labels are 0-based there and 'props' is all zero)."""
        imgs, ptrs, hs, ws = _u8_list(images)
        n = len(imgs)
        # per-image capacity of the detection list: FRCNN box_detections_per_img, RetinaNet retina_max_detections
        cap = self.cfg.retina_max_detections if self.arch_id == ARCH_RETINANET else self.cfg.box_detections_per_img
        C = self.num_classes
        counts = np.zeros(n, dtype=np.int32)
        boxes = np.zeros((n, cap, 4), dtype=np.float32)
        props = np.zeros((n, cap, 4), dtype=np.float32)
        scores = np.zeros((n, cap), dtype=np.float32)
        pmax = np.zeros((n, cap), dtype=np.float32)
        labels = np.zeros((n, cap), dtype=np.int64)
        scls = np.zeros((n, cap, C), dtype=np.float32)
        fp = POINTER(c_float)
        self._L.cald_detect.argtypes = [c_void_p, c_int, POINTER(POINTER(c_uint8)), POINTER(c_int), POINTER(c_int),
                                        POINTER(c_int), fp, fp, POINTER(c_int64), fp, fp, fp]
        self._check(self._L.cald_detect(self._h, n, ptrs, hs, ws, counts.ctypes.data_as(POINTER(c_int)),
                                        boxes.ctypes.data_as(fp), scores.ctypes.data_as(fp),
                                        labels.ctypes.data_as(POINTER(c_int64)), props.ctypes.data_as(fp),
                                        pmax.ctypes.data_as(fp), scls.ctypes.data_as(fp)))
        out = []
        for i in range(n):
            k = int(counts[i])
            out.append({"boxes": boxes[i, :k], "labels": labels[i, :k], "scores": scores[i, :k],
                        "props": props[i, :k], "prob_max": pmax[i, :k], "scores_cls": scls[i, :k]})
        return out

    def debug_views(self, n_images, n_augs):
        """debug=True: detections of every view of the last score() call -> list (per image) of lists (reference
        view first, then the augmented views) of dicts boxes / scores / labels / prob_max."""
        nv = n_images * (1 + n_augs)
        self._L.cald_debug_views.restype = c_longlong
        self._L.cald_debug_views.argtypes = [c_void_p, POINTER(c_int), c_int, POINTER(c_float), POINTER(c_float),
                                             POINTER(c_int), POINTER(c_float), c_longlong]
        rows = self._L.cald_debug_views(self._h, None, 0, None, None, None, None, 0)
        counts = np.zeros(nv, dtype=np.int32)
        boxes = np.zeros((rows, 4), dtype=np.float32)
        scores = np.zeros(rows, dtype=np.float32)
        pmax = np.zeros(rows, dtype=np.float32)
        labels = np.zeros(rows, dtype=np.int32)
        fp = POINTER(c_float)
        self._L.cald_debug_views(self._h, counts.ctypes.data_as(POINTER(c_int)), nv, boxes.ctypes.data_as(fp),
                                 scores.ctypes.data_as(fp), labels.ctypes.data_as(POINTER(c_int)),
                                 pmax.ctypes.data_as(fp), rows)
        out, pos, k = [], 0, 0
        for _ in range(n_images):
            views = []
            for _ in range(1 + n_augs):
                n = int(counts[k]) if k < len(counts) else 0
                views.append({"boxes": boxes[pos:pos + n], "scores": scores[pos:pos + n],
                              "labels": labels[pos:pos + n], "prob_max": pmax[pos:pos + n]})
                pos += n
                k += 1
            out.append(views)
        return out

    def views_per_pass(self):
        return int(self._L.cald_views_per_pass(self._h))

    def images_per_chunk(self, n_augs=0):
        """Images the engine scores per chunk: as many as one pass holds views (a chunk's reference views run as one
        full pass, its augmented views as n_augs more)."""
        return max(1, self.views_per_pass())

    def arena_peak(self):
        self._L.cald_arena_peak.restype = c_longlong
        self._L.cald_arena_peak.argtypes = [c_void_p]
        return int(self._L.cald_arena_peak(self._h))

    def debug_fetch(self, name):
        self._L.cald_debug_fetch.restype = c_longlong
        self._L.cald_debug_fetch.argtypes = [c_void_p, c_char_p, POINTER(c_float), c_longlong]
        n = self._L.cald_debug_fetch(self._h, name.encode(), None, 0)
        if n < 0:
            raise CaldError(self._L.cald_last_error(self._h).decode())
        buf = np.zeros(n, dtype=np.float32)
        self._L.cald_debug_fetch(self._h, name.encode(), buf.ctypes.data_as(POINTER(c_float)), n)
        return buf

    def counters(self):
        k = c_longlong(0)
        f = c_double(0)
        self._L.cald_counters.argtypes = [c_void_p, POINTER(c_longlong), POINTER(c_double)]
        self._L.cald_counters(self._h, ctypes.byref(k), ctypes.byref(f))
        return k.value, f.value

    # ------------------------------------------------------------------ measurement hooks
    def profile(self, enable=True):
        self._L.cald_profile.argtypes = [c_void_p, c_int]
        self._L.cald_profile(self._h, int(enable))

    def profile_read(self):
        """-> (conv kernel ms, conv launches, conv algorithmic FLOPs) since the last read."""
        ms, n, fl = c_double(0), c_longlong(0), c_double(0)
        self._L.cald_profile_read.argtypes = [c_void_p, POINTER(c_double), POINTER(c_longlong), POINTER(c_double)]
        self._check(self._L.cald_profile_read(self._h, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl)))
        return ms.value, n.value, fl.value

    def profile_layers(self):
        """Per-layer rows of the last profile_read(): list of (signature, count, ms, gflop, mbyte)."""
        self._L.cald_profile_layers.restype = c_longlong
        self._L.cald_profile_layers.argtypes = [c_void_p, c_char_p, c_longlong]
        n = self._L.cald_profile_layers(self._h, None, 0)
        buf = ctypes.create_string_buffer(int(n))
        self._L.cald_profile_layers(self._h, buf, n)
        rows = []
        for line in buf.value.decode().splitlines()[1:]:
            sig, cnt, ms, gf, mb = line.split("\t")
            rows.append((sig, int(cnt), float(ms), float(gf), float(mb)))
        return rows

    def event_record(self, slot):
        self._L.cald_event_record.argtypes = [c_void_p, c_int]
        self._check(self._L.cald_event_record(self._h, slot))

    def event_elapsed_ms(self, a, b):
        ms = c_float(0)
        self._L.cald_event_elapsed_ms.argtypes = [c_void_p, c_int, c_int, POINTER(c_float)]
        self._check(self._L.cald_event_elapsed_ms(self._h, a, b, ctypes.byref(ms)))
        return ms.value
