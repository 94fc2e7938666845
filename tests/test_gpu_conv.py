"""Stage parity: implicit-GEMM conv / linear kernels vs torch CPU fp32 conv2d.

Oracle for this stage = torch.nn.functional.conv2d on CPU in fp32 (the exact op the
reference's CPU run executes: tv resnet.py / feature_pyramid_network.py / rpn.py).
Tolerances: the split-half x3 mode carries 22 significand bits per operand and the truncating accumulate is
pre-compensated: error at the fp32 level (~1e-6 of the output scale); single-pass half ~1e-3.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x, w, b, stride, relu, res=None, res_mode=0):
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    y = F.conv2d(xt, torch.from_numpy(w), None if b is None else torch.from_numpy(b), stride=stride,
                 padding=w.shape[-1] // 2)
    if res is not None:
        r = torch.from_numpy(res).permute(0, 3, 1, 2)
        if res_mode == 2:
            r = F.interpolate(r, size=y.shape[-2:], mode="nearest")
        y = y + r
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def _case(seed, n, h, w, cin, cout, k, stride=1, relu=True, bias=True, res_mode=0, res_hw=None):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    wt = (rs.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
    b = rs.standard_normal(cout).astype(np.float32) if bias else None
    ho, wo = ((h + 1) // 2, (w + 1) // 2) if stride == 2 else (h, w)
    res = None
    if res_mode == 1:
        res = rs.standard_normal((n, ho, wo, cout)).astype(np.float32)
    elif res_mode == 2:
        res = rs.standard_normal((n,) + tuple(res_hw) + (cout,)).astype(np.float32)
    return x, wt, b, res


CASES = [
    # n, h, w, cin, cout, k, stride, res_mode, res_hw
    (1, 16, 16, 64, 64, 1, 1, 0, None),
    (2, 19, 25, 64, 256, 1, 1, 1, None),
    (1, 38, 50, 256, 64, 1, 1, 0, None),
    (2, 20, 28, 64, 64, 3, 1, 0, None),
    (1, 19, 25, 128, 128, 3, 1, 0, None),
    (1, 25, 42, 256, 256, 3, 1, 0, None),
    (2, 19, 25, 128, 128, 3, 2, 0, None),
    (1, 20, 26, 256, 512, 1, 2, 0, None),
    (1, 38, 50, 512, 256, 1, 1, 2, (19, 25)),
    (2, 19, 25, 256, 1024, 1, 1, 1, None),      # layer3 expand conv: 4 conv + 2 residual k-blocks, 8 n-blocks
    (1, 13, 21, 256, 16, 1, 1, 0, None),
    (1, 10, 100, 1024, 112, 1, 1, 0, None),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("impl", [1, 0])
def test_conv_split_matches_fp32(case, impl):
    from cald_b200 import ops
    n, h, w, cin, cout, k, stride, res_mode, res_hw = case
    x, wt, b, res = _case(hash(case) % 1000, n, h, w, cin, cout, k, stride, True, True, res_mode, res_hw)
    want = _ref(x, wt, b, stride, True, res, res_mode)
    got = ops.conv2d(x, wt, b, stride=stride, relu=True, res=res, res_mode=res_mode, prec=0, impl=impl)
    scale = np.abs(want).max()
    err = np.abs(got - want).max()
    assert err <= 5e-6 * scale + 1e-6, (err, scale)


@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_conv_block_n_variants(block_n):
    from cald_b200 import ops
    from cald_b200._lib import CaldError
    x, wt, b, _ = _case(7, 1, 24, 40, 128, 256, 3)
    want = _ref(x, wt, b, 1, False)
    for prec, tol in ((0, 5e-6), (1, 2e-3)):
        if prec == 0 and block_n == 256:
            # the split-half format keeps the (scaled) cross terms in their own accumulator columns: BLOCK_N <= 128
            with pytest.raises(CaldError):
                ops.conv2d(x, wt, b, prec=prec, impl=0, block_n=block_n)
            continue
        got = ops.conv2d(x, wt, b, prec=prec, impl=0, block_n=block_n)
        assert np.abs(got - want).max() <= tol * np.abs(want).max()


def test_conv_stride2_odd_sizes():
    """stride-2 convs read the full-resolution tensor through an element-strided TMA map: odd extents, where the last
    tap row / column falls into the zero padding, against torch fp32"""
    from cald_b200 import ops
    for (n, h, w, cin, cout, k) in [(2, 19, 25, 64, 128, 3), (1, 25, 42, 128, 64, 3), (1, 33, 17, 64, 256, 1),
                                    (1, 50, 84, 256, 256, 3)]:
        x, wt, b, _ = _case(n + h + w, n, h, w, cin, cout, k, stride=2)
        want = _ref(x, wt, b, 2, True)
        for impl in (0, 1):
            got = ops.conv2d(x, wt, b, stride=2, relu=True, prec=0, impl=impl)
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= 5e-6 * np.abs(want).max() + 1e-6, (n, h, w, cin, cout, k, impl)


def test_conv_large_k_chunked_accumulation():
    """fc6-like contraction (K = 12544).  The tensor core truncates when adding into the TMEM accumulator; the
    chunked-accumulation epilogue bounds that bias.  Unchunked error is reported for the record."""
    from cald_b200 import ops
    rs = np.random.RandomState(3)
    x = rs.standard_normal((1, 1, 300, 12544)).astype(np.float32)
    wt = (rs.standard_normal((128, 12544, 1, 1)) * 0.01).astype(np.float32)
    want = _ref(x, wt, None, 1, False)
    scale = np.abs(want).max()
    errs = {}
    for kc in (0, 16, 8, 4, 2, 1):
        got = ops.conv2d(x, wt, None, prec=0, impl=0, kc=kc)
        errs[kc] = float(np.abs(got - want).max() / scale)
    simt = ops.conv2d(x, wt, None, prec=0, impl=1)
    errs["simt"] = float(np.abs(simt - want).max() / scale)
    print("relative error by chunk length (k-blocks of 64):", errs)
    assert errs[4] <= 1e-5, errs
    got = ops.conv2d(x, wt, None, prec=0, impl=0)  # engine default
    assert np.abs(got - want).max() <= 1e-5 * scale



PAIR_CASES = [
    # n, h, w, cin, cout, k, stride, res_mode, res_hw      (all map to BLOCK_N = 128, TMA-store, unchunked launches)
    (1, 25, 42, 256, 256, 3, 1, 0, None),
    (3, 19, 25, 128, 128, 3, 1, 0, None),       # odd tile count: the peer of the last pair idles
    (2, 19, 25, 128, 128, 3, 2, 0, None),
    (1, 20, 26, 256, 512, 1, 2, 0, None),
    (1, 38, 50, 512, 256, 1, 1, 2, (19, 25)),   # FPN lateral: nearest-upsampled top-down add in the epilogue
    (1, 1, 300, 1024, 1024, 1, 1, 0, None),     # fc7-like linear mode, 3 row tiles
    (2, 100, 168, 256, 256, 3, 1, 0, None),     # 526 tiles: every cluster walks several pairs, all ring phases
    (1, 7, 9, 64, 128, 1, 1, 0, None),          # one tile, one k-block
    (2, 25, 42, 512, 512, 3, 1, 0, None),       # 72 k-blocks: chunked accumulation across the pair
    (1, 1, 200, 12544, 128, 1, 1, 0, None),     # fc6-like, 196 k-blocks, chunked
    (1, 25, 42, 256, 819, 3, 1, 0, None),       # RetinaNet cls_logits: 7 n-blocks, the last one ragged, direct fp32 stores
    (1, 10, 100, 1024, 112, 1, 1, 0, None),     # one ragged n-block (predictor-like)
    (2, 20, 28, 64, 64, 3, 1, 0, None),         # BLOCK_N = 64 pair instantiation (layer1 3x3), 9 k-blocks, 4 smem stages
    (3, 25, 42, 128, 64, 3, 2, 0, None),        # BLOCK_N = 64, stride 2, odd tile count
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_cta_pair_kernel_matches_fp32_and_single_cta(case, monkeypatch):
    """The cta_group::2 kernel (igemm2.cuh) against torch fp32 and against the one-CTA kernel on the same operands."""
    import ctypes
    from cald_b200 import ops
    from cald_b200._lib import lib
    L = lib()
    L.cald_ops_pair_launches.restype = ctypes.c_longlong
    n, h, w, cin, cout, k, stride, res_mode, res_hw = case
    x, wt, b, res = _case(hash(case) % 1000, n, h, w, cin, cout, k, stride, True, True, res_mode, res_hw)
    want = _ref(x, wt, b, stride, True, res, res_mode)
    monkeypatch.setenv("CALD_TFORM", "0")   # the 64-channel 3x3 cases would take the transposed-role kernel otherwise
    monkeypatch.setenv("CALD_CTA2", "0")
    before = L.cald_ops_pair_launches()
    single = ops.conv2d(x, wt, b, stride=stride, relu=True, res=res, res_mode=res_mode, prec=0, impl=0)
    assert L.cald_ops_pair_launches() == before
    monkeypatch.setenv("CALD_CTA2", "1")
    monkeypatch.setenv("CALD_CTA2_MIN_KB", "1")   # the engine default only pairs launches of >= 16 k-blocks
    monkeypatch.setenv("CALD_CTA2_MIN_KB64", "1")
    got = ops.conv2d(x, wt, b, stride=stride, relu=True, res=res, res_mode=res_mode, prec=0, impl=0)
    assert L.cald_ops_pair_launches() == before + 1, "the launch did not take the pair kernel"
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 5e-6 * scale + 1e-6
    # same split operands, same cross-term-separated accumulation: the two kernels agree far below the fp32 tolerance
    assert np.abs(got - single).max() <= 2e-6 * scale + 1e-7


TFORM_CASES = [
    # n, h, w, cin, cout, k: spatial 64-output-channel convs (layer1 3x3) -> igemm_t.cuh, 16 x 16 pixel patches
    (1, 16, 16, 64, 64, 3),        # exactly one patch
    (2, 20, 28, 64, 64, 3),        # ragged right and bottom edges
    (1, 7, 9, 64, 64, 3),          # one partial patch: 4 of 8 slabs, the last one half a patch row pair
    (3, 33, 47, 64, 64, 3),        # odd extents: a one-row slab at the bottom
    (2, 200, 336, 64, 64, 3),      # the real layer1 shape: 546 patches, every CTA walks several, all ring phases
    (1, 40, 40, 128, 64, 3),       # two k-blocks per tap
]


@pytest.mark.parametrize("case", TFORM_CASES)
def test_conv_transposed_role_kernel_matches_fp32_and_pixel_major_kernels(case, monkeypatch):
    """igemm_t.cuh (channels on M with the hi / lo weight planes stacked, 256 pixels on N) against torch fp32 and
    against the pixel-major kernel on the same operands: same products and accumulate counts, main sums in the same
    order, the two cross-term partial sums of a k-block in the other one (measured: <= 1e-7 of the output scale)."""
    import ctypes
    from cald_b200 import ops
    from cald_b200._lib import lib
    L = lib()
    L.cald_ops_tform_launches.restype = ctypes.c_longlong
    n, h, w, cin, cout, k = case
    x, wt, b, _ = _case(sum(case), n, h, w, cin, cout, k)
    want = _ref(x, wt, b, 1, True)
    monkeypatch.setenv("CALD_TFORM", "0")
    before = L.cald_ops_tform_launches()
    old = ops.conv2d(x, wt, b, relu=True, prec=0, impl=0)
    assert L.cald_ops_tform_launches() == before
    monkeypatch.setenv("CALD_TFORM", "1")
    got = ops.conv2d(x, wt, b, relu=True, prec=0, impl=0)
    assert L.cald_ops_tform_launches() == before + 1, "the launch did not take the transposed-role kernel"
    scale = np.abs(want).max()
    print("tform %s: max |got - fp32| %.2e, max |got - pixel-major kernel| %.2e (scale %.2f)" % (
        case, np.abs(got - want).max(), np.abs(got - old).max(), scale))
    assert np.abs(got - want).max() <= 5e-6 * scale + 1e-6
    assert np.abs(got - old).max() <= 3e-7 * scale
    # no bias / no ReLU path
    got2 = ops.conv2d(x, wt, None, relu=False, prec=0, impl=0)
    want2 = _ref(x, wt, None, 1, False)
    assert np.abs(got2 - want2).max() <= 5e-6 * np.abs(want2).max() + 1e-6


DUAL_CASES = [
    # n, h, w, cin (conv3 input), cout, h2, w2, cin2 (block input), stride2
    (2, 19, 25, 64, 256, 19, 25, 64, 1),        # layer1.0: linear mode, 1 + 1 k-blocks
    (1, 19, 25, 128, 512, 38, 50, 256, 2),      # layer2.0: stride-2 shortcut read through an element-strided map
    (1, 13, 21, 256, 1024, 25, 42, 512, 2),     # layer3.0, odd source extents
    (1, 20, 26, 128, 512, 39, 51, 256, 2),
]


@pytest.mark.parametrize("case", DUAL_CASES)
def test_conv_fused_projection_shortcut(case):
    """relu(conv3(y) + b3 + downsample(x) + bd) as ONE launch against the two torch fp32 convs added."""
    from cald_b200 import ops
    n, h, w, cin, cout, h2, w2, cin2, s2 = case
    rs = np.random.RandomState(sum(case))
    y = rs.standard_normal((n, h, w, cin)).astype(np.float32)
    x = rs.standard_normal((n, h2, w2, cin2)).astype(np.float32)
    w3 = (rs.standard_normal((cout, cin, 1, 1)) * np.sqrt(2.0 / cin)).astype(np.float32)
    wd = (rs.standard_normal((cout, cin2, 1, 1)) * np.sqrt(2.0 / cin2)).astype(np.float32)
    b3 = rs.standard_normal(cout).astype(np.float32)
    bd = rs.standard_normal(cout).astype(np.float32)
    a = F.conv2d(torch.from_numpy(y).permute(0, 3, 1, 2), torch.from_numpy(w3), torch.from_numpy(b3))
    b = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(wd), torch.from_numpy(bd), stride=s2)
    want = F.relu(a + b).permute(0, 2, 3, 1).contiguous().numpy()
    got = ops.conv2d_dual(y, w3, b3, x, wd, bd, stride2=s2, relu=True)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 5e-6 * np.abs(want).max() + 1e-6
