"""CPU-side checks of the boundary: the shared library loads, exports every declared symbol, and the
host logic (weight key mapping, shard arithmetic, augmentation ordering) behaves.  No GPU compute."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    so = os.path.join(ROOT, "cald_b200", "libcald_b200.so")
    if not os.path.exists(so):
        subprocess.check_call(["bash", os.path.join(ROOT, "build.sh")])
    return so


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cald_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    names = _declared("cald_b200.h") + _declared("cald_b200_ops.h")
    assert "cald_score" in names and "cald_create" in names and "cald_op_conv2d" in names
    for n in names:
        assert hasattr(lib, n), n


def test_create_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cald_b200.engine import Engine
    from cald_b200._lib import CaldError
    with pytest.raises(CaldError):
        Engine(depth=50, num_classes=21)


def test_sass_contains_blackwell_tensor_and_tma_instructions(built):
    out = subprocess.run(["cuobjdump", "-sass", built], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out and "UTMALDG" in out and "LDTM" in out
    # CTA-pair conv kernel (igemm2.cuh): cta_group::2 MMA, multicast commit, pair TMA loads, cluster barrier
    for mnemonic in ("UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST", "UTMALDG.4D.2CTA", "UCGABAR_ARV"):
        assert mnemonic in out, mnemonic
    assert "HMMA.16816" not in out  # no legacy mma.sync tensor path anywhere in the library


def test_weight_key_aliases():
    from cald_b200 import arch
    assert arch.canonical_key("rpn.head.conv.weight") == "rpn.head.conv.0.0.weight"
    assert arch.canonical_key("backbone.fpn.inner_blocks.2.bias") == "backbone.fpn.inner_blocks.2.0.bias"
    assert arch.canonical_key("backbone.fpn.layer_blocks.0.0.weight") == "backbone.fpn.layer_blocks.0.0.weight"
    assert len(arch.frcnn_params(50, 21)) == 295 and len(arch.frcnn_params(101, 21)) == 550
    assert len(arch.retinanet_params(50, 21)) == 301


def test_planted_weights_are_deterministic_and_complete():
    from cald_b200 import arch, synth
    a = synth.planted_frcnn_weights(50, 21, 0)
    b = synth.planted_frcnn_weights(50, 21, 0)
    shapes = arch.frcnn_params(50, 21)
    assert list(a.keys()) == list(shapes.keys())
    for k in a:
        assert a[k].shape == tuple(shapes[k]) and a[k].dtype == np.float32 and np.array_equal(a[k], b[k])


def test_aug_order_follows_reference():
    from cald_b200 import api
    from cald_b200 import engine as E
    assert api._aug_kinds(['rotation', 'flip', 'cut_out', 'smaller_resize']) == \
        [(E.AUG_FLIP, 0.0), (E.AUG_CUTOUT, 2.0), (E.AUG_RESIZE, 0.8), (E.AUG_ROTATION, 5.0)]
    v = api._aug_kinds(['multi_resize', 'ga', 'multi_cut_out', 'sp'])
    assert v == [(E.AUG_GAUSS, 16.0), (E.AUG_SALTPEPPER, 0.1)] + [(E.AUG_CUTOUT, float(i)) for i in range(1, 5)] + \
        [(E.AUG_RESIZE, i * 0.1) for i in range(7, 10)]
    assert len(api._aug_kinds(['multi_ga', 'multi_sp'])) == 12
    # colour views sit between the Gaussian and the salt-pepper views (cald_train.py:136-149)
    assert api._aug_kinds(['sp', 'color_swap', 'ga', 'color_adjust']) == \
        [(E.AUG_GAUSS, 16.0), (E.AUG_COLOR_ADJUST, 1.5), (E.AUG_COLOR_SWAP, 0.0), (E.AUG_SALTPEPPER, 0.1)]
    with pytest.raises(NameError):  # cald_train.py:148 references an undefined name: unreachable upstream too
        api._aug_kinds(['multi_color_adjust'])


def test_shard_partition_covers_pool_in_order():
    from cald_b200 import shard
    for n in (0, 1, 7, 100, 10001):
        for world in (1, 2, 3, 8):
            parts = [shard.shard_indices(n, r, world) for r in range(world)]
            flat = np.concatenate(parts) if parts else np.zeros(0, int)
            assert sorted(flat.tolist()) == list(range(n))
            merged = shard.merge_shards([np.asarray(p, float) * 2 for p in parts], n, world)
            assert np.array_equal(merged, np.arange(n) * 2.0)


def test_bench_prints_exactly_one_json_line_on_stdout():
    """bench.py's contract is ONE JSON line on stdout; anything a library prints to file descriptor 1 during the run
    (NCCL's version banner under torchrun) must end up on stderr instead."""
    code = ("import os, sys, json; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "print('library noise'); os.write(1, b'raw fd-1 noise\\n'); bench.emit({'metric': 'm', 'value': 1.5})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"metric": "m", "value": 1.5}\n'
    assert "library noise" in r.stderr and "raw fd-1 noise" in r.stderr


def test_bench_reference_arm_emits_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the engine arm): one JSON line with the contract
    keys, the same `config` object the engine arm emits, kind "port"."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg4",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    want = bench.common_config(bench.CONFIGS["cfg4"], argparse.Namespace(mode="weak"))
    assert d["config"] == want            # identical in both arms (the driver's same_config check)
