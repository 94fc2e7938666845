#!/usr/bin/env python
"""bench.py -- unlabeled images scored / s.

Default workload = BASELINE.json configs[1] ("cfg2"): Faster R-CNN R50-FPN, nc = 91, synthetic 1333x800 (W x H) u8
pool, min/max size 800/1333, augmentations F, C, D, R, bp = 1.3, planted weights.  One "step" = one pass of the hot
path (1 reference + A augmented detector forwards + the paired-prediction reduction) over one batch of --batch images
per GPU; every step scores images it has not seen before.

    value          images/s with the u8 pool already resident in HBM (cald_score_device), NOT instrumented
    roofline       a second pass over the same steps with every tcgen05 conv/GEMM launch bracketed by CUDA events
    e2e            the public API (cald_b200.api.score_images -> cald_score) fed from page-locked HOST images, H2D and
                   D2H inside the timed region; e2e.pageable = the same from ordinary (pageable) numpy arrays
    cpu_baseline   the oracle port of cald_train.get_uncertainty on the box's host cores (N = 1 only)

All GPU legs are timed with CUDA events recorded on the engine's own stream; under torchrun the result is the max over
ranks.  Other BASELINE configs: --config cfg3 (RetinaNet R50-FPN), cfg4 (Faster R-CNN R101-FPN, VOC shape, A = 3),
cfg5 (--mode cycle: a fixed pool sharded over the ranks through cald_b200.shard.get_uncertainty_sharded, one
all-gather, then the host selection of cald_train.py:439-448 -- strong scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference] [--config cfgN] [--mode cycle]
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: model, depth, num_classes, (H, W), min/max size, augmentations, BASELINE.json configs[] index
    "cfg2": dict(model="frcnn", depth=50, nc=91, hw=(800, 1333), size=(800, 1333),
                 augs=['flip', 'cut_out', 'smaller_resize', 'rotation'], idx=1,
                 metric="unlabeled images scored/sec (FRCNN R50-FPN, 800x1333)",
                 workload="FRCNN R50-FPN nc=91, synthetic 1333x800 pool, 1 ref + 4 aug (F,C,D,R) forwards per image"),
    "cfg3": dict(model="retinanet", depth=50, nc=91, hw=(800, 1333), size=(800, 1333),
                 augs=['flip', 'cut_out', 'smaller_resize', 'rotation'], idx=2,
                 metric="unlabeled images scored/sec (RetinaNet R50-FPN, 800x1333)",
                 workload="RetinaNet R50-FPN nc=91, synthetic 1333x800 pool, 1 ref + 4 aug (F,C,D,R) forwards per image"),
    "cfg4": dict(model="frcnn", depth=101, nc=21, hw=(375, 500), size=(600, 1000),
                 augs=['flip', 'cut_out', 'smaller_resize'], idx=3,
                 metric="unlabeled images scored/sec (FRCNN R101-FPN, VOC shape 500x375, 3 augmentations)",
                 workload="FRCNN R101-FPN nc=21, synthetic 500x375 pool (600/1000), 1 ref + 3 aug (F,C,D) forwards per image"),
    "cfg5": dict(model="frcnn", depth=50, nc=91, hw=(800, 1333), size=(800, 1333),
                 augs=['flip', 'cut_out', 'smaller_resize', 'rotation'], idx=4,
                 metric="unlabeled images scored/sec, full AL cycle incl. selection (FRCNN R50-FPN, 800x1333)",
                 workload="FRCNN R50-FPN nc=91, synthetic 1333x800 pool sharded over the GPUs, score -> all-gather -> "
                          "argsort -> cls_kldiv -> select (cald_train.py:427-448)"),
}
# DRAM traffic of the conv kernels per scored image (dram__bytes_read.sum + dram__bytes_write.sum, ncu, summed over the 213
# igemm_tc_kernel / igemm_tc2_kernel / igemm_t_kernel launches of one 16-image cfg2 step: 218.63 GB / 16).  An OFFLINE ncu constant (ncu
# cannot run inside a timed bench), independent of the pass size to first order (every activation is written and read
# once per image); `roofline.traffic` = this x images per step / conv launches per step.  Only reported for cfg2.
NCU_DRAM_BYTES_PER_IMAGE = 13.664e9
NCU_DRAM_SOURCE = "profiles/r02_launches_step.csv"
BASE_IMAGES = 32
POOL_MAX_IMAGES = 1024   # distinct images per rank (3.3 GB at 1333x800); longer runs cycle through them


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1394.2), d.get("hbm_gbs", 6482.7), "MEASURED_PEAKS.json (sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class Pool:
    """Deterministic synthetic pool: image i = base image (i mod 32) rolled by an i-dependent offset with an
    i-dependent channel order -- distinct pixels for every index at a fraction of the cost of synthesising each."""

    def __init__(self, h, w, seed):
        from cald_b200 import synth
        self.base = [synth.synth_image(seed * 100000 + i, h, w, 0) for i in range(BASE_IMAGES)]
        self.perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]

    def __call__(self, i):
        b = self.base[i % BASE_IMAGES]
        k = i // BASE_IMAGES
        if k == 0:
            return b
        out = np.roll(b, ((37 * k) % b.shape[0], (101 * k) % b.shape[1]), axis=(0, 1))
        return np.ascontiguousarray(out[:, :, list(self.perms[k % 6])])


def planted(cfg):
    from cald_b200 import synth
    if cfg["model"] == "retinanet":
        return synth.planted_retinanet_weights(cfg["nc"], 0, cls_bias_shift=-11.0 if cfg["nc"] == 91 else -7.0)
    return synth.planted_frcnn_weights(cfg["depth"], cfg["nc"], 0)


def oracle_forward_fn(cfg):
    """CPU port (oracle/) of the reference path -- the checker, used here only as the timed CPU baseline."""
    import torch
    from oracle import frcnn_oracle as fo
    w = {k: torch.from_numpy(v) for k, v in planted(cfg).items()}
    if cfg["model"] == "retinanet":
        from oracle import retina_oracle as ro
        ocfg = ro.Cfg(cfg["depth"], cfg["nc"], cfg["size"][0], cfg["size"][1])
        return lambda x: ro.forward(x, w, ocfg)
    ocfg = fo.Cfg(cfg["depth"], cfg["nc"], cfg["size"][0], cfg["size"][1])
    return lambda x: fo.forward(x, w, ocfg)


def host_threads():
    """All host cores this process may use.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which
    would silently make the CPU arm single-threaded: set the count explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_port_images_per_s(cfg, n_images, warm, seed):
    import torch
    from oracle import cald_oracle as co
    from oracle import frcnn_oracle as fo
    fo.USE_TORCHVISION_OPS = True    # NMS / RoIAlign through torchvision's CPU kernels, like the reference's loop
    torch.set_num_threads(host_threads())
    fwd = oracle_forward_fn(cfg)
    pool = Pool(cfg["hw"][0], cfg["hw"][1], seed)
    random.seed(0)
    for i in range(warm):
        co.score_image(fwd, pool(i), cfg["augs"], cfg["nc"], 1.3)
    t = time.time()
    for i in range(warm, warm + n_images):
        co.score_image(fwd, pool(i), cfg["augs"], cfg["nc"], 1.3)
    dt = time.time() - t
    return n_images / dt, dt, torch.get_num_threads()


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL writes its version banner to stdout
    when NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for the whole run and the
    result line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(obj) + "\n").encode())


def common_config(cfg, args):
    """The `config` object, identical in both arms (the arms describe their own batching elsewhere)."""
    return {"workload": cfg["workload"], "baseline_config_index": cfg["idx"], "views_per_image": 1 + len(cfg["augs"]),
            "augmentations": cfg["augs"], "num_classes": cfg["nc"], "image_hw": list(cfg["hw"]),
            "min_max_size": list(cfg["size"]), "mode": args.mode,
            "l2": "every step scores other images than the step before (a pool of up to 1024 distinct images per GPU, "
                  "3.3 GB, is walked in order); one step's input images (>100 MB) and activations (>10 GB) exceed "
                  "the 126 MB L2 many times over"}


def run_reference(args, cfg, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the same ATen CPU kernels in the
    same order as cald_train.get_uncertainty, profiles/r02_cpu_port_vs_reference.md), all host threads."""
    if rank != 0:
        return
    v, dt, cores = cpu_port_images_per_s(cfg, args.steps, args.warmup, seed=11)
    sample = "%d images (1 per step, %d warm-up), oracle port of cald_train.get_uncertainty on torch CPU fp32" % (
        args.steps, args.warmup)
    emit({
        "impl": "reference", "metric": cfg["metric"], "value": v, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak" if args.mode == "weak" else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": common_config(cfg, args),
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="images per step per GPU (two engine chunks of 32 images)")
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="BASELINE.json configuration (default cfg2)")
    ap.add_argument("--model", default=None, choices=["frcnn", "retinanet"], help="shorthand: retinanet = --config cfg3")
    ap.add_argument("--mode", default=None, choices=["weak", "cycle"],
                    help="weak: every rank scores its own --batch images per step (default); cycle: one fixed pool of "
                         "--pool images sharded over the ranks + all-gather + host selection (default for cfg5)")
    ap.add_argument("--pool", type=int, default=0, help="cycle mode: pool size (default 128 images per GPU-step budget)")
    ap.add_argument("--budget", type=int, default=0, help="cycle mode: images to select (default pool / 8, cfg-5: 1000)")
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "f16"],
                    help="f16x3 = split-half operands, fp32-faithful (the product); f16 = single pass, not parity")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-images", type=int, default=5)
    ap.add_argument("--workspace-gb", type=float, default=0.0, help="device arena size (0 = half of free memory)")
    ap.add_argument("--views-per-pass", type=int, default=0, help="engine max_views_per_pass (0 = the engine's own auto)")
    ap.add_argument("--layers", default=None, help="write the per-layer conv timing table (TSV) to this path")
    ap.add_argument("--quick", action="store_true",
                    help="skip the instrumented (roofline) and pageable-host legs: value + e2e only (multi-GPU runs)")
    ap.add_argument("--only-value", action="store_true",
                    help="run the device-resident leg only (for short runs under ncu; prints no JSON line)")
    args = ap.parse_args()
    name = args.config or ("cfg3" if args.model == "retinanet" else "cfg2")
    cfg = CONFIGS[name]
    if args.mode is None:
        args.mode = "cycle" if name == "cfg5" else "weak"

    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from cald_b200 import api, shard
    from cald_b200.engine import Engine, PREC_F16, PREC_F16X3, ARCH_FRCNN, ARCH_RETINANET, expand_augs
    AUGS = cfg["augs"]
    H, W = cfg["hw"]
    kinds = expand_augs(AUGS)
    retina = cfg["model"] == "retinanet"
    # the engine is built the way the drop-in builds it (api.engine_for): views per pass come from the engine's own
    # sizing unless --views-per-pass overrides it
    eng = Engine(depth=cfg["depth"], num_classes=cfg["nc"], min_size=cfg["size"][0], max_size=cfg["size"][1],
                 device=local_rank, precision=PREC_F16 if args.precision == "f16" else PREC_F16X3,
                 max_views_per_pass=args.views_per_pass, arch_id=ARCH_RETINANET if retina else ARCH_FRCNN,
                 workspace_bytes=int(args.workspace_gb * (1 << 30)))
    eng.load_state_dict(planted(cfg))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    if args.mode == "cycle":
        run_cycle(args, cfg, eng, rank, local_rank, world, barrier, max_over_ranks)
        if world > 1:
            dist.destroy_process_group()
        return

    B = args.batch
    n_steps_total = args.warmup + args.steps
    # every rank scores its own shard of the pool: distinct images for every step (weak scaling).  Host copies in
    # page-locked memory (e2e leg) and in ordinary pageable memory (e2e.pageable); a device copy for the resident legs.
    make = Pool(H, W, seed=rank + 1)
    n_distinct = min(B * n_steps_total, max(B, POOL_MAX_IMAGES // B * B))   # bounded host / device memory
    pinned = [torch.from_numpy(make(i)).pin_memory() for i in range(n_distinct)]
    pool = [t.numpy() for t in pinned]
    dev_pool = [t.cuda() for t in pinned]

    def step_images(s):
        return [(s * B + j) % n_distinct for j in range(B)]

    def uniforms(s):
        return np.random.RandomState(1234 + s).random_sample(200 * B)

    def resident_pass(instrument):
        for s in range(args.warmup):
            idx = step_images(s)
            eng.score_device([dev_pool[i].data_ptr() for i in idx], [H] * B, [W] * B, kinds, 1.3, uniforms(s))
        barrier()
        eng.profile(instrument)
        k0, _ = eng.counters()
        eng.event_record(0)
        scores = []
        for s in range(args.warmup, n_steps_total):
            idx = step_images(s)
            c, v, _ = eng.score_device([dev_pool[i].data_ptr() for i in idx], [H] * B, [W] * B, kinds, 1.3, uniforms(s))
            scores.append(c)
        if world > 1:
            # the path's single collective: all-gather of the per-image scores
            t = torch.tensor(np.concatenate(scores), device="cuda")
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t)
        eng.event_record(1)
        barrier()
        ms = max_over_ranks(eng.event_elapsed_ms(0, 1))
        k1, _ = eng.counters()
        return ms, int(k1 - k0)

    # ---------------- leg 1: value (not instrumented)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches = resident_pass(False)
    clocks = sampler.stop()
    if args.only_value:
        print("only-value: %.2f images/s" % (world * B * args.steps / (ms / 1000.0)), file=sys.stderr)
        return
    value = world * B * args.steps / (ms / 1000.0)
    # ---------------- leg 2: the same steps with per-launch CUDA events on the conv kernels (roofline)
    if args.quick:
        ms_instr, conv_ms, conv_launches, conv_flops, conv_bytes = ms, 0.0, 0, 0.0, 0.0
    else:
        ms_instr, _ = resident_pass(True)
        conv_ms, conv_launches, conv_flops = eng.profile_read()
        conv_bytes = sum(r[4] for r in eng.profile_layers()) * 1e6  # algorithmic HBM bytes of the timed conv launches
    if args.layers and rank == 0 and not args.quick:
        rows = sorted(eng.profile_layers(), key=lambda r: -r[2])
        with open(args.layers, "w") as f:
            f.write("layer\tcount\tms_total\tus_per_launch\tTFLOPs_algorithmic\tGBs_algorithmic\tshare\n")
            for sig, cnt, lms, gf, mb in rows:
                f.write("%s\t%d\t%.3f\t%.1f\t%.1f\t%.0f\t%.3f\n" % (
                    sig, cnt, lms, 1000.0 * lms / cnt, gf / lms if lms else 0, mb / lms if lms else 0,
                    lms / conv_ms if conv_ms else 0))
    eng.profile(False)

    # ---------------- legs 3 + 4: end to end through the public API with host buffers (page-locked, then pageable)
    def e2e_pass(images):
        barrier()
        for s in range(min(2, args.warmup)):  # warm the host-buffer path (staging buffers, first touch)
            random.seed(1000 + s)
            api.score_images(eng, [images[i] for i in step_images(s)], AUGS, chunk=B)
        barrier()
        t_wall = time.time()
        eng.event_record(2)
        for s in range(args.warmup, n_steps_total):
            random.seed(s)
            api.score_images(eng, [images[i] for i in step_images(s)], AUGS, chunk=B)
        eng.event_record(3)
        barrier()
        t_wall = time.time() - t_wall
        # the call is synchronous: wall clock bounds it too
        return world * B * args.steps / (max_over_ranks(max(eng.event_elapsed_ms(2, 3), 1000.0 * t_wall)) / 1000.0)

    e2e = e2e_pass(pool)
    if args.quick:
        e2e_pageable = None
    else:
        pageable = [np.array(a, copy=True) for a in pool]
        e2e_pageable = e2e_pass(pageable)
        del pageable

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_hbm, peak_src = peaks()
    achieved = conv_flops / (conv_ms / 1000.0) / 1e12 if conv_ms > 0 else 0.0
    A = len(AUGS)
    out = {
        "metric": cfg["metric"], "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f16" if args.precision == "f16" else "f16x3 (split-half operands: 22-bit significand, fp32 accumulate)",
        "data": "synthetic", "config": common_config(cfg, args),
        "engine": {"images_per_step_per_gpu": B, "views_per_pass": eng.views_per_pass(), "precision": args.precision,
                   "arena_peak_gb": eng.arena_peak() / 2 ** 30,
                   "value_instrumented": world * B * args.steps / (ms_instr / 1000.0)},
        "gpu_launches": launches,
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": B * H * W * 3 + 200 * B * 8,
                "d2h_bytes_per_step": B * (A + (1 + A) * (cfg["nc"] - 1) + 1) * 4 + 4,
                "host_memory": "page-locked",
                "pageable": None if e2e_pageable is None else {"value": e2e_pageable, "unit": "images/s"}},
        "roofline": None if args.quick else {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf if peak_tf else None,
                     "frac_of_x3_ceiling": 3.0 * achieved / peak_tf if peak_tf else None,
                     "traffic": (NCU_DRAM_BYTES_PER_IMAGE * B * args.steps / conv_launches
                                 if (name == "cfg2" and conv_launches) else None),
                     "traffic_source": "offline ncu constant: conv-kernel DRAM bytes (read + write) per scored image "
                                       "from " + NCU_DRAM_SOURCE + ", scaled to this run's images and launches per step",
                     "algorithmic_bytes_per_launch": conv_bytes / conv_launches if conv_launches else None,
                     "kernel": "igemm_tc_kernel + igemm_tc2_kernel + igemm_t_kernel (tcgen05 implicit-GEMM conv/GEMM: "
                               "one-CTA, CTA-pair cta_group::2 and transposed-role instantiations)",
                     "launches": int(conv_launches), "kernel_ms_per_step": conv_ms / args.steps,
                     "algorithmic_gflop_per_step": conv_flops / args.steps / 1e9,
                     "share_of_step": conv_ms / ms_instr if ms_instr else None, "peak_source": peak_src,
                     "note": "achieved = algorithmic 2*MAC of the reference convs/GEMMs / summed CUDA-event kernel "
                             "time of the instrumented pass; the fp32-faithful split arithmetic issues 3 tensor-core "
                             "MACs per algorithmic MAC, so the kernel's own ceiling is peak/3"},
    }
    if not args.no_cpu_baseline and world == 1:  # the CPU baseline is timed at N = 1 only
        v, dt, cores = cpu_port_images_per_s(cfg, args.cpu_images, warm=1, seed=7)
        out["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                               "sample": "%d images of the same workload after 1 warm-up image (%.0f s), oracle port "
                                         "(torch CPU fp32)" % (args.cpu_images, dt)}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def run_cycle(args, cfg, eng, rank, local_rank, world, barrier, max_over_ranks):
    """One selection cycle over a FIXED pool (strong scaling), through the product's own multi-GPU path:
    shard.get_uncertainty_sharded (round-robin shard, chunked scoring from host images, one all-gather, un-permute)
    followed by the host selection api.select (argsort -> 1.2 x budget candidates -> cls_kldiv)."""
    import torch
    from cald_b200 import api, shard
    AUGS = cfg["augs"]
    H, W = cfg["hw"]
    P = args.pool or 2048
    budget = args.budget or (1000 if P >= 2400 else max(1, P // 8))
    make = Pool(H, W, seed=1)          # the same pool on every rank; a rank only touches its own shard
    # the loader stand-in: this rank's shard decoded into (pageable) host memory before the clock starts, like images
    # handed over by DataLoader workers; fetching one inside the timed region is a dictionary lookup
    mine = {int(i): make(int(i)) for i in shard.shard_indices(P, rank, world)}
    fetch = lambda i: mine[i]  # noqa: E731
    rs = np.random.RandomState(3)
    labeled = [((None,), ({"labels": torch.from_numpy(rs.randint(1, cfg["nc"], rs.randint(1, 8)))},)) for _ in range(500)]
    subset = list(range(10 ** 6, 10 ** 6 + P))
    random.seed(17 + rank)
    for s in range(args.warmup):       # warm-up: a few chunks outside the timed region
        api.score_images(eng, [make(P + s * args.batch + j) for j in range(args.batch)], AUGS, chunk=args.batch)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.time()
    eng.event_record(0)
    cons, cls = shard.get_uncertainty_sharded(eng, fetch, AUGS, rank, world, n=P, chunk=2 * args.batch)
    eng.event_record(1)
    t_score = time.time() - t0
    t1 = time.time()
    picked = api.select(cons, cls, subset, labeled, budget)
    t_select = time.time() - t1
    barrier()
    clocks = sampler.stop()
    # the cycle is host-synchronous: wall clock of score + gather + select, max over ranks
    ms = max_over_ranks(1000.0 * (t_score + t_select))
    ms_select = max_over_ranks(1000.0 * t_select)
    if rank != 0:
        return
    k, _ = eng.counters()
    emit({
        "metric": cfg["metric"], "value": P / (ms / 1000.0), "unit": "images/s", "n_gpus": world,
        "steps": (P + world * args.batch - 1) // (world * args.batch), "warmup": args.warmup,
        "ms_per_step": ms / max(1, (P + world * args.batch - 1) // (world * args.batch)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "f16x3 (split-half operands: 22-bit significand, fp32 accumulate)", "data": "synthetic",
        "config": common_config(cfg, args),
        "engine": {"images_per_chunk_per_gpu": args.batch, "views_per_pass": eng.views_per_pass()},
        "cycle": {"pool": P, "budget": budget, "selected": len(picked), "score_gather_ms": 1000.0 * t_score,
                  "select_host_ms": ms_select, "select_share": ms_select / ms,
                  "note": "the rank's shard sits decoded in pageable host memory (loader stand-in); the timed region "
                          "uploads, scores, all-gathers and selects; select = argsort + cls_kldiv on every rank's "
                          "copy of the gathered rows"},
        "gpu_launches": int(k), "clocks": clocks,
        "e2e": {"value": P / (ms / 1000.0), "unit": "images/s", "h2d_bytes_per_step": args.batch * H * W * 3,
                "d2h_bytes_per_step": args.batch * (len(AUGS) + (1 + len(AUGS)) * (cfg["nc"] - 1) + 1) * 4},
    })


if __name__ == "__main__":
    main()
