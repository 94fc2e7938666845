#!/usr/bin/env python
"""bench.py -- unlabeled images scored / s (Faster R-CNN R50-FPN, 800x1333, A = 4 augmentations).

Workload (BASELINE.json configs[1]): synthetic 1333x800 (W x H) u8 pool, nc = 91, min/max size
800/1333, augmentations F, C, D, R, bp = 1.3, planted weights.  One "step" = one pass of the hot
path (1 reference + 4 augmented detector forwards + the paired-prediction reduction) over one
batch of --batch images.  `value` times the step with the u8 pool already resident in HBM
(cald_score_device); `e2e` times the public API (cald_b200.api.score_images -> cald_score) with
HOST images, H2D and D2H inside the timed region.  Both are timed with CUDA events recorded on
the engine's own stream; under torchrun the result is the max over ranks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
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

H, W = 800, 1333
NUM_CLASSES = 91
MIN_SIZE, MAX_SIZE = 800, 1333
AUGS = ['flip', 'cut_out', 'smaller_resize', 'rotation']
METRIC = "unlabeled images scored/sec (FRCNN R50-FPN, 800x1333)"
WORKLOAD = "FRCNN R50-FPN nc=91, synthetic 1333x800 pool, 1 ref + 4 aug (F,C,D,R) forwards per image"
# --model retinanet = BASELINE.json configs[2] (RetinaNet R50-FPN, retinanet_cal.py), same pool shape and augmentations
# DRAM bytes per conv launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu) averaged over the 142 igemm_tc_kernel /
# igemm_tc2_kernel launches of one default step (batch 16: a 16-view reference pass + a 64-view augmented pass); source:
# profiles/r01_igemm_dram_step.csv, captured with tools/final_measure.sh.  Only valid for the default FRCNN workload.
NCU_DRAM_BYTES_PER_CONV_LAUNCH = 1556.2e6
NCU_DRAM_BATCH = 16
METRIC_RETINA = "unlabeled images scored/sec (RetinaNet R50-FPN, 800x1333)"
WORKLOAD_RETINA = "RetinaNet R50-FPN nc=91, synthetic 1333x800 pool, 1 ref + 4 aug (F,C,D,R) forwards per image"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1394.2), d.get("hbm_gbs", 6482.7), "measured"
    return 1400.0, 6650.0, "fallback"


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


def make_pool(n, seed=0):
    from cald_b200 import synth
    return [synth.synth_image(i, H, W, seed) for i in range(n)]


def planted(model):
    from cald_b200 import synth
    if model == "retinanet":
        return synth.planted_retinanet_weights(NUM_CLASSES, 0, cls_bias_shift=-11.0)
    return synth.planted_frcnn_weights(50, NUM_CLASSES, 0)


def oracle_forward_fn(model="frcnn"):
    """CPU port (oracle/) of the reference path -- the checker, used here only as the timed CPU baseline."""
    import torch
    from cald_b200 import synth
    from oracle import frcnn_oracle as fo
    if model == "retinanet":
        from oracle import retina_oracle as ro
        w = {k: torch.from_numpy(v) for k, v in planted(model).items()}
        cfg = ro.Cfg(50, NUM_CLASSES, MIN_SIZE, MAX_SIZE)
        return lambda x: ro.forward(x, w, cfg)
    w = {k: torch.from_numpy(v) for k, v in synth.planted_frcnn_weights(50, NUM_CLASSES, 0).items()}
    cfg = fo.Cfg(50, NUM_CLASSES, MIN_SIZE, MAX_SIZE)
    return lambda x: fo.forward(x, w, cfg)


def host_threads():
    """All host cores this process may use.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which
    would silently make the CPU arm single-threaded: set the count explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(n_images, threads=None, model="frcnn"):
    import torch
    from oracle import cald_oracle as co
    torch.set_num_threads(threads or host_threads())
    fwd = oracle_forward_fn(model)
    imgs = make_pool(n_images + 1, seed=7)
    random.seed(0)
    co.score_image(fwd, imgs[0][:200, :334].copy(), AUGS, NUM_CLASSES, 1.3)  # warm-up on a small crop
    t = time.time()
    for im in imgs[1:]:
        co.score_image(fwd, im, AUGS, NUM_CLASSES, 1.3)
    dt = time.time() - t
    return n_images / dt, torch.get_num_threads()


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


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads."""
    if rank != 0:
        return
    import torch
    from oracle import cald_oracle as co
    torch.set_num_threads(host_threads())
    fwd = oracle_forward_fn(args.model)
    imgs = make_pool(args.warmup + args.steps, seed=11)
    random.seed(0)
    for im in imgs[:args.warmup]:
        co.score_image(fwd, im, AUGS, NUM_CLASSES, 1.3)
    t = time.time()
    for im in imgs[args.warmup:]:
        co.score_image(fwd, im, AUGS, NUM_CLASSES, 1.3)
    dt = time.time() - t
    v = args.steps / dt
    cores = torch.get_num_threads()
    sample = "%d images (1 per step), oracle port of cald_train.get_uncertainty on torch CPU fp32" % args.steps
    emit(({
        "impl": "reference", "metric": METRIC_RETINA if args.model == "retinanet" else METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_RETINA if args.model == "retinanet" else WORKLOAD, "images_per_step": 1},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=16, help="images per step per GPU")
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--model", default="frcnn", choices=["frcnn", "retinanet"],
                    help="frcnn = BASELINE.json configs[1] (the headline); retinanet = configs[2]")
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-images", type=int, default=1)
    ap.add_argument("--workspace-gb", type=float, default=0.0, help="device arena size (0 = half of free memory)")
    ap.add_argument("--layers", default=None, help="write the per-layer conv timing table (TSV) to this path")
    args = ap.parse_args()

    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from cald_b200 import api, synth
    from cald_b200.engine import Engine, PREC_BF16, PREC_BF16X3, ARCH_FRCNN, ARCH_RETINANET, expand_augs
    kinds = expand_augs(AUGS)
    retina = args.model == "retinanet"
    eng = Engine(depth=50, num_classes=NUM_CLASSES, min_size=MIN_SIZE, max_size=MAX_SIZE, device=local_rank,
                 precision=PREC_BF16 if args.precision == "bf16" else PREC_BF16X3,
                 max_views_per_pass=args.batch * len(AUGS), arch_id=ARCH_RETINANET if retina else ARCH_FRCNN,
                 workspace_bytes=int(args.workspace_gb * (1 << 30)))
    eng.load_state_dict(planted(args.model))

    B = args.batch
    n_steps_total = args.warmup + args.steps
    # every rank scores its own shard of the pool (distinct images per step; weak scaling)
    # host pool in page-locked memory (the e2e leg copies from it every step); device copy for the resident leg
    pinned = [torch.from_numpy(synth.synth_image(rank * 100000 + i, H, W, 0)).pin_memory()
              for i in range(B * min(n_steps_total, 4))]
    pool = [t.numpy() for t in pinned]
    dev_pool = [t.cuda() for t in pinned]

    def step_images(s):
        idx = [(s * B + j) % len(pool) for j in range(B)]
        return idx

    def uniforms(s):
        rs = np.random.RandomState(1234 + s)
        return rs.random_sample(200 * B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident run: value + roofline
    for s in range(args.warmup):
        idx = step_images(s)
        eng.score_device([dev_pool[i].data_ptr() for i in idx], [H] * B, [W] * B, kinds, 1.3, uniforms(s))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.profile(True)
    k0, _ = eng.counters()
    eng.event_record(0)
    scores = []
    for s in range(args.warmup, n_steps_total):
        idx = step_images(s)
        c, v, _ = eng.score_device([dev_pool[i].data_ptr() for i in idx], [H] * B, [W] * B, kinds, 1.3, uniforms(s))
        scores.append(c)
    gathered = None
    if world > 1:
        # the path's single collective: all-gather of the per-image scores
        t = torch.tensor(np.concatenate(scores), device="cuda")
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        gathered = torch.cat(out)
    eng.event_record(1)
    barrier()
    ms = eng.event_elapsed_ms(0, 1)
    k1, _ = eng.counters()
    conv_ms, conv_launches, conv_flops = eng.profile_read()
    conv_bytes = sum(r[4] for r in eng.profile_layers()) * 1e6  # algorithmic HBM bytes of the timed conv launches
    if args.layers and rank == 0:
        rows = sorted(eng.profile_layers(), key=lambda r: -r[2])
        with open(args.layers, "w") as f:
            f.write("layer\tcount\tms_total\tus_per_launch\tTFLOPs_algorithmic\tGBs_algorithmic\tshare\n")
            for sig, cnt, lms, gf, mb in rows:
                f.write("%s\t%d\t%.3f\t%.1f\t%.1f\t%.0f\t%.3f\n" % (
                    sig, cnt, lms, 1000.0 * lms / cnt, gf / lms if lms else 0, mb / lms if lms else 0,
                    lms / conv_ms if conv_ms else 0))
    eng.profile(False)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms / 1000.0)

    # ---------------- end-to-end run through the public API with host buffers
    barrier()
    for s in range(min(2, args.warmup)):  # warm the host-buffer path (pinned staging, first-touch)
        random.seed(1000 + s)
        api.score_images(eng, [pool[i] for i in step_images(s)], AUGS, chunk=B)
    barrier()
    t_wall = time.time()
    eng.event_record(2)
    for s in range(args.warmup, n_steps_total):
        idx = step_images(s)
        random.seed(s)
        api.score_images(eng, [pool[i] for i in idx], AUGS, chunk=B)
    eng.event_record(3)
    barrier()
    t_wall = time.time() - t_wall
    ms_e2e = max(eng.event_elapsed_ms(2, 3), 1000.0 * t_wall)  # the call is synchronous: wall clock bounds it too
    if world > 1:
        t = torch.tensor([ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e = world * B * args.steps / (ms_e2e / 1000.0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_hbm, peak_src = peaks()
    achieved = conv_flops / (conv_ms / 1000.0) / 1e12 if conv_ms > 0 else 0.0
    out = {
        "metric": METRIC_RETINA if retina else METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (split-bf16, fp32 accumulate)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_RETINA if retina else WORKLOAD, "images_per_step_per_gpu": B, "views_per_image": 1 + len(AUGS),
                   "precision": args.precision,
                   "l2": "working set per step (activations of %d views, >10 GB) far exceeds the 126 MB L2; "
                         "each step scores different images" % (B * (1 + len(AUGS)))},
        "gpu_launches": int(k1 - k0),
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": B * H * W * 3 + 200 * B * 8,
                "d2h_bytes_per_step": B * (len(AUGS) + (1 + len(AUGS)) * (NUM_CLASSES - 1) + 1) * 4 + 4},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf if peak_tf else None,
                     "frac_of_bf16x3_ceiling": 3.0 * achieved / peak_tf if peak_tf else None,
                     "traffic": NCU_DRAM_BYTES_PER_CONV_LAUNCH if (not retina and B == NCU_DRAM_BATCH) else None,
                     "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read+write, mean over the 142 conv "
                                     "launches of one step; profiles/r01_igemm_dram_step.csv)",
                     "algorithmic_bytes_per_launch": conv_bytes / conv_launches if conv_launches else None,
                     "kernel": "igemm_tc_kernel + igemm_tc2_kernel (tcgen05 implicit-GEMM conv/GEMM: one-CTA and CTA-pair cta_group::2 instantiations)",
                     "launches": int(conv_launches), "kernel_ms_per_step": conv_ms / args.steps,
                     "algorithmic_gflop_per_step": conv_flops / args.steps / 1e9,
                     "share_of_step": conv_ms / ms if ms else None, "peak_source": peak_src,
                     "note": "achieved = algorithmic 2*MAC of the reference convs/GEMMs / summed CUDA-event kernel "
                             "time; the fp32-faithful bf16x3 arithmetic issues 3 tensor-core MACs per algorithmic MAC, "
                             "so the kernel's own ceiling is peak/3"},
    }
    if not args.no_cpu_baseline and world == 1:  # the CPU baseline is timed at N = 1 only
        v, cores = cpu_baseline(args.cpu_images, model=args.model)
        out["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                               "sample": "%d image(s) of the same workload, oracle port (torch CPU fp32)" % args.cpu_images}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
