#!/usr/bin/env python
"""Headline benchmark: image-PGD-iterations / second of the SD-1.5 VAE-encoder attack
(BASELINE.json metric; config[1]: batch 64 x 512^2, bf16 activations, fp32 iterate).

    python bench.py --gpus N --steps K --warmup W              # this framework (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle port)

One "step" = one PGD iteration for every image of the batch: encoder forward, latent loss and its
gradient, encoder input-gradient backward, fused sign-step / eps-projection / clamp.  Images are
independent: rank r owns images r::N with no data-path collective.

    --scaling strong (default)  the GLOBAL batch is fixed (configs[1]: 64 images "sharded across 1/2/4/8"):
                                64 / 32 / 16 / 8 images per GPU; at N > 1 the same run also times the weak
                                case (64 per GPU) and reports it under "weak".
    --scaling weak              64 images per GPU whatever N.
    --mode universal            configs[3]: shared-perturbation training (old/train_noise.py) over a synthetic
                                dataset sharded r::N, one NCCL all-reduce of the 3xHxW gradient per step.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FLOP_PER_IMG_ITER = {512: 2.2677e12, 1024: 10.3077e12}   # fwd + dgrad-only bwd, algorithmic
EPS, STEP, LO, HI = 32 / 255, 4 / 255, -1.0, 1.0          # eps 16/255, step 2/255 on a [0,1] scale


def read_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  The sampler is started
    before the warm-up (nvidia-smi needs ~0.2 s to come up) and every line is time-stamped; `stop(t0, t1)` keeps the
    samples that fall inside the timed region.  A region shorter than the 200 ms sampling period may contain none: the
    samples of the warm-up steps (the same work, back to back) are used then and the line says so."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []      # (host time, csv line)

    def start(self):
        if self.proc is not None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float = None, t1: float = None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.proc = None

        def parse(lines):
            sm, smax, reasons = [], None, set()
            for _, ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax = float(f[2])
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            sm.sort()
            return sm, smax, reasons

        inside = [x for x in self.lines if t0 is None or (t0 <= x[0] <= t1)]
        window = "timed region"
        if not inside:   # region shorter than the sampling period: fall back to the warm-up steps right before it
            inside = [x for x in self.lines if x[0] <= (t1 if t1 is not None else float("inf"))][-3:]
            window = "warm-up steps immediately before the timed region (region shorter than the 200 ms sampling period)"
        sm, smax, reasons = parse(inside)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm), "window": window}


def workload_config(res: int, B: int, world: int, mb: int, streams: int = 1, scaling: str = "strong"):
    name = "SDXL VAE-encoder PGD attack (BASELINE configs[2])" if res == 1024 else \
        "SD-1.5 VAE-encoder PGD attack (BASELINE configs[1])"
    return {"workload": f"{name}: global batch {B * world} x {res}^2 ({B} per GPU, {scaling} scaling), "
                        f"bf16 activations/weights, fp32 accumulate, fp32 iterate, linf eps=16/255 step=2/255, "
                        f"random-init weights",
            "global_batch": B * world, "per_gpu_batch": B, "micro_batch": min(mb, B), "streams": streams,
            "resolution": res,
            "parallelism": f"dp{world} (images r::{world}, independent, no data-path collective)",
            "l2_policy": "inputs larger than L2 (per-step working set >> 126 MB)"}


def per_gpu_batch(args, world: int) -> int:
    """--batch is the configuration's batch: the global one under strong scaling, the per-GPU one under weak."""
    if args.scaling == "weak":
        return args.batch
    if args.batch % world:
        raise SystemExit(f"--scaling strong: global batch {args.batch} is not divisible by {world} GPUs")
    return args.batch // world


def synth_inputs(batch: int, res: int, seed: int, pin: bool):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((batch, 3, res, res), generator=g) * 2 - 1
    tgt = torch.randn((batch, 4, res // 8, res // 8), generator=g)
    noise = torch.randn((batch, 4, res // 8, res // 8), generator=g)
    if pin:
        x, tgt, noise = x.pin_memory(), tgt.pin_memory(), noise.pin_memory()
    return x, tgt, noise


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU path (oracle port; the reference cannot be imported: no diffusers)
# --------------------------------------------------------------------------------------------------
def time_cpu_oracle(res: int, images_per_step: int, steps: int, warmup: int, threads: int):
    from oracle.encoder_oracle import make_oracle
    from oracle.pgd_oracle import encoder_attack
    torch.set_num_threads(threads)
    model = make_oracle(0)
    x, tgt, noise = synth_inputs(images_per_step, res, 0, pin=False)
    if warmup:
        encoder_attack(model, x, tgt, noise, warmup, EPS, STEP, LO, HI, kind=0)
    t0 = time.perf_counter()
    encoder_attack(model, x, tgt, noise, steps, EPS, STEP, LO, HI, kind=0)
    dt = time.perf_counter() - t0
    return images_per_step * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    per_step = 1
    value, dt = time_cpu_oracle(args.res, per_step, args.steps, min(args.warmup, 1), threads)
    sample = (f"bounded sample of the workload: {per_step} image(s) x {args.res}^2 fp32 per step, {args.steps} timed PGD steps after "
              f"{min(args.warmup, 1)} warm-up, oracle port (torch CPU, {threads} threads)")
    line = {
        "impl": "reference", "metric": "image-PGD-iters/sec", "value": value, "unit": "image-PGD-iters/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.res, per_gpu_batch(args, max(1, args.gpus)), max(1, args.gpus), args.micro_batch,
                                  args.streams, args.scaling),
        "cpu_baseline": {"value": value, "unit": "image-PGD-iters/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "image-PGD-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
# this framework
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from tml_image_editing_defense_b200 import _lib, ops
    from tml_image_editing_defense_b200.configs import TrainConfig
    from tml_image_editing_defense_b200.trainer import Trainer
    from tml_image_editing_defense_b200.vae import AutoencoderKL
    from tml_image_editing_defense_b200.weights import random_init_state_dict

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    res, mb = args.res, args.micro_batch
    B = per_gpu_batch(args, world)
    weights = random_init_state_dict(seed=0, include_decoder=args.loss == "images")
    vae = AutoencoderKL(device=str(dev)).load_state_dict(weights)
    del weights
    cfg = TrainConfig(norm_type="linf", eps=EPS, step_size=STEP, grad_reps=1, override_from_norm_type=False,
                      n_optimization_steps=1, device=str(dev), apply_loss_on_images=args.loss == "images",
                      apply_loss_on_latents=args.loss != "images",      # the encoder attack sets the latent mode explicitly
                      perturbation_loss_lambda=1.0 if args.loss == "images" else 0.0)
    tr = Trainer(cfg, vae, micro_batch=mb, num_streams=args.streams)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_device(Bl: int, steps: int, warm: int, timed_gemms: bool):
        """Device-resident throughput of `steps` PGD iterations over Bl images per GPU (inputs already in HBM).
        Returns (ms max over ranks, this rank's tensors, GEMM timing tuple or None, launches)."""
        xh_, th_, nh_ = synth_inputs(Bl, res, 1000 + rank, pin=True)
        x_ = xh_.to(dev)
        tgt_, noise_ = th_.to(dev), nh_.to(dev)
        xa_ = x_.clone()
        grad_ = torch.empty_like(x_)
        tgt_img_ = None
        if args.loss == "images":   # the reference's default losses on decoded images (main.py:156-171), UNet removed
            tgt_img_ = (torch.rand((Bl, 3, res, res), generator=torch.Generator().manual_seed(7)) * 2 - 1).to(dev)

        def step_device():
            tr.compute_grad(xa_, None, x_, tgt_img_, tgt_, [noise_], grad_out=grad_, beta=0.0)
            tr.perturbation_step(xa_, grad_, x_, None)

        if timed_gemms and rank == 0:
            sampler.start()                  # up before the warm-up; only the samples inside the timed region are kept
        for _ in range(warm):
            step_device()
        barrier()
        c0 = _lib.launch_counts()
        if timed_gemms:
            lib.tml_gemm_timing_enable(200000)
        barrier()
        t_start = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_device()
        e1.record()
        barrier()
        t_end = time.time()
        ms_ = e0.elapsed_time(e1)
        gt_ = None
        if timed_gemms:
            if rank == 0:
                clock_box.append(sampler.stop(t_start, t_end))
            gt_ = (C.c_double * 4)()
            lib.tml_gemm_timing_collect(gt_)
        c1 = _lib.launch_counts()
        t_ = torch.tensor([ms_], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item()), (xh_, th_, nh_, x_, tgt_, noise_, xa_, grad_), gt_, (c1[0] - c0[0]) + (c1[1] - c0[1])

    # ------------------------------------------------------------------ device-resident throughput
    n_warm = args.warmup if args.quick else max(args.warmup, 3)
    sampler = ClockSampler(local)
    clock_box = []
    ms_max, tensors, gt, launches = time_device(B, args.steps, n_warm, True)
    xh, th, nh, x, tgt, noise, x_adv, grad = tensors
    if args.gemm_table and rank == 0:
        print_gemm_table(lib, args.steps)
    lib.tml_gemm_timing_enable(0)
    clocks = clock_box[0] if clock_box else None
    value = world * B * args.steps / (ms_max / 1e3)

    # the other scaling mode, timed in the same run (N > 1 only: at N = 1 the two coincide)
    other = None
    if world > 1 and not args.quick and args.loss == "latents":
        Bo = args.batch if args.scaling == "strong" else (args.batch // world if args.batch % world == 0 else 0)
        if Bo > 0 and Bo != B:
            ko = max(2, min(args.steps, 5))
            ms_o, _t, _g, _l = time_device(Bo, ko, 2, False)
            del _t
            torch.cuda.empty_cache()
            other = {"scaling": "weak" if args.scaling == "strong" else "strong", "per_gpu_batch": Bo,
                     "global_batch": Bo * world, "steps": ko, "ms_per_step": ms_o / ko,
                     "value": world * Bo * ko / (ms_o / 1e3), "unit": "image-PGD-iters/s"}

    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "value": value, "ms_per_step": ms_max / args.steps,
                              "gemm_ms_per_step": gt[0] / args.steps, "gemm_launches": int(gt[2]),
                              "launches": launches, "clocks": clocks}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------ end to end with host buffers
    # Per step: H2D of the step's inputs (current iterate, source image, target latent, noise) from
    # pinned memory, the PGD iteration through the public Trainer API, D2H of the new iterate + losses.
    # Every byte is copied every step; only the order is pipelined.
    # The iterate ping-pongs between two pinned host buffers (the result of step i is the input of
    # step i+1); the step-invariant inputs (source, target, noise) are uploaded on a side stream while
    # the previous step computes, and within a step the iterate moves micro-batch by micro-batch.
    xa_bufs = [xh.clone().pin_memory(), torch.empty_like(xh).pin_memory()]
    loss_host = torch.empty(B, dtype=torch.float32).pin_memory()
    h2d = xh.numel() * 4 + xh.numel() * 4 + th.numel() * 4 + nh.numel() * 4
    d2h = xh.numel() * 4 + loss_host.numel() * 4
    copy_stream = torch.cuda.Stream(device=dev)
    state = {"cur": 0, "pre": None}

    def upload_invariants():
        with torch.cuda.stream(copy_stream):
            x_d = xh.to(dev, non_blocking=True)
            t_d = th.to(dev, non_blocking=True)
            n_d = nh.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return x_d, t_d, n_d, ev

    d2h_stream = torch.cuda.Stream(device=dev)
    chunks = [(s_, min(B, s_ + mb)) for s_ in range(0, B, mb)]

    def step_e2e():
        """Images are independent, so the step is pipelined per micro-batch: the iterate of micro-batch m+1 is uploaded
        while m computes, the result of m is downloaded while m+1 computes.  Every byte still moves every step."""
        main = torch.cuda.current_stream()
        x_d, t_d, n_d, ev = state["pre"] if state["pre"] is not None else upload_invariants()
        src, dst = xa_bufs[state["cur"]], xa_bufs[state["cur"] ^ 1]
        parts = []
        with torch.cuda.stream(copy_stream):            # the iterate first, micro-batch by micro-batch ...
            for s_, e_ in chunks:
                xa_m = src[s_:e_].to(dev, non_blocking=True)
                ev_m = torch.cuda.Event()
                ev_m.record(copy_stream)
                parts.append((xa_m, ev_m))
        main.wait_event(ev)
        state["pre"] = upload_invariants()              # ... then the next step's invariants, under this step's kernels
        for (s_, e_), (xa_m, ev_m) in zip(chunks, parts):
            main.wait_event(ev_m)
            xa_m.record_stream(main)
            g_m, _, _, ld = tr.compute_grad(xa_m, None, x_d[s_:e_], None, t_d[s_:e_], [n_d[s_:e_]])
            xa_m = tr.perturbation_step(xa_m, g_m, x_d[s_:e_], None)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                dst[s_:e_].copy_(xa_m, non_blocking=True)
                loss_host[s_:e_].copy_(ld["per_image"], non_blocking=True)
            xa_m.record_stream(d2h_stream)
            ld["per_image"].record_stream(d2h_stream)
        for t_ in (x_d, t_d, n_d):
            t_.record_stream(main)
        d2h_stream.synchronize()                        # the caller owns the result after this
        main.synchronize()
        state["cur"] ^= 1

    for _ in range(2):
        step_e2e()
    barrier()
    k2 = max(2, args.steps)      # the end-to-end number is timed over all --steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k2):
        step_e2e()
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * k2 / (float(t2.item()) / 1e3)
    state["pre"] = None

    # ------------------------------------------------------------------ K9 standalone (HBM roofline)
    pgd = None
    if rank == 0:
        n = x_adv.numel()
        for _ in range(3):
            ops.pgd_step_linf_(x_adv, grad, x, EPS, STEP, LO, HI)
        torch.cuda.synchronize()
        reps = 20
        e0.record()
        for _ in range(reps):
            ops.pgd_step_linf_(x_adv, grad, x, EPS, STEP, LO, HI)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        pgd = {"bytes": 16 * n, "us": us, "gbs": 16 * n / (us * 1e-6) / 1e9}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_src = read_peaks()
    gemm_ms, _, gemm_n, gemm_drop = gt[0], gt[1], int(gt[2]), int(gt[3])
    algo_flops_step = FLOP_PER_IMG_ITER.get(res, 0.0) * B   # every conv / linear / attention product is a launch of this kernel
    gemm_ms_step = gemm_ms / args.steps if args.steps else 0.0
    achieved_tf = algo_flops_step / (gemm_ms_step * 1e-3) / 1e12 if gemm_ms_step > 0 else None
    peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    roofline = {
        "bound": "tensor", "kernel": "conv3x3_swapped_kernel + conv_gemm_tcgen05_kernel (all tcgen05 GEMM launches)",
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": (achieved_tf / peak_tf) if achieved_tf else None,
        # dram bytes of ONE launch of the dominant instantiation (128->128 3x3 at 512^2, 16 images), ncu --set full
        "traffic": 2.127e9 if res == 512 else None,
        "traffic_source": "profiles/r01_prof_swap_c128_r1.txt: dram read 1.087 GB + write 1.040 GB per launch of "
                          "conv3x3_swapped_kernel<1,0,0,0> on 16 x 512^2 x 128 ch = its algorithmic bytes (input once, "
                          "output once: 2 x 1.074 GB); the aggregate `achieved` above covers every tcgen05 launch",
        "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
        "launches_timed": gemm_n, "launches_dropped": gemm_drop,
        "gemm_ms_per_step": gemm_ms_step, "gemm_share_of_step": gemm_ms_step / (ms_max / args.steps),
        "algorithmic_flops_per_step": algo_flops_step,
        "whole_step_frac_of_burst": (value / world) * FLOP_PER_IMG_ITER.get(res, 0.0) / (peaks["bf16_tflops"] * 1e12),
    }
    roofline_pgd = None
    if pgd:
        roofline_pgd = {"bound": "hbm", "kernel": "pgd_linf_kernel", "achieved": pgd["gbs"], "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": pgd["gbs"] / peaks["hbm_gbs"], "bytes_per_launch": pgd["bytes"],
                        "us_per_launch": pgd["us"], "peak_source": f"{peak_src} hbm_gbs",
                        "traffic": 762.7e6 if (res == 512 and B == 64) else None,
                        "traffic_source": "ncu --set full, profiles/r01_prof_pgd_r1_b64.txt (dram read 604.0 MB + write 158.7 MB; "
                                          "the remaining writes were still in L2 at kernel end)",
                        "note": "16 B/element algorithmic; working set 805 MB > 126 MB L2"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = time_cpu_oracle(res, 1, 2, 1, threads)
        cpu = {"value": v, "unit": "image-PGD-iters/s", "cores": threads, "kind": "port",
               "sample": f"1 image x {res}^2 fp32, 2 timed PGD steps after 1 warm-up ({dt:.1f} s), oracle port, "
                         f"torch CPU {threads} threads"}

    line = {
        "metric": "image-PGD-iters/sec", "value": value, "unit": "image-PGD-iters/s", "n_gpus": world,
        "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(res, B, world, mb, args.streams, args.scaling),
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "image-PGD-iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": k2},
        "roofline": roofline, "roofline_pgd_update": roofline_pgd, "cpu_baseline": cpu,
    }
    if other is not None:
        line[other["scaling"]] = other
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_universal(args):
    """BASELINE configs[3]: universal-perturbation training (old/train_noise.py:115-185) over a synthetic dataset of
    --dataset images sharded r::N.  One step = one update of the shared delta: every rank evaluates the encoder-attack
    gradient of its images at x_i + delta, sums them in image order, ONE fp32 [1,3,H,W] NCCL all-reduce, then the
    identical L2-normalised update, +-eps clamp and image-range re-projection on every rank."""
    import torch.distributed as dist
    from tml_image_editing_defense_b200 import _lib
    from tml_image_editing_defense_b200.configs import UniversalConfig
    from tml_image_editing_defense_b200.dataset import SyntheticImageDataset, shard_indices
    from tml_image_editing_defense_b200.universal import UniversalTrainer
    from tml_image_editing_defense_b200.vae import AutoencoderKL
    from tml_image_editing_defense_b200.weights import random_init_state_dict

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    res, n_global, mb = args.res, args.dataset, args.micro_batch
    vae = AutoencoderKL(device=str(dev)).load_state_dict(random_init_state_dict(seed=0))
    ucfg = UniversalConfig(grad_reps=1, eps=16 / 255 * 2, step_size=1.0, resolution=res, device=str(dev))
    ut = UniversalTrainer.for_b200(ucfg, vae)
    ut.comm_events = []
    idx = shard_indices(n_global, rank, world)
    ds = SyntheticImageDataset(n_global, resolution=res, seed=0)
    images = ds.batch(idx).to(dev)
    g = torch.Generator().manual_seed(17)
    tg = torch.randn((1, 4, res // 8, res // 8), generator=g).to(dev).expand(len(idx), -1, -1, -1).contiguous()
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    delta = torch.zeros((1, 3, res, res), device=dev)
    ut.prepare_projection(images)

    def step(d):
        nz = torch.randn(tg.shape, generator=gen, device=dev)      # latent_dist.sample(generator), :133
        return ut.step(d, images, tg, nz, n_global=n_global, micro_batch=mb)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 1)):
        delta = step(delta)
    barrier()
    ut.comm_events = []
    c0 = _lib.launch_counts()
    t_start = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        delta = step(delta)
    e1.record()
    barrier()
    clocks = sampler.stop(t_start, time.time()) if rank == 0 else None
    c1 = _lib.launch_counts()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    comm_us = [1e3 * a.elapsed_time(b) for a, b in ut.comm_events]
    # per step, the rank that arrived last waited least: the minimum over ranks of that step's interval ~ the collective
    # itself; the other ranks' intervals include the wait for the slowest rank
    tc = torch.tensor(comm_us if comm_us else [0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tc, op=dist.ReduceOp.MIN)
    tc = tc.mean().reshape(1)
    same = ut.check_replicas_identical(delta)
    in_range = bool(float((images + delta).abs().max()) <= 1.0 + 1e-6) if len(idx) else True
    if rank == 0:
        ms = float(t.item())
        payload = delta.numel() * 4
        line = {
            "mode": "universal", "metric": "image-grad-evals/sec (universal perturbation step)",
            "value": n_global * args.steps / (ms / 1e3), "unit": "image-grad-evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[3]: universal perturbation (old/train_noise.py) over {n_global} synthetic "
                                   f"{res}^2 images sharded r::{world}, grad_reps 1, one fp32 [1,3,{res},{res}] NCCL "
                                   f"all-reduce per step, eps 32/255 step 1.0, image-range projection on",
                       "dataset": n_global, "images_per_gpu": len(idx), "micro_batch": mb, "resolution": res},
            "allreduce": {"payload_bytes": payload, "per_step": 1,
                          "us_mean_rank0_incl_wait": sum(comm_us) / max(1, len(comm_us)),
                          "us_mean_of_per_step_min_over_ranks": float(tc.item()),
                          "share_of_step": (float(tc.item()) * 1e-3) / (ms / args.steps) if world > 1 else 0.0,
                          "backend": "nccl" if world > 1 else "none (single rank)"},
            "replicas_identical": same, "images_plus_delta_in_range": in_range,
            "delta_abs_max": float(delta.abs().max()), "clocks": clocks,
            "gpu_launches": (c1[0] - c0[0]) + (c1[1] - c0[1]),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def print_gemm_table(lib, steps):
    """Per-shape in-situ GEMM times of the launches recorded since tml_gemm_timing_enable (stderr)."""
    buf = C.create_string_buffer(1 << 18)
    n = lib.tml_gemm_timing_report(buf, len(buf))
    rows = []
    for ln in buf.raw[:n].decode().strip().split("\n"):
        f = ln.split("|")
        rows.append((f[0], int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5]), float(f[6]), float(f[7])))
    rows.sort(key=lambda r: -r[6])
    tot_ms = sum(r[6] for r in rows)
    tot_fl = sum(r[7] for r in rows)
    print(f"{'gemm':28s} {'M':>9s} {'N':>6s} {'K':>6s} {'mode':>4s} {'n':>5s} {'ms/step':>8s} {'TFLOP/s':>8s}", file=sys.stderr)
    for r in rows:
        print(f"{r[0]:28s} {r[1]:9d} {r[2]:6d} {r[3]:6d} {r[4]:4d} {r[5]:5d} {r[6] / steps:8.3f} "
              f"{r[7] / (r[6] * 1e-3) / 1e12:8.1f}", file=sys.stderr)
    print(f"{'total':28s} {'':9s} {'':6s} {'':6s} {'':4s} {'':5s} {tot_ms / steps:8.3f} "
          f"{tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms else 0:8.1f}", file=sys.stderr)


def run_diffusion(args):
    """BASELINE configs[4], informational: backprop through the img2img DDIM steps of the SD-1.5 UNet at 512^2 with
    activation checkpointing.  Encoder / decoder / PGD update are this repo's kernels; the UNet is the PyTorch
    library module (unet_torch.py) with random-init weights."""
    from tml_image_editing_defense_b200 import _lib, ops
    from tml_image_editing_defense_b200.configs import TrainConfig
    from tml_image_editing_defense_b200.diffusion import DiffusionAttack
    from tml_image_editing_defense_b200.schedulers import DDIMScheduler
    from tml_image_editing_defense_b200.unet_torch import UNet2DConditionModel
    from tml_image_editing_defense_b200.vae import AutoencoderKL
    from tml_image_editing_defense_b200.weights import random_init_state_dict
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:   # images shard r::world with no data-path collective, exactly like the encoder attack
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, res = args.batch, args.res   # images per GPU (weak scaling: the step of one image is what the memory bounds)
    vae = AutoencoderKL(device=str(dev)).load_state_dict(random_init_state_dict(seed=0, include_decoder=True))
    native = args.unet == "native"
    if native:
        # random-init weights of the SD-1.5 topology, generated on the device by the PyTorch restatement and handed to
        # the native module as a diffusers state dict
        from tml_image_editing_defense_b200.unet import UNet2DConditionModel as NativeUNet
        with torch.device(dev):
            torch.manual_seed(0)
            src = UNet2DConditionModel().requires_grad_(False)
        unet = NativeUNet(device=str(dev), keep_activations=not args.unet_recompute).load_state_dict(src.state_dict())
        del src
        torch.cuda.empty_cache()
    else:
        with torch.device(dev):
            torch.manual_seed(0)
            unet = UNet2DConditionModel().to(torch.bfloat16).requires_grad_(False)
    cfg = TrainConfig(norm_type="linf", eps=EPS, step_size=STEP, grad_reps=1, override_from_norm_type=False,
                      device=str(dev), apply_loss_on_images=True, apply_loss_on_latents=False,
                      perturbation_loss_lambda=1.0, n_denoising_steps_per_iteration=4, limit_timesteps=False)
    da = DiffusionAttack(cfg, vae, unet, DDIMScheduler(), use_checkpointing=not native,
                         unet_dtype=torch.float32 if native else torch.bfloat16)
    g = torch.Generator().manual_seed(rank)
    x = (torch.rand((B, 3, res, res), generator=g) * 2 - 1).to(dev)
    tgt = (torch.rand((B, 3, res, res), generator=g) * 2 - 1).to(dev)
    pe = torch.randn((2, 77, 768), generator=g).to(dev)
    nz = [torch.randn((B, 4, res // 8, res // 8), generator=g).to(dev)]
    x_adv = x.clone()

    def step():
        grad, loss, _, _ = da.compute_grad(x_adv, pe, x, tgt, None, nz)
        ops.pgd_step_linf_(x_adv, grad.contiguous(), x, EPS, STEP, LO, HI)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):   # (0 is allowed: profiler runs)
        step()
    barrier()
    c0 = _lib.launch_counts()
    if args.gemm_table:
        _lib.load().tml_gemm_timing_enable(200000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    c1 = _lib.launch_counts()
    ms = e0.elapsed_time(e1)
    if world > 1:   # device time, max over ranks
        t_ = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        ms = float(t_.item())
    extra = {}
    if args.gemm_table:
        gt = (C.c_double * 4)()
        _lib.load().tml_gemm_timing_collect(gt)
        extra = {"gemm_ms_per_step": gt[0] / args.steps, "gemm_tflops_in_situ": gt[1] / (gt[0] * 1e-3) / 1e12 if gt[0] else None,
                 "gemm_launches_timed": int(gt[2])}
        print_gemm_table(_lib.load(), args.steps)
        _lib.load().tml_gemm_timing_enable(0)
    unet_desc = ("this repo's native UNet (csrc/unet.cu: tcgen05 GEMMs, bf16 activations, "
                 + ("activations recomputed per denoising step in the backward)" if args.unet_recompute else
                    "activations of all denoising steps kept in HBM, attention probabilities recomputed)") if native else
                 "the PyTorch library UNet (unet_torch.py: cuDNN / cuBLAS / SDPA, bf16 autocast, torch.utils.checkpoint)")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    print(json.dumps({"mode": "diffusion", "unet": args.unet, "metric": "image-PGD-iters/sec",
                      "value": world * B * args.steps / (ms / 1e3), "n_gpus": world, "scaling": "weak",
                      "per_gpu_batch": B,
                      "ms_per_step": ms / args.steps, "steps": args.steps, "loss": float(loss),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                      "gpu_launches": (c1[0] - c0[0]) + (c1[1] - c0[1]), **extra,
                      "config": {"workload": f"BASELINE configs[4]: diffusion attack, 4 DDIM steps of the SD-1.5 UNet with "
                                             f"classifier-free guidance on {unet_desc} + encoder/decoder/update on this "
                                             f"repo's kernels, batch {B} x {res}^2, image-space losses, random-init weights"}}),
          flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--batch", type=int, default=64,
                    help="the configuration's batch (configs[1]: 64; configs[2]: --res 1024 --batch 16): the GLOBAL batch "
                         "under --scaling strong, the per-GPU batch under --scaling weak")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: global batch fixed, images sharded r::N (BASELINE configs[1]); weak: --batch per GPU")
    ap.add_argument("--dataset", type=int, default=512, help="--mode universal: images in the synthetic dataset (global)")
    ap.add_argument("--micro_batch", type=int, default=0,
                    help="images per encoder pass (0 = 32 up to 512^2, 16 above: measured 16 -> 32: +1 %% device, +1.7 %% end to end)")
    ap.add_argument("--streams", type=int, default=1, help="CUDA streams the micro-batches alternate on")
    ap.add_argument("--loss", default="latents", choices=["latents", "images"],
                    help="latents: the BASELINE encoder attack; images: + decoder and image-space losses (needs --quick)")
    ap.add_argument("--mode", default="encoder", choices=["encoder", "universal", "diffusion"],
                    help="universal: BASELINE configs[3] (shared perturbation, one NCCL all-reduce per step); "
                         "diffusion: BASELINE configs[4] (4 DDIM steps of the SD-1.5 UNet, checkpointed; informational)")
    ap.add_argument("--unet", default="native", choices=["native", "torch"],
                    help="--mode diffusion: this repo's UNet kernels (default) or the PyTorch library module (baseline)")
    ap.add_argument("--unet_recompute", action="store_true",
                    help="--mode diffusion --unet native: re-run each UNet call's forward in the backward (checkpointing) "
                         "instead of keeping its activations (about 10 GB per call at batch 8 with guidance)")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--gemm_table", action="store_true", help="print per-shape GEMM times (stderr)")
    ap.add_argument("--quick", action="store_true", help="device-resident timing only (profiling runs)")
    args = ap.parse_args()
    if args.micro_batch <= 0:
        args.micro_batch = 32 if args.res <= 512 else 16
    if args.mode == "diffusion":
        return run_diffusion(args)
    if args.mode == "universal":
        if args.impl == "reference":
            raise SystemExit("--mode universal has no reference arm")
        return run_universal(args)
    if args.loss == "images" and not args.quick:
        raise SystemExit("--loss images is an informational mode: use it with --quick")
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
