"""Command-line driver of the PGD immunization hot path (the `__main__` of the reference's main.py:592-651
and the image loop of run_all.py:23-93, for the VAE-encoder attack).

    python -m tml_image_editing_defense_b200.main --num_images 64 --max_train_steps 200
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 -m tml_image_editing_defense_b200.main --num_images 512
    torchrun ... -m tml_image_editing_defense_b200.main --universal        # shared perturbation (old/train_noise.py)
    python -m tml_image_editing_defense_b200.main --diffusion --num_images 8 --images_per_pass 8 --max_train_steps 50

One process per GPU; rank r immunizes images r::world (no communication; run_all.py:14-21 split the list by
hand over two GPUs).  Results: `<output_dir>/adversarial_rank{r}.pt` (+ PNGs when torchvision is present, as
main.py:618) and one JSON line of metrics per rank (the reference logged to wandb).
"""
from __future__ import annotations

import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

from .configs import TrainConfig, UniversalConfig
from .dataset import ImagePromptDataset, SyntheticImageDataset, shard_indices
from .loader import ShardedImageLoader
from .parser import parse_args
from .trainer import Trainer
from .universal import UniversalTrainer
from .vae import SD15_VAE, SDXL_VAE, AutoencoderKL
from .weights import load_checkpoint, random_init_state_dict


def main(argv=None) -> int:
    args = parse_args(argv)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0")) if args.local_rank < 0 else args.local_rank
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    out_dir = Path(args.output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)

    cfg_vae = SDXL_VAE if args.sdxl else SD15_VAE
    sd = load_checkpoint(args.pretrained_vae_model_name_or_path) if args.pretrained_vae_model_name_or_path \
        else random_init_state_dict(cfg_vae, seed=args.seed, include_decoder=args.diffusion)
    vae = AutoencoderKL(cfg_vae, device=dev).load_state_dict(sd)

    if args.train_data_dir:
        ds = ImagePromptDataset(args.train_data_dir, "", resolution=args.resolution)
        fetch = None
    else:
        ds = SyntheticImageDataset(args.num_images, resolution=args.resolution, seed=args.seed)
        fetch = ds.batch
    n = len(ds)
    idx = shard_indices(n, rank, world)
    # target latent: encoding of a fixed other image (main.py:75); synthetic: a seeded N(0,1) latent
    g = torch.Generator().manual_seed(args.seed + 17)
    lat = (1, 4, args.resolution // 8, args.resolution // 8)
    target = torch.randn(lat, generator=g).to(dev)

    t0 = time.perf_counter()
    if args.diffusion:
        # the reference's default mode (main.py:144-246) with every network on this repo's kernels
        from .diffusion import DiffusionAttack
        from .schedulers import DDIMScheduler
        from .unet import UNet2DConditionModel
        from .unet_torch import UNet2DConditionModel as TorchUNet
        if not vae.has_decoder:
            raise SystemExit("--diffusion needs the VAE decoder weights (decoder.*, post_quant_conv.*)")
        if args.pretrained_unet_model_name_or_path:
            usd = load_checkpoint(args.pretrained_unet_model_name_or_path)
        else:   # random init of the SD-1.5 topology (no network for checkpoints)
            with torch.device(dev):
                torch.manual_seed(args.seed)
                usd = TorchUNet().requires_grad_(False).state_dict()
        unet = UNet2DConditionModel(device=dev, keep_activations=not args.gradient_checkpointing).load_state_dict(usd)
        del usd
        cfg = TrainConfig(norm_type=args.norm_type, eps=args.eps, step_size=args.step_size, grad_reps=args.grad_reps,
                          min_value=args.min_value, max_value=args.max_value, override_from_norm_type=False,
                          n_optimization_steps=args.max_train_steps, seed=args.seed, device=dev,
                          resolution=args.resolution, output_path=out_dir,
                          n_denoising_steps_per_iteration=args.n_denoising_steps_per_iteration,
                          guidance_scale=args.guidance_scale)
        da = DiffusionAttack(cfg, vae, unet, DDIMScheduler(), use_checkpointing=False, unet_dtype=torch.float32)
        ge = torch.Generator().manual_seed(args.seed + 29)
        prompt_embeds = torch.randn((2, 77, unet.config.cross_attention_dim), generator=ge).to(dev)   # no CLIP offline
        target_image = (torch.rand((1, 3, args.resolution, args.resolution), generator=ge) * 2 - 1).to(dev)
        loader = ShardedImageLoader(ds, max(1, args.images_per_pass), rank, world, dev, fetch=fetch)
        outs, hist = [], []
        for _, images, _ in loader:
            outs.append(da.run(images, target_image, prompt_embeds).cpu())
            hist = hist or list(da.loss_history)
        x_adv = torch.cat(outs) if outs else torch.empty((0, 3, args.resolution, args.resolution))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        torch.save({"indices": idx, "x_adv": x_adv}, out_dir / f"adversarial_rank{rank}.pt")
        print(json.dumps({"rank": rank, "mode": "diffusion", "steps": args.max_train_steps, "images": len(idx),
                          "seconds": dt, "image_pgd_iters_per_s": len(idx) * args.max_train_steps / dt,
                          "loss_first": hist[0] if hist else None, "loss_last": hist[-1] if hist else None}), flush=True)
    elif args.universal:
        images = torch.cat([im for _, im, _ in ShardedImageLoader(ds, max(1, args.train_batch_size), rank, world, dev, fetch=fetch)]) \
            if idx else torch.empty((0, 3, args.resolution, args.resolution), device=dev)
        ucfg = UniversalConfig(grad_reps=args.grad_reps, eps=args.eps, step_size=args.step_size,
                               resolution=args.resolution, latent_loss=args.latent_loss, max_steps=args.max_train_steps)
        ut = UniversalTrainer.for_b200(ucfg, vae)
        delta = torch.zeros((1, 3, args.resolution, args.resolution), device=dev)
        tg = target.expand(len(idx), -1, -1, -1).contiguous()
        ut.prepare_projection(images)
        # latent_dist.sample(generator) (old/train_noise.py:133): seeded posterior noise, fresh every step, per image
        gen = torch.Generator(device=dev).manual_seed(args.seed * 7919 + rank)
        for _ in range(args.max_train_steps):
            nz = torch.randn(tg.shape, generator=gen, device=dev)
            delta = ut.step(delta, images, tg, nz, n_global=n, micro_batch=args.train_batch_size)
        same = ut.check_replicas_identical(delta)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        torch.save(delta.cpu(), out_dir / f"universal_delta_rank{rank}.pt")
        print(json.dumps({"rank": rank, "mode": "universal", "steps": args.max_train_steps, "images": len(idx),
                          "seconds": dt, "image_grad_evals_per_s": len(idx) * args.max_train_steps * args.grad_reps / dt,
                          "replicas_identical": same, "delta_abs_max": float(delta.abs().max())}), flush=True)
    else:
        cfg = TrainConfig.encoder_attack(norm_type=args.norm_type, eps=args.eps, step_size=args.step_size, grad_reps=args.grad_reps,
                          min_value=args.min_value, max_value=args.max_value, override_from_norm_type=False,
                          n_optimization_steps=args.max_train_steps, latent_loss=args.latent_loss, seed=args.seed,
                          device=dev, resolution=args.resolution, output_path=out_dir)
        tr = Trainer(cfg, vae, micro_batch=args.train_batch_size)
        # the shard is attacked `--images_per_pass` images at a time; the loader stages the next pass in pinned memory and
        # uploads it on a side stream while this one runs (run_all.py:23-93 walks its image list one by one)
        loader = ShardedImageLoader(ds, max(1, args.images_per_pass), rank, world, dev, fetch=fetch)
        outs, hist = [], []
        for _, images, _ in loader:
            outs.append(tr.run(images, target_latent=target).cpu())
            hist = hist or list(tr.loss_history)          # the loss trajectory reported is the first pass's
        x_adv = torch.cat(outs) if outs else torch.empty((0, 3, args.resolution, args.resolution))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        torch.save({"indices": idx, "x_adv": x_adv}, out_dir / f"adversarial_rank{rank}.pt")
        # main.py:619: the fixed training noises, so inference can replay them (main.py:622)
        torch.save([n_.cpu() for n_ in tr.noises] if tr.noises is not None else None,
                   out_dir / ("noise.pt" if world == 1 else f"noise_rank{rank}.pt"))
        try:
            for i, im in zip(idx[:4], Trainer.to_pil(x_adv[:4])):
                im.save(out_dir / f"adversarial_image_{i}.png")   # main.py:618
        except Exception:
            pass
        print(json.dumps({"rank": rank, "mode": "per-image", "steps": args.max_train_steps, "images": len(idx),
                          "seconds": dt, "image_pgd_iters_per_s": len(idx) * args.max_train_steps / dt,
                          "loss_first": hist[0] if hist else None, "loss_last": hist[-1] if hist else None}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
