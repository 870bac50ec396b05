"""Command-line flags for the launcher, named after the (unused) argparse list the reference
carries in utils/parser.py (``--resolution`` :116, ``--train_batch_size`` :139, ``--seed`` :114,
``--mixed_precision`` :322, ``--local_rank`` :332, ``--pretrained_vae_model_name_or_path``,
``--max_train_steps``, ``--output_dir``, ``--gradient_checkpointing`` :180, ``--allow_tf32`` :266) plus the
PGD options of configs.py:121-141."""
from __future__ import annotations

import argparse
import os


def parse_args(input_args=None):
    p = argparse.ArgumentParser(description="B200-native PGD image immunization (VAE-encoder attack)")
    p.add_argument("--pretrained_vae_model_name_or_path", type=str, default=None,
                   help="Path to a torch state dict (diffusers AutoencoderKL keys); random init if omitted.")
    p.add_argument("--train_data_dir", type=str, default=None, help="Folder of *.jpg images; synthetic if omitted.")
    p.add_argument("--num_images", type=int, default=64)
    p.add_argument("--images_per_pass", type=int, default=64,
                   help="Images of this rank's shard attacked together (one PGD run per pass; the next pass is prefetched).")
    p.add_argument("--output_dir", type=str, default="./output")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--resolution", type=int, default=512)
    p.add_argument("--train_batch_size", type=int, default=16, help="Images per encoder pass on each GPU.")
    p.add_argument("--max_train_steps", type=int, default=200, help="PGD steps (n_optimization_steps).")
    p.add_argument("--mixed_precision", type=str, default="bf16", choices=["bf16"],
                   help="Activations/weights inside the encoder kernels; the iterate stays fp32.")
    p.add_argument("--local_rank", type=int, default=-1)
    p.add_argument("--gradient_checkpointing", action="store_true",
                   help="Diffusion attack: recompute each UNet step in the backward (activation checkpointing, "
                        "utils/parser.py:180).  The encoder attack keeps only what its input gradient needs.")
    p.add_argument("--allow_tf32", action="store_true",
                   help="Accepted for call-site compatibility (utils/parser.py:266); the sm_100a kernels compute in "
                        "bf16 with fp32 accumulation and take no TF32 path, so the flag changes nothing here.")
    p.add_argument("--norm_type", type=str, default="linf", choices=["linf", "l2"])
    p.add_argument("--eps", type=float, default=32 / 255)
    p.add_argument("--step_size", type=float, default=4 / 255)
    p.add_argument("--min_value", type=float, default=-1.0)
    p.add_argument("--max_value", type=float, default=1.0)
    p.add_argument("--grad_reps", type=int, default=1)
    p.add_argument("--latent_loss", type=str, default="l2norm", choices=["l2norm", "mse"])
    p.add_argument("--universal", action="store_true", help="Shared-perturbation mode (old/train_noise.py).")
    p.add_argument("--diffusion", action="store_true",
                   help="The reference's full attack (main.py:144-246): encode, add_noise, UNet denoising steps with "
                        "classifier-free guidance, decode, image-space losses -- every network on the sm_100a kernels.")
    p.add_argument("--pretrained_unet_model_name_or_path", type=str, default=None,
                   help="--diffusion: torch state dict with diffusers UNet2DConditionModel keys; random init if omitted.")
    p.add_argument("--n_denoising_steps_per_iteration", type=int, default=4)
    p.add_argument("--guidance_scale", type=float, default=3.0)   # configs.py default
    p.add_argument("--sdxl", action="store_true", help="SDXL VAE scaling factor (architecture is identical).")
    args = p.parse_args(input_args)
    env_local_rank = int(os.environ.get("LOCAL_RANK", -1))
    if env_local_rank != -1 and env_local_rank != args.local_rank:
        args.local_rank = env_local_rank
    return args
