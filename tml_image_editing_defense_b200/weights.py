"""State-dict helpers: diffusers key names of the AutoencoderKL encoder,
random initialisation of that architecture (benchmarks: there is no network for checkpoints), and
loading a checkpoint saved with torch.save."""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from .vae import EncoderConfig


def encoder_param_shapes(cfg: EncoderConfig, include_decoder: bool = False) -> List[Tuple[str, Tuple[int, ...]]]:
    ch = cfg.block_out_channels
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(name, co, ci, k):
        out.append((name + ".weight", (co, ci, k, k)))
        out.append((name + ".bias", (co,)))

    def norm(name, c):
        out.append((name + ".weight", (c,)))
        out.append((name + ".bias", (c,)))

    def lin(name, co, ci):
        out.append((name + ".weight", (co, ci)))
        out.append((name + ".bias", (co,)))

    def resnet(name, ci, co):
        norm(name + ".norm1", ci)
        conv(name + ".conv1", co, ci, 3)
        norm(name + ".norm2", co)
        conv(name + ".conv2", co, co, 3)
        if ci != co:
            conv(name + ".conv_shortcut", co, ci, 1)

    conv("encoder.conv_in", ch[0], cfg.in_channels, 3)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", cin, cout)
            cin = cout
        if i != len(ch) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    resnet("encoder.mid_block.resnets.0", cin, cin)
    if cfg.mid_block_add_attention:
        a = "encoder.mid_block.attentions.0"
        norm(a + ".group_norm", cin)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(f"{a}.{n}", cin, cin)
    resnet("encoder.mid_block.resnets.1", cin, cin)
    norm("encoder.conv_norm_out", cin)
    conv("encoder.conv_out", 2 * cfg.latent_channels, cin, 3)
    conv("quant_conv", 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
    if include_decoder:
        rev = list(reversed(ch))
        conv("post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
        conv("decoder.conv_in", rev[0], cfg.latent_channels, 3)
        resnet("decoder.mid_block.resnets.0", rev[0], rev[0])
        if cfg.mid_block_add_attention:
            a = "decoder.mid_block.attentions.0"
            norm(a + ".group_norm", rev[0])
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(f"{a}.{n}", rev[0], rev[0])
        resnet("decoder.mid_block.resnets.1", rev[0], rev[0])
        cin = rev[0]
        for i, cout in enumerate(rev):
            for j in range(cfg.layers_per_block + 1):
                resnet(f"decoder.up_blocks.{i}.resnets.{j}", cin, cout)
                cin = cout
            if i != len(rev) - 1:
                conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
        norm("decoder.conv_norm_out", cin)
        conv("decoder.conv_out", cfg.out_channels, cin, 3)
    return out


def random_init_state_dict(cfg: EncoderConfig = None, seed: int = 0, include_decoder: bool = False) -> Dict[str, torch.Tensor]:
    """torch.nn default initialisation (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv / linear weights
    and biases, GroupNorm gamma = 1, beta = 0), reproducible from `seed`."""
    cfg = cfg or EncoderConfig()
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    shapes = encoder_param_shapes(cfg, include_decoder)
    fan_in = {}
    for name, shape in shapes:
        if name.endswith(".weight") and len(shape) > 1:
            fan_in[name[:-7]] = int(math.prod(shape[1:]))
    for name, shape in shapes:
        base = name.rsplit(".", 1)[0]
        if base in fan_in:
            b = 1.0 / math.sqrt(fan_in[base])
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif name.endswith(".weight"):
            sd[name] = torch.ones(shape)
        else:
            sd[name] = torch.zeros(shape)
    return sd


def load_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """A torch.save'd diffusers AutoencoderKL state dict (decoder keys are ignored downstream)."""
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if "state_dict" in sd:
        sd = sd["state_dict"]
    return sd
