"""ORACLE (test infrastructure, not product code).

Plain-``torch.nn`` fp32 restatement of the AutoencoderKL *decoder* the reference calls at
``main.py:156`` (``self.pipeline.vae.decode(output_latent).sample``) and uses for its image-space
losses (``main.py:160`` rec_loss on images, ``:168`` perturbation_loss = ``losses/losses.py:39-41``).

As for the encoder, the arithmetic lives in ``diffusers`` (not vendored, not installed): this file
restates the published ``AutoencoderKL.decode`` -> ``post_quant_conv`` -> ``Decoder`` ->
``UNetMidBlock2D`` / ``UpDecoderBlock2D`` / ``ResnetBlock2D`` / ``Upsample2D`` (nearest 2x + conv3x3).
**Parity UNPINNED** against the reference (no tests / golden vectors there); state-dict keys follow
diffusers (``decoder.*``, ``post_quant_conv.*``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline legs may import this.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .encoder_oracle import (Attention, EncoderConfig, OracleAutoencoderKL, ResnetBlock2D, UNetMidBlock2D,
                             DiagonalGaussianDistribution)


class Upsample2D(nn.Module):
    # F.interpolate(scale_factor=2, mode="nearest") then conv3x3 s1 p1
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 1, 1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, n_layers, add_upsample, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, groups, eps) for i in range(n_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Decoder(nn.Module):
    def __init__(self, cfg: EncoderConfig):
        super().__init__()
        ch = cfg.block_out_channels
        g, e = cfg.norm_num_groups, cfg.norm_eps
        rev = list(reversed(ch))
        self.conv_in = nn.Conv2d(cfg.latent_channels, rev[0], 3, 1, 1)
        self.mid_block = UNetMidBlock2D(rev[0], g, e, cfg.mid_block_add_attention)
        blocks = []
        cin = rev[0]
        for i, cout in enumerate(rev):
            blocks.append(UpDecoderBlock2D(cin, cout, cfg.layers_per_block + 1, i != len(rev) - 1, g, e))
            cin = cout
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(g, rev[-1], eps=e, affine=True)
        self.conv_out = nn.Conv2d(rev[-1], cfg.in_channels, 3, 1, 1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class OracleVAE(OracleAutoencoderKL):
    """Encoder oracle + decoder: ``decode(z)`` = ``decoder(post_quant_conv(z))`` (``.sample`` in diffusers)."""

    def __init__(self, cfg: Optional[EncoderConfig] = None):
        super().__init__(cfg)
        self.post_quant_conv = nn.Conv2d(self.cfg.latent_channels, self.cfg.latent_channels, 1)
        self.decoder = Decoder(self.cfg)

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        return self.decoder(self.post_quant_conv(z))


def make_vae_oracle(seed: int = 0, cfg: Optional[EncoderConfig] = None) -> OracleVAE:
    torch.manual_seed(seed)
    m = OracleVAE(cfg)
    m.requires_grad_(False)
    m.eval()
    return m


def image_losses(out_image, target_image, source_image, rec_lambda: float, pert_lambda: float):
    """Per-image version of main.py:159-171: rec = ||out - target||_2, pert = mse(out, source)."""
    B = out_image.shape[0]
    rec = (out_image - target_image).reshape(B, -1).norm(p=2, dim=1)
    if pert_lambda > 0:
        pert = ((out_image - source_image) ** 2).reshape(B, -1).mean(dim=1)
    else:
        pert = torch.zeros_like(rec)
    return rec_lambda * rec + pert_lambda * pert, rec, pert


def autoencoder_attack_grad(model: OracleVAE, x_adv, target_image, source_image, noise, rec_lambda=1.0,
                            pert_lambda=1.0):
    """compute_grad (main.py:144-177) with the UNet loop removed: encode -> sample -> decode -> image losses."""
    with torch.enable_grad():
        x = x_adv.clone().requires_grad_(True)
        dist = model.encode(x).latent_dist
        z = dist.mode() if noise is None else dist.sample(noise=noise)
        out = model.decode(z)
        loss, rec, pert = image_losses(out, target_image, source_image, rec_lambda, pert_lambda)
        (g,) = torch.autograd.grad(loss.sum(), [x])
    return g.detach(), loss.detach(), out.detach()
