"""ORACLE (test infrastructure, not product code).

Plain-``torch.nn`` fp32 restatement of the AutoencoderKL *encoder* + posterior that the
reference calls at ``main.py:75,191``, ``old/train_noise.py:133`` and
``pipelines/pipeline_stable_diffusion_img2img.py:751,756`` (through ``retrieve_latents``,
``pipelines/pipeline_stable_diffusion_img2img.py:77-87``).

The arithmetic itself lives in a third-party dependency that is NOT vendored in
``/root/reference`` and is not installed in this image: ``diffusers`` (PyPI, version unpinned by
the reference; API usage implies ~0.30).  This file restates its published algorithm
(``AutoencoderKL`` -> ``Encoder`` -> ``DownEncoderBlock2D`` / ``UNetMidBlock2D`` /
``ResnetBlock2D`` / ``Downsample2D`` / ``Attention`` and ``DiagonalGaussianDistribution``) from
SURVEY.md Appendix A.  **Parity for the encoder is therefore UNPINNED** against the reference
(the reference ships no tests or golden vectors); the only offline cross-check is the parameter
count (34 163 664).  The PGD step and the losses ARE pinned against the reference's own code,
see ``oracle/pgd_oracle.py`` and ``oracle/gen_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  State-dict keys follow diffusers so a real checkpoint loads as is.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class EncoderConfig:
    """SURVEY Appendix A.1 (SD-1.5 ``sd-vae-ft-mse``; the SDXL VAE is architecturally identical)."""
    in_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.18215
    mid_block_add_attention: bool = True


class ResnetBlock2D(nn.Module):
    # h = conv1(silu(GN(x))); h = conv2(silu(GN(h))); return shortcut(x) + h
    def __init__(self, cin: int, cout: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1, 1, 0) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    # conv3x3 stride 2 pad 0 on F.pad(x, (0,1,0,1)): pad right & bottom only
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 2, 0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0.0))


class Attention(nn.Module):
    # single head, head_dim = C, GroupNorm (no SiLU) before q/k/v, residual connection
    def __init__(self, c: int, groups: int, eps: float):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps, affine=True)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c)])  # key "to_out.0" as in diffusers

    def forward(self, x):
        B, C, H, W = x.shape
        r = x
        t = self.group_norm(x.view(B, C, H * W)).transpose(1, 2)  # [B, HW, C]
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        s = torch.matmul(q, k.transpose(1, 2)) * (1.0 / math.sqrt(C))
        a = torch.matmul(torch.softmax(s, dim=-1), v)
        o = self.to_out[0](a)
        return o.transpose(1, 2).reshape(B, C, H, W) + r


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, n_layers, add_downsample, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, groups, eps) for i in range(n_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UNetMidBlock2D(nn.Module):
    def __init__(self, c, groups, eps, add_attention=True):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, groups, eps), ResnetBlock2D(c, c, groups, eps)])
        self.attentions = nn.ModuleList([Attention(c, groups, eps)]) if add_attention else None

    def forward(self, x):
        x = self.resnets[0](x)
        if self.attentions is not None:
            x = self.attentions[0](x)
        return self.resnets[1](x)


class Encoder(nn.Module):
    def __init__(self, cfg: EncoderConfig):
        super().__init__()
        ch = cfg.block_out_channels
        g, e = cfg.norm_num_groups, cfg.norm_eps
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, 1, 1)
        blocks = []
        cin = ch[0]
        for i, cout in enumerate(ch):
            blocks.append(DownEncoderBlock2D(cin, cout, cfg.layers_per_block, i != len(ch) - 1, g, e))
            cin = cout
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = UNetMidBlock2D(ch[-1], g, e, cfg.mid_block_add_attention)
        self.conv_norm_out = nn.GroupNorm(g, ch[-1], eps=e, affine=True)
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg.latent_channels, 3, 1, 1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    """mean, logvar = chunk(moments, 2, 1); logvar clamp(-30, 20); std = exp(.5 logvar)."""

    def __init__(self, moments: torch.Tensor):
        self.parameters = moments
        self.mean, logvar = torch.chunk(moments, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None, noise: Optional[torch.Tensor] = None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device,
                                dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution


class OracleAutoencoderKL(nn.Module):
    """encode(x) -> .latent_dist.{sample,mode}; keys ``encoder.*`` / ``quant_conv.*`` as diffusers."""

    def __init__(self, cfg: Optional[EncoderConfig] = None):
        super().__init__()
        self.cfg = cfg or EncoderConfig()
        self.encoder = Encoder(self.cfg)
        self.quant_conv = nn.Conv2d(2 * self.cfg.latent_channels, 2 * self.cfg.latent_channels, 1)

    def moments(self, x: torch.Tensor) -> torch.Tensor:
        return self.quant_conv(self.encoder(x))

    def encode(self, x: torch.Tensor) -> AutoencoderKLOutput:
        return AutoencoderKLOutput(DiagonalGaussianDistribution(self.moments(x)))


def make_oracle(seed: int = 0, cfg: Optional[EncoderConfig] = None, dtype=torch.float32) -> OracleAutoencoderKL:
    """Random-init weights exactly as SURVEY 8(d) cfg 1: ``torch.manual_seed(seed)`` then construct."""
    torch.manual_seed(seed)
    m = OracleAutoencoderKL(cfg).to(dtype)
    m.requires_grad_(False)
    m.eval()
    return m


def perturb_affine_params(model: nn.Module, seed: int = 1234) -> None:
    """Make GroupNorm gamma/beta and biases non-trivial (default init gamma=1, beta=0 hides bugs)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, nn.GroupNorm):
                mod.weight.add_(0.2 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.add_(0.1 * torch.randn(mod.bias.shape, generator=g))


def latent_loss(z: torch.Tensor, target: torch.Tensor, kind: int) -> torch.Tensor:
    """Per-image losses [B].  kind 0: ||z-t||_2 (main.py:162, per image since the reference has B=1);
    kind 1: mean((z-t)^2) (losses/losses.py:39-41 ``perturbation_loss`` = F.mse_loss)."""
    d = (z - target).reshape(z.shape[0], -1)
    if kind == 0:
        return d.norm(p=2, dim=1)
    return (d * d).mean(dim=1)


def encoder_attack_grad(model: OracleAutoencoderKL, x_adv: torch.Tensor, target: torch.Tensor,
                        noise: Optional[torch.Tensor], kind: int = 0):
    """One gradient evaluation of the VAE-encoder attack (SURVEY 8c oracle definition).

    Mirrors main.py:151-153 (clone + requires_grad), :191 (encode -> sample), :162 (loss), :176
    (autograd.grad w.r.t. the image).  Returns (grad, per-image loss, z).
    """
    with torch.enable_grad():
        x = x_adv.clone().requires_grad_(True)
        dist = model.encode(x).latent_dist
        z = dist.mode() if noise is None else dist.sample(noise=noise)
        losses = latent_loss(z, target, kind)
        (g,) = torch.autograd.grad(losses.sum(), [x])
    return g.detach(), losses.detach(), z.detach()
