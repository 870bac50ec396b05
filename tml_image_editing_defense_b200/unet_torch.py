"""SD-1.5 ``UNet2DConditionModel`` as plain ``torch.nn`` — the LIBRARY path of the diffusion attack.

SURVEY 8(f) n2, step 1: "start by running the UNet in PyTorch (bf16, cuDNN/SDPA) between the custom encoder
and the custom update".  The reference calls ``self.pipeline.unet(latent_model_input, t,
encoder_hidden_states=prompt_embeds).sample`` (main.py:233-238); the module itself lives in ``diffusers``
(absent here).  This file restates its published architecture with diffusers' state-dict key names, so a real
checkpoint loads with ``load_state_dict``; its only offline cross-check is the parameter count
(859 520 964 for SD-1.5).  It is NOT one of this repo's sm_100a kernels: convolutions / linears / attention go to
cuDNN / cuBLAS / SDPA through PyTorch, exactly as they do in the reference.  Replacing it is future work.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    attention_head_dim: int = 8          # SD-1.5: this is the NUMBER of heads (diffusers naming quirk)
    norm_num_groups: int = 32
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    up_has_attn: Tuple[bool, ...] = (False, True, True, True)


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers ``Timesteps(flip_sin_to_cos=True, downscale_freq_shift=0)``: [cos | sin] of t * 10000^(-i/half)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, T, C = x.shape
        q = self.to_q(x).view(B, T, self.heads, C // self.heads).transpose(1, 2)
        k = self.to_k(ctx).view(B, ctx.shape[1], self.heads, C // self.heads).transpose(1, 2)
        v = self.to_v(ctx).view(B, ctx.shape[1], self.heads, C // self.heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        return self.to_out[0](o.transpose(1, 2).reshape(B, T, C))


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, ctx_dim, heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    def __init__(self, dim, ctx_dim, heads, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Conv2d(dim, dim, 1)       # SD-1.5: use_linear_projection = False
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, ctx_dim, heads)])
        self.proj_out = nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        r = x
        h = self.proj_in(self.norm(x)).permute(0, 2, 3, 1).reshape(B, H * W, C)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
        return self.proj_out(h) + r


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 2, 1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 1, 1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, cfg: UNetConfig, has_attn, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, cfg.norm_num_groups)
                                      for i in range(cfg.layers_per_block)])
        self.attentions = nn.ModuleList([Transformer2DModel(cout, cfg.cross_attention_dim, cfg.attention_head_dim,
                                                            cfg.norm_num_groups)
                                         for _ in range(cfg.layers_per_block)]) if has_attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, c, temb, cfg: UNetConfig):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, cfg.norm_num_groups) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(c, cfg.cross_attention_dim, cfg.attention_head_dim,
                                                            cfg.norm_num_groups)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cin_prev, cout, skip_channels, temb, cfg: UNetConfig, has_attn, add_up):
        super().__init__()
        n = cfg.layers_per_block + 1
        self.resnets = nn.ModuleList()
        for i in range(n):
            rin = cin_prev if i == 0 else cout
            self.resnets.append(ResnetBlock2D(rin + skip_channels[i], cout, temb, cfg.norm_num_groups))
        self.attentions = nn.ModuleList([Transformer2DModel(cout, cfg.cross_attention_dim, cfg.attention_head_dim,
                                                            cfg.norm_num_groups) for _ in range(n)]) if has_attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, ctx):
        for i, r in enumerate(self.resnets):
            x = r(torch.cat([x, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


@dataclass
class UNetOutput:
    sample: torch.Tensor


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: Optional[UNetConfig] = None):
        super().__init__()
        self.cfg = cfg = cfg or UNetConfig()
        ch = cfg.block_out_channels
        temb = ch[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, 1, 1)
        self.time_embedding = TimestepEmbedding(ch[0], temb)
        self.down_blocks = nn.ModuleList()
        cin = ch[0]
        skip = [ch[0]]
        for i, cout in enumerate(ch):
            last = i == len(ch) - 1
            self.down_blocks.append(DownBlock(cin, cout, temb, cfg, cfg.down_has_attn[i], not last))
            skip += [cout] * cfg.layers_per_block + ([] if last else [cout])
            cin = cout
        self.mid_block = MidBlock(ch[-1], temb, cfg)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        prev = ch[-1]
        for i, cout in enumerate(rev):
            last = i == len(rev) - 1
            sk = [skip.pop() for _ in range(cfg.layers_per_block + 1)]
            self.up_blocks.append(UpBlock(prev, cout, sk, temb, cfg, cfg.up_has_attn[i], not last))
            prev = cout
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=1e-5)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, 1, 1)

    def forward(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor, **kwargs) -> UNetOutput:
        t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], device=sample.device)
        t = t.reshape(-1).expand(sample.shape[0])
        temb = self.time_embedding(timestep_embedding(t, self.cfg.block_out_channels[0]).to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states)
            skips += outs
        x = self.mid_block(x, temb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, temb, encoder_hidden_states)
        return UNetOutput(self.conv_out(F.silu(self.conv_norm_out(x))))


def tiny_unet_config() -> UNetConfig:
    """A few-million-parameter configuration with the same topology (tests)."""
    return UNetConfig(block_out_channels=(32, 64, 64, 64), cross_attention_dim=32, attention_head_dim=2,
                      norm_num_groups=8)
