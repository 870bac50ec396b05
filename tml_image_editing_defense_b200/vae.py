"""Drop-in for the ``AutoencoderKL`` object the reference calls (``pipeline.vae``).

Seams mirrored (file:line in /root/reference):
  * ``vae.encode(x).latent_dist.sample()`` / ``.mode()``            main.py:75,191; old/train_noise.py:133
  * ``retrieve_latents`` (``latent_dist`` attribute)                 pipelines/pipeline_stable_diffusion_img2img.py:77-87
  * ``vae.config.scaling_factor`` / ``.block_out_channels``          old/train_noise.py:133; pipeline :307,:758
  * ``vae.dtype``, ``vae.to(...)``, ``vae.requires_grad_(False)``    main.py:290,302; old/train_noise.py:89
  * differentiable w.r.t. the image under ``torch.autograd.grad``    main.py:176

The encoder forward and its input-gradient backward run in the sm_100a kernels behind the C ABI
(``include/tml_b200.h``); so do ``vae.decode`` (main.py:156) and its gradient w.r.t. the latent when the
state dict carries the decoder ("decoder.*", "post_quant_conv.*") -- without those keys ``decode`` raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch

from . import _lib


@dataclass
class EncoderConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.18215
    mid_block_add_attention: bool = True
    force_upcast: bool = False


SD15_VAE = EncoderConfig(scaling_factor=0.18215)
SDXL_VAE = EncoderConfig(scaling_factor=0.13025)

_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


class DiagonalGaussianDistribution:
    """Same surface as diffusers' class: ``mean``, ``logvar``, ``std``, ``var``, ``sample``, ``mode``."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution


class _EncodeFn(torch.autograd.Function):
    """moments = encoder(x); backward = input gradient only (the reference never needs weight grads)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, vae: "AutoencoderKL"):
        need_grad = x.requires_grad
        moments, saved = vae._forward_raw(x, keep=need_grad)
        ctx.vae = vae
        ctx.saved_buf = saved
        ctx.shape = tuple(x.shape)
        return moments

    @staticmethod
    def backward(ctx, dmoments: torch.Tensor):
        dx = ctx.vae._backward_raw(dmoments.contiguous().float(), ctx.saved_buf, ctx.shape)
        ctx.saved_buf = None
        return dx, None


class _DecodeFn(torch.autograd.Function):
    """image = decoder(post_quant_conv(z)); backward = gradient w.r.t. z only."""

    @staticmethod
    def forward(ctx, z: torch.Tensor, vae: "AutoencoderKL"):
        image, saved = vae._decode_raw(z, keep=z.requires_grad)
        ctx.vae = vae
        ctx.saved_buf = saved
        ctx.shape = tuple(z.shape)
        return image

    @staticmethod
    def backward(ctx, dimage: torch.Tensor):
        dz = ctx.vae._decode_backward_raw(dimage.contiguous().float(), ctx.saved_buf, ctx.shape)
        ctx.saved_buf = None
        return dz, None


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class AutoencoderKL:
    """B200-native encoder behind the reference's ``vae`` interface."""

    def __init__(self, config: Optional[EncoderConfig] = None, device: str = "cuda:0"):
        self.config = config or EncoderConfig()
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None and torch.cuda.is_available():
            self.device = torch.device("cuda", torch.cuda.current_device())   # a concrete index: streams are per device
        self.dtype = torch.float32          # image / moments dtype at the seam (reference runs fp32, main.py:33)
        self.compute_dtype = torch.bfloat16  # activations and weights inside the kernels; fp32 accumulate
        self._lib = _lib.load()
        if self.device.type != "cuda":
            raise _lib.TmlError("AutoencoderKL (B200) needs a CUDA device; there is no CPU path")
        cfg = _lib.TmlEncoderCfg()
        cfg.in_channels = self.config.in_channels
        cfg.latent_channels = self.config.latent_channels
        cfg.num_blocks = len(self.config.block_out_channels)
        for i, c in enumerate(self.config.block_out_channels):
            cfg.block_out_channels[i] = c
        cfg.layers_per_block = self.config.layers_per_block
        cfg.norm_num_groups = self.config.norm_num_groups
        cfg.norm_eps = self.config.norm_eps
        cfg.mid_block_add_attention = int(self.config.mid_block_add_attention)
        h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(self._lib.tml_encoder_create(C.byref(cfg), idx, C.byref(h)))
        self._h = h
        # scratch is per CUDA stream: independent micro-batches may run concurrently on several
        # streams (Trainer), each with its own workspace / saved-state buffer
        self._ws: Dict[int, torch.Tensor] = {}
        self._scratch_saved: Dict[int, torch.Tensor] = {}
        self._scratch_saved_dec: Dict[int, torch.Tensor] = {}
        self._finalized = False
        self.has_decoder = False

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False):
        """Accepts a diffusers AutoencoderKL state dict.  The decoder ("decoder.*", "post_quant_conv.*") is
        optional: without it ``decode`` raises."""
        for k, v in sd.items():
            if not k.startswith(("encoder.", "quant_conv.", "decoder.", "post_quant_conv.")):
                continue
            t = v.detach()
            if t.dtype not in _DTYPES:
                t = t.float()
            t = t.contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self._lib.tml_encoder_set_weight(self._h, k.encode(), t.data_ptr(), _DTYPES[t.dtype], shape,
                                                        t.dim()))
        with torch.cuda.device(self.device):
            _lib.check(self._lib.tml_encoder_finalize(self._h, None))
        self._finalized = True
        self.has_decoder = any(k.startswith("decoder.") for k in sd)
        return self

    @classmethod
    def from_state_dict(cls, sd, config: Optional[EncoderConfig] = None, device: str = "cuda:0") -> "AutoencoderKL":
        return cls(config, device).load_state_dict(sd)

    # ------------------------------------------------------------------ nn.Module-like surface
    def to(self, *args, **kwargs):
        return self

    def requires_grad_(self, flag: bool = False):
        return self

    def eval(self):
        return self

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.tml_encoder_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ raw entry points
    def _buffers(self, B: int, H: int, W: int, decoder: bool = False):
        """(workspace tensor of the current stream, bytes of saved state).  H, W: image size (encoder) or
        latent size (decoder).  Encoder and decoder share the per-stream workspace (stream-ordered use)."""
        ws_b, sv_b = C.c_size_t(), C.c_size_t()
        q = self._lib.tml_decoder_query if decoder else self._lib.tml_encoder_query
        _lib.check(q(self._h, B, H, W, C.byref(ws_b), C.byref(sv_b)))
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < ws_b.value:
            self._ws.pop(key, None)
            ws = None
            ws = torch.empty(ws_b.value, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws, sv_b.value

    def _forward_raw(self, x: torch.Tensor, keep: bool, saved: Optional[torch.Tensor] = None):
        if not self._finalized:
            raise _lib.TmlError("load_state_dict() must be called before encode()")
        if not x.is_cuda:
            raise _lib.TmlError("encode() needs a CUDA tensor; the B200 path has no CPU fallback")
        x = x.detach().to(torch.float32).contiguous()
        B, Cc, H, W = x.shape
        ws, sv_bytes = self._buffers(B, H, W)
        if saved is None:
            if keep:
                saved = torch.empty(sv_bytes, dtype=torch.uint8, device=self.device)
            else:
                key = torch.cuda.current_stream(self.device).cuda_stream
                saved = self._scratch_saved.get(key)
                if saved is None or saved.numel() < sv_bytes:
                    self._scratch_saved.pop(key, None)
                    saved = None
                    saved = torch.empty(sv_bytes, dtype=torch.uint8, device=self.device)
                    self._scratch_saved[key] = saved
        L2 = 2 * self.config.latent_channels
        f = 2 ** (len(self.config.block_out_channels) - 1)
        moments = torch.empty((B, L2, H // f, W // f), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.tml_encoder_forward(self._h, x.data_ptr(), B, H, W, moments.data_ptr(), saved.data_ptr(),
                                                 ws.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        return moments, saved

    def _backward_raw(self, dmoments: torch.Tensor, saved: torch.Tensor, shape, out: Optional[torch.Tensor] = None,
                      beta: float = 0.0) -> torch.Tensor:
        B, Cc, H, W = shape
        ws, _ = self._buffers(B, H, W)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
            beta = 0.0
        _lib.check(self._lib.tml_encoder_backward(self._h, dmoments.data_ptr(), B, H, W, saved.data_ptr(),
                                                  out.data_ptr(), beta, ws.data_ptr(),
                                                  torch.cuda.current_stream(self.device).cuda_stream))
        return out

    # ------------------------------------------------------------------ reference-facing API
    def moments(self, x: torch.Tensor) -> torch.Tensor:
        """quant_conv(encoder(x)): [B, 8, H/8, W/8] fp32, differentiable w.r.t. x."""
        if torch.is_grad_enabled() and x.requires_grad:
            return _EncodeFn.apply(x, self)
        return self._forward_raw(x, keep=False)[0]

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        out = AutoencoderKLOutput(DiagonalGaussianDistribution(self.moments(x)))
        return out if return_dict else (out.latent_dist,)

    # ------------------------------------------------------------------ decoder (main.py:156)
    def _decode_raw(self, z: torch.Tensor, keep: bool):
        if not self.has_decoder:
            raise _lib.TmlError("decoder weights were not loaded (state dict had no 'decoder.*' keys)")
        if not z.is_cuda:
            raise _lib.TmlError("decode() needs a CUDA tensor; the B200 path has no CPU fallback")
        z = z.detach().to(torch.float32).contiguous()
        B, Cz, h, w = z.shape
        ws, sv_bytes = self._buffers(B, h, w, decoder=True)
        if keep:
            saved = torch.empty(sv_bytes, dtype=torch.uint8, device=self.device)
        else:
            key = torch.cuda.current_stream(self.device).cuda_stream
            saved = self._scratch_saved_dec.get(key)
            if saved is None or saved.numel() < sv_bytes:
                self._scratch_saved_dec.pop(key, None)
                saved = None
                saved = torch.empty(sv_bytes, dtype=torch.uint8, device=self.device)
                self._scratch_saved_dec[key] = saved
        f = 2 ** (len(self.config.block_out_channels) - 1)
        image = torch.empty((B, 3, h * f, w * f), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.tml_decoder_forward(self._h, z.data_ptr(), B, h, w, image.data_ptr(), saved.data_ptr(),
                                                 ws.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        return image, saved

    def _decode_backward_raw(self, dimage: torch.Tensor, saved: torch.Tensor, zshape) -> torch.Tensor:
        B, Cz, h, w = zshape
        ws, _ = self._buffers(B, h, w, decoder=True)
        dz = torch.empty(zshape, dtype=torch.float32, device=self.device)
        _lib.check(self._lib.tml_decoder_backward(self._h, dimage.data_ptr(), B, h, w, saved.data_ptr(), dz.data_ptr(),
                                                  ws.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        return dz

    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        """``vae.decode(z).sample`` (main.py:156), differentiable w.r.t. z."""
        if torch.is_grad_enabled() and z.requires_grad:
            img = _DecodeFn.apply(z, self)
        else:
            img = self._decode_raw(z, keep=False)[0]
        return DecoderOutput(sample=img) if return_dict else (img,)

    def attack_grad_images(self, x_adv: torch.Tensor, target_image: torch.Tensor, source_image: Optional[torch.Tensor],
                           noise: Optional[torch.Tensor], rec_lambda: float = 1.0, pert_lambda: float = 1.0,
                           grad_out: Optional[torch.Tensor] = None, beta: float = 0.0):
        """compute_grad of the reference with its default image-space losses (main.py:144-177, UNet removed):
        encode -> sample -> decode -> rec_lambda*||out-target||_2 + pert_lambda*mse(out, source) -> backward
        through decoder and encoder.  Returns (grad, rec [B], pert [B], output_image)."""
        from . import ops
        moments, saved_e = self._forward_raw(x_adv, keep=False)
        z = ops.posterior_sample(moments, noise)
        img, saved_d = self._decode_raw(z, keep=False)
        rec, pert, dimg = ops.image_loss(img, target_image, source_image, rec_lambda, pert_lambda)
        dz = self._decode_backward_raw(dimg, saved_d, tuple(z.shape))
        dm = ops.posterior_sample_backward(moments, noise, dz)
        g = self._backward_raw(dm, saved_e, tuple(x_adv.shape), out=grad_out, beta=beta)
        return g, rec, pert, img

    # ------------------------------------------------------------------ fused attack gradient (no autograd)
    def attack_grad(self, x_adv: torch.Tensor, target: torch.Tensor, noise: Optional[torch.Tensor], kind: int = 0,
                    grad_out: Optional[torch.Tensor] = None, beta: float = 0.0, grad_scale: float = 1.0):
        """One gradient evaluation of the encoder attack: encoder fwd -> sample + latent loss +
        dmoments (one kernel) -> encoder input-gradient bwd.  Mirrors compute_grad (main.py:144-177)
        with the UNet removed.  Returns (grad [B,3,H,W], per-image loss [B], z)."""
        from . import ops
        moments, saved = self._forward_raw(x_adv, keep=False)
        z, loss, dm = ops.latent_loss(moments, noise, target, kind, grad_scale, need_grad=True)
        g = self._backward_raw(dm, saved, tuple(x_adv.shape), out=grad_out, beta=beta)
        return g, loss, z
