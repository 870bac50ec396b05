"""SD-1.5 ``UNet2DConditionModel`` on this repo's sm_100a kernels, behind the call the reference makes:
``self.pipeline.unet(latent_model_input, t, encoder_hidden_states=prompt_embeds).sample`` (main.py:233-238).

Forward and the gradient w.r.t. the sample (``torch.autograd.grad(loss, [cur_image])`` through the denoising loop,
main.py:176,229-243) run in ``csrc/unet.cu`` through the C ABI (``tml_unet_*``).  Timestep, prompt embeddings and
weights are constants of the attack, so no other gradient exists.  By default the activations of every call whose sample
requires a gradient stay in HBM until its backward (as in the reference, which keeps everything; ~10 GB per call at 16
samples of 64 x 64 latents -- the attention probabilities are recomputed, not kept); ``keep_activations=False`` re-runs
the forward in the backward instead (activation checkpointing per UNet call, BASELINE configs[4]), so only the sample, the
timestep and the prompt embeddings stay alive between the denoising steps.

There is no fallback: without the CUDA library or an sm_100a device construction raises.  ``unet_torch.py`` is the
PyTorch restatement of the same module (the oracle of the tests and the library baseline), not a code path of this one.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib
from .unet_torch import UNetConfig

_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


@dataclass
class UNet2DConditionOutput:
    sample: torch.Tensor


class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sample: torch.Tensor, unet: "UNet2DConditionModel", t: float, ctx_emb: torch.Tensor):
        keep = sample.requires_grad and unet.keep_activations
        out, saved = unet._forward_raw(sample, t, ctx_emb, keep=keep)
        ctx.unet, ctx.t, ctx.ctx_emb = unet, t, ctx_emb
        ctx.saved_buf = saved if keep else None
        ctx.save_for_backward(sample)
        return out

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        (sample,) = ctx.saved_tensors
        unet = ctx.unet
        saved = ctx.saved_buf
        if saved is None:   # checkpointing: recompute the activations of this call
            _, saved = unet._forward_raw(sample, ctx.t, ctx.ctx_emb, keep=False)
        dx = unet._backward_raw(dout.contiguous().float(), saved, tuple(sample.shape), ctx.ctx_emb.shape[1])
        ctx.saved_buf = None
        return dx, None, None, None


class UNet2DConditionModel:
    """B200-native denoiser with the reference-facing surface of diffusers' module (``__call__`` -> ``.sample``)."""
    native = True   # DiffusionAttack passes the timestep as a host scalar

    def __init__(self, config: Optional[UNetConfig] = None, device: str = "cuda:0", keep_activations: bool = True):
        self.config = config or UNetConfig()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.TmlError("UNet2DConditionModel (B200) needs a CUDA device; there is no CPU path")
        if self.device.index is None and torch.cuda.is_available():
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.keep_activations = keep_activations
        self.dtype = torch.float32
        self._lib = _lib.load()
        c = self.config
        cfg = _lib.TmlUnetCfg()
        cfg.in_channels, cfg.out_channels = c.in_channels, c.out_channels
        cfg.num_blocks = len(c.block_out_channels)
        for i, ch in enumerate(c.block_out_channels):
            cfg.block_out_channels[i] = ch
            cfg.down_has_attn[i] = int(c.down_has_attn[i])
            cfg.up_has_attn[i] = int(c.up_has_attn[i])
        cfg.layers_per_block = c.layers_per_block
        cfg.cross_attention_dim = c.cross_attention_dim
        cfg.num_heads = c.attention_head_dim
        cfg.norm_num_groups = c.norm_num_groups
        h = C.c_void_p()
        _lib.check(self._lib.tml_unet_create(C.byref(cfg), self.device.index or 0, C.byref(h)))
        self._h = h
        self._ws: Dict[int, torch.Tensor] = {}
        self._saved: Dict[int, torch.Tensor] = {}
        self._finalized = False

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False):
        """Accepts a diffusers ``UNet2DConditionModel`` state dict (SD-1.5 topology)."""
        for k, v in sd.items():
            t = v.detach()
            if t.dtype not in _DTYPES:
                t = t.float()
            t = t.contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self._lib.tml_unet_set_weight(self._h, k.encode(), t.data_ptr(), _DTYPES[t.dtype], shape, t.dim()))
        with torch.cuda.device(self.device):
            _lib.check(self._lib.tml_unet_finalize(self._h, None))
        self._finalized = True
        return self

    def to(self, *args, **kwargs):
        return self

    def requires_grad_(self, flag: bool = False):
        return self

    def eval(self):
        return self

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.tml_unet_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ raw entry points
    def _buffers(self, B: int, h: int, w: int, T: int):
        ws_b, sv_b = C.c_size_t(), C.c_size_t()
        _lib.check(self._lib.tml_unet_query(self._h, B, h, w, T, C.byref(ws_b), C.byref(sv_b)))
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < ws_b.value:
            self._ws.pop(key, None)
            ws = None
            ws = torch.empty(ws_b.value, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws, sv_b.value

    def _forward_raw(self, sample: torch.Tensor, t: float, ctx_emb: torch.Tensor, keep: bool):
        if not self._finalized:
            raise _lib.TmlError("load_state_dict() must be called before the UNet is used")
        if not sample.is_cuda:
            raise _lib.TmlError("the B200 UNet needs CUDA tensors; there is no CPU fallback")
        x = sample.detach().to(torch.float32).contiguous()
        ce = ctx_emb.detach().to(torch.float32).contiguous()
        B, _, h, w = x.shape
        if ce.shape[0] != B or ce.shape[2] != self.config.cross_attention_dim:
            raise ValueError(f"encoder_hidden_states {tuple(ce.shape)} does not match batch {B} / "
                             f"cross_attention_dim {self.config.cross_attention_dim}")
        T = ce.shape[1]
        ws, sv_bytes = self._buffers(B, h, w, T)
        key = torch.cuda.current_stream(self.device).cuda_stream
        if keep:
            saved = torch.empty(sv_bytes, dtype=torch.uint8, device=self.device)
        else:
            saved = self._saved.get(key)
            if saved is None or saved.numel() < sv_bytes:
                self._saved.pop(key, None)
                saved = None
                saved = torch.empty(sv_bytes, dtype=torch.uint8, device=self.device)
                self._saved[key] = saved
        out = torch.empty((B, self.config.out_channels, h, w), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.tml_unet_forward(self._h, x.data_ptr(), float(t), ce.data_ptr(), B, h, w, T, out.data_ptr(),
                                              saved.data_ptr(), ws.data_ptr(), key))
        return out, saved

    def _backward_raw(self, dout: torch.Tensor, saved: torch.Tensor, shape, T: int) -> torch.Tensor:
        B, _, h, w = shape
        ws, _ = self._buffers(B, h, w, T)
        dx = torch.empty(shape, dtype=torch.float32, device=self.device)
        _lib.check(self._lib.tml_unet_backward(self._h, dout.data_ptr(), B, h, w, T, saved.data_ptr(), dx.data_ptr(),
                                               ws.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        return dx

    def saved_tensor(self, saved: torch.Tensor, name: str, index: int = 0) -> torch.Tensor:
        """bf16 NHWC view of a kept activation (tests)."""
        off = C.c_size_t()
        dims = (C.c_int * 4)()
        _lib.check(self._lib.tml_debug_unet_saved_tensor(self._h, name.encode(), index, C.byref(off), dims))
        n = dims[0] * dims[1] * dims[2] * dims[3]
        return saved[off.value: off.value + 2 * n].view(torch.bfloat16).view(dims[0], dims[1], dims[2], dims[3])

    def count(self, what: str) -> int:
        off = C.c_size_t()
        dims = (C.c_int * 4)()
        _lib.check(self._lib.tml_debug_unet_saved_tensor(self._h, f"count_{what}".encode(), 0, C.byref(off), dims))
        return int(off.value)

    # ------------------------------------------------------------------ the reference-facing call
    def __call__(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor, **kwargs):
        t = float(timestep.reshape(-1)[0]) if torch.is_tensor(timestep) else float(timestep)
        if torch.is_tensor(timestep) and timestep.numel() > 1 and not bool((timestep.reshape(-1) == timestep.reshape(-1)[0]).all()):
            raise ValueError("the B200 UNet takes one timestep per call (as the reference passes it, main.py:233)")
        if sample.requires_grad and torch.is_grad_enabled():
            out = _UNetFn.apply(sample, self, t, encoder_hidden_states.detach())
        else:
            out, _ = self._forward_raw(sample, t, encoder_hidden_states, keep=False)
        return UNet2DConditionOutput(out)

    forward = __call__
