"""Tensor-level wrappers over the C ABI (device pointers + the current CUDA stream).

PyTorch is used for device memory and streams only; all arithmetic happens in
``csrc/libtml_b200.so``.  Every function raises if given CPU tensors: there is no CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


def _stream(t: torch.Tensor) -> int:
    """The current stream of the tensor's own device (not of whatever device happens to be current)."""
    return torch.cuda.current_stream(t.device).cuda_stream


def _on(t: torch.Tensor):
    """Pointer-only entry points launch on the current device: make it the tensor's device for the call."""
    return torch.cuda.device(t.device)


def _chk(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.TmlError(f"{name} must be a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.TmlError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.TmlError(f"{name} must be contiguous")
    return t


def pgd_step_linf_(x_adv: torch.Tensor, grad: torch.Tensor, x: torch.Tensor, eps: float, step: float, lo: float,
                   hi: float) -> torch.Tensor:
    """In-place fused L-inf PGD step (reference main.py:272-274)."""
    _chk(x_adv, "x_adv"), _chk(grad, "grad"), _chk(x, "x")
    assert x_adv.shape == grad.shape == x.shape
    lib = _lib.load()
    with _on(x_adv):
        _lib.check(lib.tml_pgd_step_linf(x_adv.data_ptr(), grad.data_ptr(), x.data_ptr(), eps, step, lo, hi,
                                         x_adv.numel(), _stream(x_adv)))
    return x_adv


def pgd_step_l2_(x_adv: torch.Tensor, grad: torch.Tensor, x: torch.Tensor, mask: Optional[torch.Tensor], eps: float,
                 step: float, lo: float, hi: float) -> torch.Tensor:
    """In-place L2 PGD step (reference main.py:254-268)."""
    _chk(x_adv, "x_adv"), _chk(grad, "grad"), _chk(x, "x")
    B, Cc = x.shape[0], x.shape[1]
    hw = x[0, 0].numel()
    if mask is not None:
        _chk(mask, "mask")
        assert mask.numel() == B * hw
    lib = _lib.load()
    ws = torch.empty(lib.tml_pgd_l2_workspace(B), dtype=torch.uint8, device=x.device)
    with _on(x_adv):
        _lib.check(lib.tml_pgd_step_l2(x_adv.data_ptr(), grad.data_ptr(), x.data_ptr(),
                                       mask.data_ptr() if mask is not None else None, eps, step, lo, hi, B, Cc, hw,
                                       ws.data_ptr(), _stream(x_adv)))
    return x_adv


def latent_loss(moments: torch.Tensor, noise: Optional[torch.Tensor], target: torch.Tensor, kind: int = 0,
                grad_scale: float = 1.0, need_grad: bool = True):
    """Posterior sample + per-image latent loss + d loss / d moments (main.py:162,191; losses.py:39-41).

    Returns (z [B,4,h,w], loss [B], dmoments [B,8,h,w] or None)."""
    _chk(moments, "moments"), _chk(target, "target")
    B, c2, h, w = moments.shape
    assert c2 == 8 and target.shape == (B, 4, h, w)
    if noise is not None:
        _chk(noise, "noise")
        assert noise.shape == target.shape
    z = torch.empty_like(target)
    loss = torch.empty(B, dtype=torch.float32, device=moments.device)
    dm = torch.empty_like(moments) if need_grad else None
    lib = _lib.load()
    with _on(moments):
        _lib.check(lib.tml_latent_loss(kind, moments.data_ptr(), noise.data_ptr() if noise is not None else None,
                                       target.data_ptr(), B, h, w, grad_scale, z.data_ptr(), loss.data_ptr(),
                                       dm.data_ptr() if dm is not None else None, _stream(moments)))
    return z, loss, dm


def add_delta(x: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
    """x[b] + delta for a shared perturbation (old/train_noise.py:132)."""
    _chk(x, "x"), _chk(delta, "delta")
    out = torch.empty_like(x)
    with _on(x):
        _lib.check(_lib.load().tml_add_delta(x.data_ptr(), delta.data_ptr(), out.data_ptr(), x.shape[0], x[0].numel(),
                                             _stream(x)))
    return out


def batch_sum(g: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """scale * sum over the batch dimension in image order (gradient of the shared delta)."""
    _chk(g, "g")
    out = torch.empty((1,) + tuple(g.shape[1:]), dtype=torch.float32, device=g.device)
    with _on(g):
        _lib.check(_lib.load().tml_batch_sum(g.data_ptr(), out.data_ptr(), g.shape[0], g[0].numel(), scale, _stream(g)))
    return out


def universal_step_(delta: torch.Tensor, grad: torch.Tensor, source: Optional[torch.Tensor], eps: float, step: float,
                    lo: float = -1.0, hi: float = 1.0) -> torch.Tensor:
    """In-place universal-perturbation update (old/train_noise.py:173-185)."""
    _chk(delta, "delta"), _chk(grad, "grad")
    if source is not None:
        _chk(source, "source")
    ws = torch.empty(4096, dtype=torch.uint8, device=delta.device)
    with _on(delta):
        _lib.check(_lib.load().tml_universal_step(delta.data_ptr(), grad.data_ptr(),
                                                  source.data_ptr() if source is not None else None, eps, step, lo, hi,
                                                  delta.numel(), ws.data_ptr(), _stream(delta)))
    return delta


def universal_project_(delta: torch.Tensor, sources: torch.Tensor, lo: float = -1.0, hi: float = 1.0) -> torch.Tensor:
    """In place, for each source image in order: delta = clamp(source + delta, lo, hi) - source
    (old/train_noise.py:183-185).  ``sources``: [n, *delta.shape[1:]]."""
    _chk(delta, "delta"), _chk(sources, "sources")
    assert sources.numel() % delta.numel() == 0
    with _on(delta):
        _lib.check(_lib.load().tml_universal_project(delta.data_ptr(), sources.data_ptr(),
                                                     sources.numel() // delta.numel(), lo, hi, delta.numel(),
                                                     _stream(delta)))
    return delta


def posterior_sample(moments: torch.Tensor, noise: Optional[torch.Tensor]) -> torch.Tensor:
    """latent_dist.sample() with explicit noise / .mode() when noise is None (main.py:191)."""
    _chk(moments, "moments")
    B, c2, h, w = moments.shape
    z = torch.empty((B, c2 // 2, h, w), dtype=torch.float32, device=moments.device)
    if noise is not None:
        _chk(noise, "noise")
    with _on(moments):
        _lib.check(_lib.load().tml_posterior_sample(moments.data_ptr(), noise.data_ptr() if noise is not None else None,
                                                    z.data_ptr(), B, h, w, _stream(moments)))
    return z


def posterior_sample_backward(moments: torch.Tensor, noise: Optional[torch.Tensor], dz: torch.Tensor) -> torch.Tensor:
    _chk(moments, "moments"), _chk(dz, "dz")
    B, c2, h, w = moments.shape
    dm = torch.empty_like(moments)
    with _on(moments):
        _lib.check(_lib.load().tml_posterior_sample_backward(moments.data_ptr(),
                                                             noise.data_ptr() if noise is not None else None,
                                                             dz.data_ptr(), dm.data_ptr(), B, h, w, _stream(moments)))
    return dm


def image_loss(out: torch.Tensor, target: torch.Tensor, source: Optional[torch.Tensor], rec_lambda: float = 1.0,
               pert_lambda: float = 1.0, need_grad: bool = True):
    """Per-image rec = ||out - target||_2 (main.py:160), pert = mse(out, source) (losses.py:39-41) and the
    gradient of rec_lambda*rec + pert_lambda*pert w.r.t. out.  Returns (rec [B], pert [B], dout or None)."""
    _chk(out, "out"), _chk(target, "target")
    assert out.shape == target.shape
    if source is not None:
        _chk(source, "source")
    B = out.shape[0]
    rec = torch.empty(B, dtype=torch.float32, device=out.device)
    pert = torch.zeros(B, dtype=torch.float32, device=out.device)
    dout = torch.empty_like(out) if need_grad else None
    lib = _lib.load()
    ws = torch.empty(lib.tml_image_loss_workspace(B), dtype=torch.uint8, device=out.device)
    with _on(out):
        _lib.check(lib.tml_image_loss(out.data_ptr(), target.data_ptr(), source.data_ptr() if source is not None else None,
                                      B, out[0].numel(), rec_lambda, pert_lambda, rec.data_ptr(), pert.data_ptr(),
                                      dout.data_ptr() if dout is not None else None, ws.data_ptr(), _stream(out)))
    return rec, pert, dout
