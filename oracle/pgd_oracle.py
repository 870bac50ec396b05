"""ORACLE (test infrastructure, not product code).

CPU restatement of the reference's PGD update rules, loss functions and universal-perturbation
update.  Pinned against the reference's own code by ``oracle/gen_golden.py`` (which executes
``Trainer.perturbation_step`` from ``/root/reference/main.py`` and the update statements of
``/root/reference/old/train_noise.py`` and stores their outputs under ``tests/golden/``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# PGD steps  (reference: main.py:248-276; second witness old/yuval_playground.py:250-259,309-313)
# ----------------------------------------------------------------------------------------------
def pgd_step_linf(x_adv: torch.Tensor, grad: torch.Tensor, x: torch.Tensor, eps: float, step: float,
                  lo: float, hi: float) -> torch.Tensor:
    """main.py:272-274: sign step, L-inf eps-ball projection around X, clamp to [lo, hi]."""
    x_adv = x_adv - grad.detach().sign() * step
    x_adv = torch.minimum(torch.maximum(x_adv, x - eps), x + eps)
    return torch.clamp(x_adv, lo, hi)


def pgd_step_l2(x_adv: torch.Tensor, grad: torch.Tensor, x: torch.Tensor, mask: Optional[torch.Tensor],
                eps: float, step: float, lo: float, hi: float) -> torch.Tensor:
    """main.py:254-268: per-sample normalised step, optional mask (x3 channels), renorm, clamp."""
    nd = x.dim() - 1
    gnorm = torch.norm(grad.detach().reshape(grad.shape[0], -1), dim=1).view(-1, *([1] * nd))
    gn = grad.detach() / (gnorm + 1e-10)
    if mask is not None:
        gn = gn * mask.repeat(1, 3, 1, 1)
    x_adv = x_adv - gn * step
    d = x_adv - x.detach()
    d = torch.renorm(d, p=2, dim=0, maxnorm=eps)
    return torch.clamp(x + d, lo, hi)


def pgd_step_linf_numpy(x_adv: np.ndarray, grad: np.ndarray, x: np.ndarray, eps: float, step: float,
                        lo: float, hi: float) -> np.ndarray:
    """Independent fp32 numpy restatement of main.py:272-274 with ATen's special-value semantics:
    sign(+-0)=0, sign(NaN)=0, minimum/maximum/clamp propagate NaN.  Used to cross-check the torch
    restatement above bit for bit."""
    f = np.float32
    x_adv, grad, x = x_adv.astype(f), grad.astype(f), x.astype(f)
    with np.errstate(invalid="ignore"):
        sgn = (grad > 0).astype(f) - (grad < 0).astype(f)          # NaN -> 0, +-0 -> 0
        v = (x_adv - sgn * f(step)).astype(f)
        lo_b = (x - f(eps)).astype(f)
        hi_b = (x + f(eps)).astype(f)
        v = np.maximum(v, lo_b)                                     # np.maximum propagates NaN
        v = np.minimum(v, hi_b)
        nan = np.isnan(v)
        v = np.minimum(np.maximum(v, f(lo)), f(hi))
        v[nan] = np.nan
    return v.astype(f)


# ----------------------------------------------------------------------------------------------
# Universal perturbation update (reference: old/train_noise.py:173-185)
# ----------------------------------------------------------------------------------------------
def universal_update(delta: torch.Tensor, grad: torch.Tensor, source: Optional[torch.Tensor], eps: float,
                     step: float, lo: float = -1.0, hi: float = 1.0) -> torch.Tensor:
    """L2-normalised gradient step (:173-177), clamp to +-eps (:180), and — when a source image is
    given (``apply_image_pertubation``) — re-projection so that source+delta stays in [lo, hi] (:183-185)."""
    nd = grad.dim() - 1
    gnorm = torch.norm(grad.view(grad.shape[0], -1), dim=1).view(-1, *([1] * nd))
    gn = grad / (gnorm + 1e-10)
    delta = delta - gn * step
    delta = torch.clamp(delta, -eps, eps)
    if source is not None:
        delta = torch.clamp(source + delta, lo, hi) - source
    return delta


def universal_project(delta: torch.Tensor, sources: torch.Tensor, lo: float = -1.0, hi: float = 1.0) -> torch.Tensor:
    """old/train_noise.py:183-185 applied once per source image, in order:
    ``perturbed = clamp(source + perturbation, lo, hi); perturbation = perturbed - source``."""
    for s in sources:
        delta = torch.clamp(s[None] + delta, lo, hi) - s[None]
    return delta


# ----------------------------------------------------------------------------------------------
# Losses (reference: losses/losses.py:6-41)
# ----------------------------------------------------------------------------------------------
def perturbation_loss(adv_image: torch.Tensor, source_image: torch.Tensor) -> torch.Tensor:
    return F.mse_loss(adv_image, source_image)


def lp_distance(x: torch.Tensor, y: torch.Tensor, p) -> torch.Tensor:
    return torch.norm(x - y, p)


def lp_regularization(params, p) -> torch.Tensor:
    if isinstance(params, torch.Tensor):
        params = [params]
    return sum(torch.norm(q, p) for q in params)


def cosine_similarity_plus_one(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return (F.cosine_similarity(x, y) + 1).mean()


# ----------------------------------------------------------------------------------------------
# Full encoder-attack loop: the CPU baseline / --impl reference workload
# ----------------------------------------------------------------------------------------------
def encoder_attack(model, x: torch.Tensor, target: torch.Tensor, noise: Optional[torch.Tensor], steps: int,
                   eps: float, step: float, lo: float, hi: float, kind: int = 0, norm_type: str = "linf",
                   grad_reps: int = 1, record=None) -> torch.Tensor:
    """PGD loop of main.py:79-115 with the UNet removed (PhotoGuard encoder attack)."""
    from .encoder_oracle import encoder_attack_grad
    x_adv = x.clone()
    for _ in range(steps):
        grads, losses = [], []
        for _ in range(grad_reps):
            g, l, _ = encoder_attack_grad(model, x_adv, target, noise, kind)
            grads.append(g)
            losses.append(l)
        g = torch.stack(grads).mean(0)                      # main.py:102
        if record is not None:
            record(x_adv, g, torch.stack(losses).mean(0))
        if norm_type == "linf":
            x_adv = pgd_step_linf(x_adv, g, x, eps, step, lo, hi)
        else:
            x_adv = pgd_step_l2(x_adv, g, x, None, eps, step, lo, hi)
    return x_adv
