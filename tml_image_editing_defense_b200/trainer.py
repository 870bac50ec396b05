"""``Trainer``: the reference's PGD driver (main.py:25-276) for the VAE-encoder attack.

Method names, argument names and return shapes follow the reference so call sites read the same:

  * ``run()``                main.py:47-142   loop: grad_reps x compute_grad -> mean -> perturbation_step
  * ``compute_grad(...)``    main.py:144-177  -> (grad, loss: float, output_image, {'rec_loss','pert_loss'})
  * ``attack_forward(...)``  main.py:179-246  encoder line :191 (the UNet loop lives in diffusion.py::DiffusionAttack)
  * ``perturbation_step()``  main.py:248-276  linf / l2

All arithmetic on images and latents runs in the CUDA kernels behind the C ABI; this file is
control flow only.  Batches of B > 1 images are supported (the reference is B = 1): losses are per
image and the scalar returned is their mean.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from . import ops
from .configs import TrainConfig
from .vae import AutoencoderKL


class Trainer:
    def __init__(self, cfg: TrainConfig, vae: AutoencoderKL, use_sdxl: bool = False, use_lcm: bool = False,
                 micro_batch: int = 16, num_streams: int = 1):
        self.cfg = cfg
        self.vae = vae
        self.use_sdxl = use_sdxl
        self.use_lcm = use_lcm
        self.device = torch.device(cfg.device)
        self.dtype = torch.float32  # main.py:33
        self.micro_batch = micro_batch
        # Images are independent, so consecutive micro-batches run on alternating CUDA streams: the
        # HBM-bound GroupNorm / SiLU passes of one overlap the tensor-core-bound GEMMs of the other.
        self._streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, num_streams))] \
            if num_streams > 1 else []
        self.noises: Optional[List[torch.Tensor]] = None
        self._noise_shape = None
        self._rng = torch.Generator(device="cpu").manual_seed(cfg.seed)
        self.loss_history: List[float] = []

    # ------------------------------------------------------------------ fixed noises (main.py:40-45)
    def _ensure_noises(self, latent_shape):
        if not self.cfg.use_fixed_noise:
            return
        if self.noises is None or self._noise_shape != tuple(latent_shape):
            g = torch.Generator(device="cpu").manual_seed(self.cfg.seed + 1)
            self.noises = [torch.randn(tuple(latent_shape), generator=g).to(self.device)
                           for _ in range(self.cfg.n_noise)]
            self._noise_shape = tuple(latent_shape)

    def _pick_noise(self, noise, latent_shape):
        if noise is None:
            return torch.randn(tuple(latent_shape), device=self.device, dtype=self.dtype)
        idx = int(torch.randint(0, len(noise), (1,), generator=self._rng))  # main.py:215 (np.random there)
        return noise[idx]

    # ------------------------------------------------------------------ main.py:179-246
    def attack_forward(self, prompt, image: torch.Tensor, noise: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        """Encoder part of attack_forward: vae.encode(image).latent_dist.sample() (unscaled latent;
        the reference multiplies by 0.18215 before the UNet and divides again after, main.py:191,245)."""
        m = self.vae.moments(image)
        mean, logvar = torch.chunk(m, 2, dim=1)
        eps = self._pick_noise(noise, mean.shape)
        return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * eps

    # ------------------------------------------------------------------ main.py:144-177
    def compute_grad(self, cur_image: torch.Tensor, prompt, source_image: torch.Tensor, target_image,
                     target_latent: torch.Tensor, noise: Optional[List[torch.Tensor]] = None,
                     grad_out: Optional[torch.Tensor] = None, beta: float = 0.0):
        if self.cfg.apply_loss_on_images:
            return self._compute_grad_images(cur_image, source_image, target_image, noise, grad_out, beta)
        if not self.cfg.apply_loss_on_latents:
            raise ValueError("Please specify whether to apply loss on images or latents")  # main.py:164
        B = cur_image.shape[0]
        lat_shape = (B, 4, cur_image.shape[2] // 8, cur_image.shape[3] // 8)
        eps = self._pick_noise(noise, lat_shape)
        grad = grad_out if grad_out is not None else torch.empty_like(cur_image)
        losses = torch.empty(B, dtype=torch.float32, device=self.device)
        mb = self.micro_batch
        chunks = [(s, min(B, s + mb)) for s in range(0, B, mb)]
        if len(chunks) == 1 or not self._streams:
            for s, e in chunks:
                _, l, _ = self.vae.attack_grad(cur_image[s:e], target_latent[s:e], eps[s:e], kind=self.cfg.loss_kind,
                                               grad_out=grad[s:e], beta=beta, grad_scale=self.cfg.rec_loss_lambda)
                losses[s:e] = l
        else:
            main = torch.cuda.current_stream(self.device)
            ready = torch.cuda.Event()
            ready.record(main)
            used = self._streams[:min(len(self._streams), len(chunks))]
            for st in used:
                st.wait_event(ready)
            for i, (s, e) in enumerate(chunks):
                with torch.cuda.stream(used[i % len(used)]):
                    _, l, _ = self.vae.attack_grad(cur_image[s:e], target_latent[s:e], eps[s:e],
                                                   kind=self.cfg.loss_kind, grad_out=grad[s:e], beta=beta,
                                                   grad_scale=self.cfg.rec_loss_lambda)
                    losses[s:e] = l
            for st in used:
                main.wait_stream(st)
        rec = losses.mean()
        loss_dict = {"rec_loss": rec, "pert_loss": 0.0, "per_image": losses}
        return grad, rec * self.cfg.rec_loss_lambda, None, loss_dict

    def _compute_grad_images(self, cur_image, source_image, target_image, noise, grad_out, beta):
        """The reference's default: rec_loss on decoded images (main.py:156-160) + perturbation_loss (:167-169)."""
        if target_image is None:
            raise ValueError("apply_loss_on_images needs target_image")
        B = cur_image.shape[0]
        lat_shape = (B, 4, cur_image.shape[2] // 8, cur_image.shape[3] // 8)
        eps = self._pick_noise(noise, lat_shape)
        grad = grad_out if grad_out is not None else torch.empty_like(cur_image)
        rec = torch.empty(B, dtype=torch.float32, device=self.device)
        pert = torch.empty(B, dtype=torch.float32, device=self.device)
        outs = []
        mb = self.micro_batch
        c = self.cfg
        for s in range(0, B, mb):
            e = min(B, s + mb)
            tgt = target_image[s:e] if target_image.shape[0] == B else target_image.expand(e - s, -1, -1, -1).contiguous()
            src = source_image[s:e] if c.perturbation_loss_lambda > 0 else None
            _, r_, p_, img = self.vae.attack_grad_images(cur_image[s:e], tgt, src, eps[s:e], c.rec_loss_lambda,
                                                         c.perturbation_loss_lambda, grad_out=grad[s:e], beta=beta)
            rec[s:e] = r_
            pert[s:e] = p_
            outs.append(img)
        loss = c.rec_loss_lambda * rec.mean() + c.perturbation_loss_lambda * pert.mean()
        return grad, loss, torch.cat(outs), {"rec_loss": rec.mean(), "pert_loss": pert.mean(), "per_image": rec}

    # ------------------------------------------------------------------ main.py:248-276
    def perturbation_step(self, X_adv: torch.Tensor, grad: torch.Tensor, X: torch.Tensor,
                          X_mask: Optional[torch.Tensor] = None, grad_scale: float = 1.0) -> torch.Tensor:
        c = self.cfg
        if c.norm_type == "l2":
            return ops.pgd_step_l2_(X_adv, grad, X, X_mask, float(c.eps), float(c.step_size), float(c.min_value),
                                    float(c.max_value))
        elif c.norm_type == "linf":
            return ops.pgd_step_linf_(X_adv, grad, X, float(c.eps), float(c.step_size), float(c.min_value),
                                      float(c.max_value))
        raise ValueError(c.norm_type)

    # ------------------------------------------------------------------ main.py:47-142
    def run(self, source_image: torch.Tensor, target_image: Optional[torch.Tensor] = None,
            target_latent: Optional[torch.Tensor] = None, source_mask: Optional[torch.Tensor] = None,
            callback: Optional[Callable[[int, float], None]] = None, sync_every: int = 25) -> torch.Tensor:
        """PGD on a batch of images; returns X_adv in the image range [min_value, max_value]."""
        c = self.cfg
        source_image = source_image.to(self.device, self.dtype).contiguous()
        B = source_image.shape[0]
        if target_image is not None:
            target_image = target_image.to(self.device, self.dtype).contiguous()
        if target_latent is None:
            if target_image is None:
                raise ValueError("need target_image or target_latent")
            with torch.no_grad():  # main.py:75
                target_latent = self.vae.encode(target_image.to(self.device, self.dtype)).latent_dist.sample()
        target_latent = target_latent.to(self.device, self.dtype).contiguous()
        if target_latent.shape[0] == 1 and B > 1:
            target_latent = target_latent.expand(B, -1, -1, -1).contiguous()
        self._ensure_noises(target_latent.shape)
        X_adv = source_image.clone()
        grad = torch.empty_like(X_adv)
        self.loss_history = []
        pending = []
        for iteration in range(c.n_optimization_steps):
            step_losses = []
            for i in range(c.grad_reps):  # main.py:88-99; accumulation happens in the dgrad kernel (beta=1)
                _, loss, _, _ = self.compute_grad(cur_image=X_adv, prompt=None, source_image=source_image,
                                                  target_image=target_image, target_latent=target_latent, noise=self.noises,
                                                  grad_out=grad, beta=0.0 if i == 0 else 1.0)
                step_losses.append(loss)
            # main.py:102 takes the mean over reps; sign() and the L2 normalisation are invariant to
            # that positive factor, so the sum is passed on as is.
            X_adv = self.perturbation_step(X_adv=X_adv, grad=grad, X=source_image,
                                           X_mask=source_mask if c.use_segmentation_mask else None)
            pending.append(torch.stack(step_losses).mean())
            if (iteration + 1) % sync_every == 0 or iteration == c.n_optimization_steps - 1:
                vals = torch.stack(pending).tolist()      # one host sync per `sync_every` steps
                self.loss_history.extend(vals)
                pending = []
                if callback is not None:
                    callback(iteration, vals[-1])
        return X_adv

    @staticmethod
    def to_pil(X_adv: torch.Tensor):
        """main.py:139-140."""
        import torchvision.transforms as T
        img = (X_adv / 2 + 0.5).clamp(0, 1)
        return [T.ToPILImage()(im.cpu()).convert("RGB") for im in img]
