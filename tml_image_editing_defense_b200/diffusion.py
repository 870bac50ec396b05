"""The reference's full attack step (main.py:144-246): encode -> add_noise -> k UNet denoising steps with
classifier-free guidance -> decode -> image-space losses -> gradient w.r.t. the image (BASELINE
configs[4]).

Encoder and decoder (forward and backward) run on this repo's sm_100a kernels through their autograd seams
(``vae.encode(x).latent_dist`` / ``vae.decode(z).sample``), the PGD update on the fused kernel.  The UNet is either
this repo's native module (``unet.py`` -> ``csrc/unet.cu``: pass ``unet_dtype=torch.float32, use_checkpointing=False``,
it recomputes its own activations in the backward) or the PyTorch library module of ``unet_torch.py`` (cuDNN / cuBLAS /
SDPA; the oracle and the library baseline), re-run under ``torch.utils.checkpoint`` per denoising step.  Either way only
the latents between steps stay alive (the reference keeps everything; checkpointing is BASELINE configs[4]'s addition).
No CLIP weights exist offline: prompt embeddings are passed in as tensors."""
from __future__ import annotations

from typing import List, Optional

import torch
from torch.utils.checkpoint import checkpoint

from .configs import TrainConfig

SCALING = 0.18215   # hard-coded in the reference for SD-1.5 and SDXL alike (main.py:191,245)


class DiffusionAttack:
    def __init__(self, cfg: TrainConfig, vae, unet, scheduler, use_checkpointing: bool = True,
                 unet_dtype: torch.dtype = torch.bfloat16):
        self.cfg, self.vae, self.unet, self.scheduler = cfg, vae, unet, scheduler
        self.use_checkpointing = use_checkpointing
        self.unet_dtype = unet_dtype
        self.device = torch.device(cfg.device)

    # ------------------------------------------------------------------ main.py:179-246
    def attack_forward(self, prompt_embeds: torch.Tensor, image: torch.Tensor, selected_noise: torch.Tensor,
                       vae_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """prompt_embeds: [2, T, D] = (negative, positive) — main.py:187 concatenates them in that order."""
        c = self.cfg
        B = image.shape[0]
        dist = self.vae.encode(image).latent_dist
        z = dist.mean if vae_noise is None else dist.mean + dist.std * vae_noise
        latents = z * SCALING                                                     # :191
        self.scheduler.set_timesteps(c.n_denoising_steps_per_iteration)           # :194
        timesteps = [int(t) for t in self.scheduler.timesteps]
        if c.limit_timesteps:
            timesteps = [t for t in timesteps if t < 700]                         # :198-199
        latents = self.scheduler.add_noise(latents, selected_noise, timesteps[:1])   # :216
        ctx = torch.cat([prompt_embeds[0:1].expand(B, -1, -1), prompt_embeds[1:2].expand(B, -1, -1)]).to(self.unet_dtype)
        extra = {"eta": c.eta} if "eta" in self.scheduler.step.__code__.co_varnames else {}   # :218-220

        def one_step(lat, t):
            inp = self.scheduler.scale_model_input(torch.cat([lat] * 2), t)       # :230-231
            with torch.autocast("cuda", dtype=self.unet_dtype, enabled=lat.is_cuda and self.unet_dtype != torch.float32):
                # (the native UNet takes the scalar timestep as a host number: no device round trip per step)
                tt = t if getattr(self.unet, "native", False) else torch.tensor(t, device=lat.device)
                pred = self.unet(inp.to(self.unet_dtype) if lat.is_cuda else inp, tt, encoder_hidden_states=ctx).sample
            pred = pred.float()
            uncond, text = pred.chunk(2)
            return uncond + c.guidance_scale * (text - uncond)                    # :240-241

        g = torch.Generator(device=latents.device).manual_seed(c.seed)
        for t in timesteps:                                                       # :229
            if self.use_checkpointing and latents.requires_grad:
                noise_pred = checkpoint(one_step, latents, t, use_reentrant=False)
            else:
                noise_pred = one_step(latents, t)
            latents = self.scheduler.step(noise_pred, t, latents, generator=g, **extra)   # :242
        return latents / SCALING                                                  # :245

    # ------------------------------------------------------------------ main.py:144-177
    def compute_grad(self, cur_image: torch.Tensor, prompt_embeds: torch.Tensor, source_image: torch.Tensor,
                     target_image: torch.Tensor, target_latent: Optional[torch.Tensor], noise: List[torch.Tensor],
                     vae_noise: Optional[torch.Tensor] = None):
        c = self.cfg
        with torch.enable_grad():
            cur = cur_image.clone()
            cur.requires_grad = True
            sel = noise[int(torch.randint(0, len(noise), (1,)))]
            output_latent = self.attack_forward(prompt_embeds, cur, sel, vae_noise)
            dec = self.vae.decode(output_latent)                                  # :156 (always decoded)
            output_image = dec.sample if hasattr(dec, "sample") else dec
            B = cur.shape[0]
            if c.apply_loss_on_images:
                rec = (output_image - target_image).reshape(B, -1).norm(p=2, dim=1)      # :160, per image
            elif c.apply_loss_on_latents:
                rec = (output_latent - target_latent).reshape(B, -1).norm(p=2, dim=1)    # :162
            else:
                raise ValueError("Please specify whether to apply loss on images or latents")
            if c.perturbation_loss_lambda > 0:
                pert = ((output_image - source_image) ** 2).reshape(B, -1).mean(dim=1)   # :168, losses.py:39-41
                loss = c.rec_loss_lambda * rec + c.perturbation_loss_lambda * pert
            else:
                pert = torch.zeros_like(rec)
                loss = c.rec_loss_lambda * rec
            (grad,) = torch.autograd.grad(loss.sum(), [cur])                      # :176
        return grad, loss.detach().mean(), output_image.detach(), {"rec_loss": rec.detach().mean(),
                                                                    "pert_loss": pert.detach().mean()}

    # ------------------------------------------------------------------ main.py:47-142
    def run(self, source_image: torch.Tensor, target_image: torch.Tensor, prompt_embeds: torch.Tensor,
            noises: Optional[List[torch.Tensor]] = None, source_mask: Optional[torch.Tensor] = None,
            vae_noise: Optional[torch.Tensor] = None, callback=None) -> torch.Tensor:
        """The reference's PGD loop around the diffusion attack (main.py:79-135): per iteration ``grad_reps`` gradient
        evaluations through the denoising loop, their mean, one ``perturbation_step`` (the fused sm_100a update).
        ``noises``: the fixed training noises of main.py:60-66 (``n_noise`` seeded N(0,1) latents when omitted)."""
        from . import ops
        c = self.cfg
        x = source_image.to(self.device, torch.float32).contiguous()
        tgt = target_image.to(self.device, torch.float32).contiguous()
        if tgt.shape[0] == 1 and x.shape[0] > 1:
            tgt = tgt.expand(x.shape[0], -1, -1, -1).contiguous()
        lat_shape = (x.shape[0], 4, x.shape[2] // 8, x.shape[3] // 8)
        if noises is None:
            g = torch.Generator(device=self.device).manual_seed(c.seed)
            noises = [torch.randn(lat_shape, generator=g, device=self.device) for _ in range(max(1, c.n_noise))]
        self.noises = noises
        x_adv = x.clone()
        self.loss_history = []
        for it in range(c.n_optimization_steps):
            grads, losses = [], []
            for _ in range(c.grad_reps):                                        # :88-99
                g_, loss, _, _ = self.compute_grad(x_adv, prompt_embeds, x, tgt, None, noises, vae_noise)
                grads.append(g_)
                losses.append(loss)
            grad = grads[0] if len(grads) == 1 else torch.stack(grads).mean(dim=0)   # :102
            mask = source_mask if c.use_segmentation_mask else None
            if c.norm_type == "linf":                                           # :248-276
                ops.pgd_step_linf_(x_adv, grad.contiguous(), x, float(c.eps), float(c.step_size), float(c.min_value),
                                   float(c.max_value))
            else:
                ops.pgd_step_l2_(x_adv, grad.contiguous(), x, mask, float(c.eps), float(c.step_size), float(c.min_value),
                                 float(c.max_value))
            self.loss_history.append(float(torch.stack(losses).mean()))
            if callback is not None:
                callback(it, self.loss_history[-1])
        return x_adv
