"""Noise schedulers used by the reference's attack loop (``scheduler.add_noise`` main.py:216,
``scheduler.step`` main.py:242, ``set_timesteps`` main.py:194, ``scale_model_input`` :231).  The reference takes
them from ``diffusers`` (pipeline default or ``LCMScheduler``, main.py:305-308); these are restatements of the
published formulas — plain tensor algebra on [B,4,h,w] latents, differentiable."""
from __future__ import annotations

from typing import Optional

import torch


class _Base:
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012):
        self.num_train_timesteps = num_train_timesteps
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)       # "scaled_linear"
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)
        self.num_inference_steps = None

    def _ac(self, t, like: torch.Tensor) -> torch.Tensor:
        t = torch.as_tensor(t).reshape(-1).long().cpu()
        return self.alphas_cumprod[t].to(like.device, like.dtype).view(-1, 1, 1, 1)

    def scale_model_input(self, sample: torch.Tensor, t) -> torch.Tensor:
        return sample

    def add_noise(self, x0: torch.Tensor, noise: torch.Tensor, t) -> torch.Tensor:
        a = self._ac(t, x0)
        return a.sqrt() * x0 + (1 - a).sqrt() * noise


class DDIMScheduler(_Base):
    """x0_hat = (x_t - sqrt(1-a_t) e)/sqrt(a_t); sigma = eta*sqrt((1-a_p)/(1-a_t))*sqrt(1-a_t/a_p);
    x_prev = sqrt(a_p) x0_hat + sqrt(1-a_p-sigma^2) e + sigma*noise.  SD-1.5 config: steps_offset 1,
    set_alpha_to_one False, 'leading' spacing."""

    def __init__(self, steps_offset: int = 1, **kw):
        super().__init__(**kw)
        self.steps_offset = steps_offset
        self.final_alpha_cumprod = self.alphas_cumprod[0]

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        self.timesteps = (torch.arange(n) * ratio).flip(0) + self.steps_offset

    def step(self, model_output, t, sample, eta: float = 0.0, generator: Optional[torch.Generator] = None,
             variance_noise: Optional[torch.Tensor] = None):
        t = int(t)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t].to(sample.device, sample.dtype)
        a_p = (self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod).to(sample.device, sample.dtype)
        x0 = (sample - (1 - a_t).sqrt() * model_output) / a_t.sqrt()
        var = (1 - a_p) / (1 - a_t) * (1 - a_t / a_p)
        sigma = eta * var.clamp(min=0).sqrt()
        prev = a_p.sqrt() * x0 + (1 - a_p - sigma ** 2).clamp(min=0).sqrt() * model_output
        if eta > 0:
            if variance_noise is None:
                variance_noise = torch.randn(sample.shape, generator=generator, device=sample.device, dtype=sample.dtype)
            prev = prev + sigma * variance_noise
        return prev


class LCMScheduler(_Base):
    """Latent-consistency sampling (main.py:305-308, run_all.py:59): s = 10 t, c_skip = .25/(s^2+.25),
    c_out = s/sqrt(s^2+.25); denoised = c_out*x0_hat + c_skip*x_t; not last: re-noise to the next timestep."""

    def __init__(self, original_inference_steps: int = 50, timestep_scaling: float = 10.0, sigma_data: float = 0.5, **kw):
        super().__init__(**kw)
        self.original_inference_steps = original_inference_steps
        self.timestep_scaling = timestep_scaling
        self.sigma_data = sigma_data
        self._i = 0

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        k = self.num_train_timesteps // self.original_inference_steps
        origin = torch.arange(1, self.original_inference_steps + 1) * k - 1
        skip = self.original_inference_steps // n
        self.timesteps = origin.flip(0)[::skip][:n]
        self._i = 0

    def step(self, model_output, t, sample, generator: Optional[torch.Generator] = None,
             variance_noise: Optional[torch.Tensor] = None):
        t = int(t)
        idx = int((self.timesteps == t).nonzero()[0])
        last = idx == len(self.timesteps) - 1
        a_t = self.alphas_cumprod[t].to(sample.device, sample.dtype)
        s = self.timestep_scaling * t
        c_skip = self.sigma_data ** 2 / (s ** 2 + self.sigma_data ** 2)
        c_out = s / (s ** 2 + self.sigma_data ** 2) ** 0.5
        x0 = (sample - (1 - a_t).sqrt() * model_output) / a_t.sqrt()
        den = c_out * x0 + c_skip * sample
        if last:
            return den
        a_p = self.alphas_cumprod[int(self.timesteps[idx + 1])].to(sample.device, sample.dtype)
        if variance_noise is None:
            variance_noise = torch.randn(sample.shape, generator=generator, device=sample.device, dtype=sample.dtype)
        return a_p.sqrt() * den + (1 - a_p).sqrt() * variance_noise
