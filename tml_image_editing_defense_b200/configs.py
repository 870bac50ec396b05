"""Option names of the reference (configs.py:86-159 ``TrainConfig``; old/train_noise.py:20-48
``Config``) kept field for field for the PGD hot path.  Fields that only feed the UNet / wandb /
captioning parts of the reference are accepted and ignored so existing call sites keep working."""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Tuple


@dataclass
class TrainConfig:
    source_image_path: Optional[Path] = None
    target_image_path: Optional[Path] = None
    default_source_image_caption: str = ""
    output_path: Path = Path("./output")
    experiment_name: str = "experiment_l2_fixed_noise"
    n_optimization_steps: int = 200
    n_denoising_steps_per_iteration: int = 4     # UNet steps: 0 in the encoder attack
    apply_loss_on_images: bool = True            # reference default (configs.py:103): losses on vae.decode(z)
    apply_loss_on_latents: bool = False          # reference default (configs.py:105); the encoder attack sets it
    limit_timesteps: bool = True
    rec_loss_lambda: float = 1.0
    perturbation_loss_lambda: float = 1.0        # reference default (configs.py:111); image-space loss only
    seed: int = 42
    prompts: List[str] = field(default_factory=lambda: [""])
    device: str = "cuda:0"
    # optimisation parameters (configs.py:121-141)
    norm_type: str = "l2"
    eps: float = 0.1
    step_size: float = 0.006
    min_value: float = -1
    max_value: float = 1
    guidance_scale: float = 3.0
    grad_reps: int = 5
    eta: float = 0.9
    add_image_caption_to_prompts: bool = False
    use_segmentation_mask: bool = False
    use_fixed_noise: bool = True
    n_noise: int = 1
    image_visualization_interval: int = 25
    # --- additions of this build ---
    # The reference silently overrides eps / step_size / grad_reps from norm_type in __post_init__
    # (configs.py:152-159).  True reproduces that; False keeps the values given.
    override_from_norm_type: bool = True
    latent_loss: str = "l2norm"                  # "l2norm" (main.py:162) or "mse" (losses.py:39-41)
    resolution: int = 512

    def __post_init__(self):
        if self.norm_type not in ("l2", "linf"):
            raise ValueError(f"norm_type must be 'l2' or 'linf', got {self.norm_type!r}")
        if self.override_from_norm_type:
            if self.norm_type == "l2":
                self.eps, self.step_size, self.grad_reps = 32, 7.5, 10
            else:
                self.eps, self.step_size, self.grad_reps = 0.1, 0.006, 5
        if not (self.apply_loss_on_images or self.apply_loss_on_latents):
            raise ValueError("Please specify whether to apply loss on images or latents")  # main.py:164
        self.source_image = None
        self.target_image = None
        if self.source_image_path is not None and Path(self.source_image_path).exists():
            from PIL import Image
            self.source_image = Image.open(self.source_image_path).convert("RGB")
        if self.target_image_path is not None and Path(self.target_image_path).exists():
            from PIL import Image
            self.target_image = Image.open(self.target_image_path).convert("RGB")

    @property
    def loss_kind(self) -> int:
        return {"l2norm": 0, "mse": 1}[self.latent_loss]

    @classmethod
    def encoder_attack(cls, **kw) -> "TrainConfig":
        """The VAE-encoder attack of BASELINE configs[1]: loss between latents (main.py:161-162), no decode,
        no image-space perturbation loss.  The defaults of this class are the reference's (image losses)."""
        kw.setdefault("apply_loss_on_images", False)
        kw.setdefault("apply_loss_on_latents", True)
        kw.setdefault("perturbation_loss_lambda", 0.0)
        return cls(**kw)


@dataclass
class UniversalConfig:
    """old/train_noise.py:20-48 ``Config`` (fields used by the update rule)."""
    dataset_dir: Optional[str] = None
    default_prompt: str = ""
    device: str = "cuda:0"
    batch_size: int = 1
    epochs: int = 2000
    max_steps: int = 2000
    seed: int = 0
    apply_image_pertubation: bool = True          # (sic) reference spelling, old/train_noise.py:41
    grad_reps: int = 10
    eps: float = 16
    step_size: float = 1
    resolution: int = 512
    latent_loss: str = "l2norm"

    @property
    def loss_kind(self) -> int:
        return {"l2norm": 0, "mse": 1}[self.latent_loss]
