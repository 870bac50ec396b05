"""tml_image_editing_defense_b200 — B200-native (sm_100a) PGD image-immunization hot path.

Public surface mirrors the reference (OrLichter/tml_image_editing_defense): ``TrainConfig``,
``Trainer`` (``run`` / ``compute_grad`` / ``perturbation_step``), the ``vae.encode(...).latent_dist``
seam (``AutoencoderKL``), the loss names, ``ImagePromptDataset``.  Compute lives in
``csrc/libtml_b200.so`` behind the C ABI of ``include/tml_b200.h``.
"""
from .configs import TrainConfig, UniversalConfig  # noqa: F401
from .dataset import ImagePromptDataset, SyntheticImageDataset, shard_indices  # noqa: F401
from . import losses  # noqa: F401

__all__ = ["TrainConfig", "UniversalConfig", "ImagePromptDataset", "SyntheticImageDataset", "shard_indices",
           "losses", "AutoencoderKL", "EncoderConfig", "Trainer", "UniversalTrainer", "ShardedPGD", "ops"]


def __getattr__(name):  # lazy: these load the CUDA library
    if name in ("AutoencoderKL", "EncoderConfig", "SD15_VAE", "SDXL_VAE"):
        from . import vae
        return getattr(vae, name)
    if name == "Trainer":
        from .trainer import Trainer
        return Trainer
    if name in ("UniversalTrainer", "ShardedPGD"):
        from . import universal
        return getattr(universal, name)
    if name == "ops":
        import importlib
        return importlib.import_module(".ops", __name__)
    raise AttributeError(name)
