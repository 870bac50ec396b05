"""Image sources for the PGD loop.  ``ImagePromptDataset`` mirrors data/dataset.py:7-43 (rglob
``*.jpg`` -> Resize(512, bilinear) -> CenterCrop(512) -> ToTensor -> Normalize(.5,.5), values in
[-1,1]); ``SyntheticImageDataset`` is the seeded stand-in the benchmarks use (no network, no
dataset on the GPU box).  ``shard_indices`` is the multi-GPU split: independent images, no
communication (the reference's manual halving of the image list, run_all.py:14-21)."""
from __future__ import annotations

from pathlib import Path
from typing import List, Sequence

import torch
from torch.utils.data import Dataset


def shard_indices(n: int, rank: int, world_size: int) -> List[int]:
    """Images rank::world_size."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n, world_size))


def shard_counts(n: int, world_size: int) -> List[int]:
    return [len(range(r, n, world_size)) for r in range(world_size)]


class ImagePromptDataset(Dataset):
    def __init__(self, image_dir: str, default_prompt: str, resolution: int = 512):
        from PIL import Image
        self.images = []
        self.default_prompt = default_prompt
        self.resolution = resolution
        self.image_transforms = self.get_image_transforms(resolution)
        for image_path in sorted(Path(image_dir).rglob("*.jpg")):
            self.images.append(Image.open(image_path).convert("RGB"))

    @staticmethod
    def get_image_transforms(resolution: int = 512):
        from torchvision import transforms
        return transforms.Compose([
            transforms.Resize(resolution, interpolation=transforms.InterpolationMode.BILINEAR),
            transforms.CenterCrop(resolution),
            transforms.ToTensor(),
            transforms.Normalize([0.5], [0.5]),
        ])

    @staticmethod
    def get_image_transform_no_normalization(resolution: int = 512):
        from torchvision import transforms
        return transforms.Compose([
            transforms.Resize(resolution, interpolation=transforms.InterpolationMode.BILINEAR),
            transforms.CenterCrop(resolution),
            transforms.ToTensor(),
        ])

    def __len__(self):
        return len(self.images)

    def __getitem__(self, idx):
        return self.image_transforms(self.images[idx]), self.default_prompt


class SyntheticImageDataset(Dataset):
    """U[-1,1) images, reproducible per index (same image whatever the shard layout)."""

    def __init__(self, n: int, resolution: int = 512, seed: int = 0, default_prompt: str = ""):
        self.n, self.resolution, self.seed, self.default_prompt = n, resolution, seed, default_prompt

    def __len__(self):
        return self.n

    def image(self, idx: int) -> torch.Tensor:
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + idx)
        return torch.rand((3, self.resolution, self.resolution), generator=g) * 2 - 1

    def __getitem__(self, idx):
        return self.image(idx), self.default_prompt

    def batch(self, indices: Sequence[int]) -> torch.Tensor:
        return torch.stack([self.image(i) for i in indices])
