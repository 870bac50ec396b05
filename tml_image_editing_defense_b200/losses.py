"""Loss names of the reference (losses/losses.py:6-41).  These small host-side reductions run in
plain torch on whatever device their inputs live on; the latent loss that sits on the PGD hot path
is fused with the posterior sample and its gradient in ``tml_latent_loss`` (see ops.latent_loss)."""
from __future__ import annotations

from typing import List, Union

import torch
import torch.nn.functional as F


class LpRegularization:
    def __init__(self, p):
        self.p = p

    def __call__(self, regularization_parameters: Union[List[torch.Tensor], torch.Tensor]) -> torch.Tensor:
        if isinstance(regularization_parameters, torch.Tensor):
            regularization_parameters = [regularization_parameters]
        return sum(torch.norm(q, self.p) for q in regularization_parameters)


class LpDistance:
    def __init__(self, p):
        self.p = p

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return torch.norm(x - y, self.p)


class CosineSimilarity:
    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return (F.cosine_similarity(x, y) + 1).mean()


def perturbation_loss(adv_image: torch.Tensor, source_image: torch.Tensor) -> torch.Tensor:
    return F.mse_loss(adv_image, source_image)
