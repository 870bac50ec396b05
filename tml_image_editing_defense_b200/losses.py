"""Host-side loss helpers under the names the reference exports from ``losses/losses.py`` (:6-41):
``LpRegularization``, ``LpDistance``, ``CosineSimilarity``, ``perturbation_loss``.

They are thin wrappers over ``torch.linalg.vector_norm`` / ``torch.nn.functional`` and run on whatever device
their inputs live on; results are bit-identical to the reference's (checked against golden vectors produced by
the reference module itself, tests/test_host_logic.py).  The loss that sits on the PGD hot path is not computed
here: it is fused with the posterior sample and its gradient in ``tml_latent_loss`` / ``tml_image_loss``
(see ``ops.latent_loss`` / ``ops.image_loss``).
"""
from __future__ import annotations

from typing import Iterable, Union

import torch
import torch.nn.functional as F

Tensors = Union[torch.Tensor, Iterable[torch.Tensor]]


def _p_norm(t: torch.Tensor, p) -> torch.Tensor:
    # same reduction torch.norm(t, p) dispatches to (flattened vector norm)
    return torch.linalg.vector_norm(t.reshape(-1), ord=p)


class _WithOrder:
    __slots__ = ("p",)

    def __init__(self, p):
        self.p = p


class LpDistance(_WithOrder):
    """``||x - y||_p`` over all elements (reference: old/train_noise.py:109-111,153-154)."""

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return _p_norm(x - y, self.p)


class LpRegularization(_WithOrder):
    """Sum of the p-norms of one tensor or a collection of tensors."""

    def __call__(self, regularization_parameters: Tensors) -> torch.Tensor:
        items = [regularization_parameters] if torch.is_tensor(regularization_parameters) \
            else list(regularization_parameters)
        total = 0
        for t in items:
            total = total + _p_norm(t, self.p)
        return total


class CosineSimilarity:
    """Mean over the batch of ``1 + cos(x, y)`` along dim 1 (always >= 0)."""

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return (1 + F.cosine_similarity(x, y)).mean()


def perturbation_loss(adv_image: torch.Tensor, source_image: torch.Tensor) -> torch.Tensor:
    """Mean squared error between the (decoded) adversarial image and the source (main.py:168)."""
    return F.mse_loss(adv_image, source_image, reduction="mean")
