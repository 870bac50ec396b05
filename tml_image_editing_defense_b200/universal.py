"""Universal-perturbation training (old/train_noise.py:94-185) sharded over GPUs.

One shared ``delta`` [1,3,H,W]; each rank owns the images ``rank::world_size``; per step every rank
sums d loss_i / d delta over its images, the sums are all-reduced (the only collective on the whole
hot path: 3*H*W fp32 = 3.1 MB at 512^2, NCCL over NVLink/NVSwitch), divided by the global image
count, and every rank applies the identical update (L2-normalised step, +-eps clamp, optional
image-range projection, old/train_noise.py:173-185).

The compute callables are injected so the host logic (sharding, reduction, identical replicas) is
testable on CPU with the gloo backend and the oracle; the default callables are the B200 kernels.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .configs import UniversalConfig
from .dataset import shard_indices


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class UniversalTrainer:
    def __init__(self, cfg: UniversalConfig,
                 grad_fn: Callable[[torch.Tensor, torch.Tensor, Optional[torch.Tensor]], torch.Tensor],
                 sum_fn: Callable[[torch.Tensor], torch.Tensor],
                 add_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                 step_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor]):
        """grad_fn(x_perturbed, target_latent, noise) -> d sum_i loss_i / d x  [b,3,H,W]
        sum_fn(g) -> [1,3,H,W] sum over the batch in image order
        add_fn(x, delta) -> x + delta ; step_fn(delta, grad) -> updated delta (in place allowed)."""
        self.cfg = cfg
        self.grad_fn, self.sum_fn, self.add_fn, self.step_fn = grad_fn, sum_fn, add_fn, step_fn
        self.rank, self.world = _world()

    @classmethod
    def for_b200(cls, cfg: UniversalConfig, vae) -> "UniversalTrainer":
        from . import ops

        def grad_fn(x, tgt, noise):
            g, _, _ = vae.attack_grad(x, tgt, noise, kind=cfg.loss_kind)
            return g

        def step_fn(delta, grad):
            return ops.universal_step_(delta, grad, None, float(cfg.eps), float(cfg.step_size))

        return cls(cfg, grad_fn, ops.batch_sum, ops.add_delta, step_fn)

    def local_indices(self, n_images: int) -> List[int]:
        return shard_indices(n_images, self.rank, self.world)

    def step(self, delta: torch.Tensor, images: torch.Tensor, targets: torch.Tensor,
             noises: Optional[torch.Tensor], n_global: int, micro_batch: int = 16) -> torch.Tensor:
        """One update of the shared delta from this rank's shard (`images` = local shard)."""
        total = torch.zeros_like(delta)
        for _ in range(self.cfg.grad_reps):                       # old/train_noise.py:130
            for s in range(0, images.shape[0], micro_batch):
                e = min(images.shape[0], s + micro_batch)
                xp = self.add_fn(images[s:e], delta)              # :132
                g = self.grad_fn(xp, targets[s:e], None if noises is None else noises[s:e])
                total += self.sum_fn(g)
        if self.world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM)          # the one exchange step
        total /= float(n_global * self.cfg.grad_reps)             # :166 mean over reps (and images)
        return self.step_fn(delta, total)

    def check_replicas_identical(self, delta: torch.Tensor) -> bool:
        if self.world == 1:
            return True
        lo, hi = delta.clone(), delta.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return bool(torch.equal(lo, hi))


class ShardedPGD:
    """Per-image PGD over a dataset split ``rank::world_size`` — no collective on the data path;
    results are gathered once at the end (reference run_all.py:14-21)."""

    def __init__(self, run_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor]):
        self.run_fn = run_fn
        self.rank, self.world = _world()

    def run(self, dataset_batch_fn: Callable[[Sequence[int]], torch.Tensor],
            target_fn: Callable[[Sequence[int]], torch.Tensor], n_images: int):
        idx = shard_indices(n_images, self.rank, self.world)
        out = self.run_fn(dataset_batch_fn(idx), target_fn(idx)) if idx else None
        return idx, out

    def gather(self, idx: List[int], x_adv: Optional[torch.Tensor], n_images: int, shape, device) -> Optional[torch.Tensor]:
        """All ranks receive the full [n_images, ...] result in dataset order."""
        if self.world == 1:
            return x_adv
        per = (n_images + self.world - 1) // self.world
        buf = torch.zeros((per,) + tuple(shape), dtype=torch.float32, device=device)
        if x_adv is not None:
            buf[: x_adv.shape[0]] = x_adv
        parts = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(parts, buf)
        full = torch.zeros((n_images,) + tuple(shape), dtype=torch.float32, device=device)
        for r in range(self.world):
            ids = shard_indices(n_images, r, self.world)
            if ids:
                full[ids] = parts[r][: len(ids)]
        return full
