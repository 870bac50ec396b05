"""Universal-perturbation training (old/train_noise.py:94-185) sharded over GPUs.

One shared ``delta`` [1,3,H,W]; each rank owns the images ``rank::world_size``; per step every rank
sums d loss_i / d delta over its images, the sums are all-reduced (the only collective on the whole
hot path: 3*H*W fp32 = 3.1 MB at 512^2, NCCL over NVLink/NVSwitch), divided by the global image
count, and every rank applies the identical update (L2-normalised step, +-eps clamp, old/train_noise.py:173-180).

``cfg.apply_image_pertubation`` (:182-185, default True): the reference re-projects the perturbation so that the one
image of its step stays in [-1, 1]: ``perturbation = clamp(source + perturbation, -1, 1) - source``.  A sharded step
sees every image, so the same statement is applied against the two extreme images that bound them all -- the
per-pixel minimum and maximum over the whole dataset (``prepare_projection``: computed once, MIN/MAX all-reduced, so
every replica clamps to the identical interval [-1 - min_i x_i, 1 - max_i x_i], which is what applying the
reference statement image after image converges to since every per-image interval contains 0).

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
                 step_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                 project_fn: Optional[Callable[[torch.Tensor, torch.Tensor], torch.Tensor]] = None):
        """grad_fn(x_perturbed, target_latent, noise) -> d sum_i loss_i / d x  [b,3,H,W]
        sum_fn(g) -> [1,3,H,W] sum over the batch in image order
        add_fn(x, delta) -> x + delta ; step_fn(delta, grad) -> updated delta (in place allowed)
        project_fn(delta, sources [n,3,H,W]) -> delta re-projected against each source in order (:183-185)."""
        self.cfg = cfg
        self.grad_fn, self.sum_fn, self.add_fn, self.step_fn = grad_fn, sum_fn, add_fn, step_fn
        self.project_fn = project_fn
        self.rank, self.world = _world()
        self._bounds: Optional[torch.Tensor] = None     # [2,3,H,W]: per-pixel (min, max) image of the whole dataset
        # set to a list to have (start, end) CUDA events recorded around every all-reduce on the current stream
        # (bench.py --mode universal); the interval includes the wait for the slowest rank
        self.comm_events: Optional[list] = None

    @classmethod
    def for_b200(cls, cfg: UniversalConfig, vae) -> "UniversalTrainer":
        from . import ops

        def grad_fn(x, tgt, noise):
            g, _, _ = vae.attack_grad(x, tgt, noise, kind=cfg.loss_kind)
            return g

        def step_fn(delta, grad):
            return ops.universal_step_(delta, grad, None, float(cfg.eps), float(cfg.step_size))

        return cls(cfg, grad_fn, ops.batch_sum, ops.add_delta, step_fn, ops.universal_project_)

    def prepare_projection(self, images: torch.Tensor) -> None:
        """Per-pixel (min, max) over every rank's images: the bounds of old/train_noise.py:183-185 for the whole
        dataset.  Called once (the images do not change); a no-op when ``apply_image_pertubation`` is off."""
        if not self.cfg.apply_image_pertubation:
            self._bounds = None
            return
        if self.project_fn is None:
            raise ValueError("apply_image_pertubation=True needs a project_fn (old/train_noise.py:182-185)")
        if images.shape[0]:
            lo, hi = images.amin(dim=0), images.amax(dim=0)
        else:   # an empty shard must not constrain the others
            shape = images.shape[1:]
            lo = torch.full(shape, float("inf"), dtype=images.dtype, device=images.device)
            hi = torch.full(shape, float("-inf"), dtype=images.dtype, device=images.device)
        if self.world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        self._bounds = torch.stack([lo, hi]).contiguous()

    def local_indices(self, n_images: int) -> List[int]:
        return shard_indices(n_images, self.rank, self.world)

    def step(self, delta: torch.Tensor, images: torch.Tensor, targets: torch.Tensor,
             noises: Optional[torch.Tensor], n_global: int, micro_batch: int = 16) -> torch.Tensor:
        """One update of the shared delta from this rank's shard (`images` = local shard)."""
        if self.cfg.apply_image_pertubation and self._bounds is None:
            self.prepare_projection(images)
        total = torch.zeros_like(delta)
        for _ in range(self.cfg.grad_reps):                       # old/train_noise.py:130
            for s in range(0, images.shape[0], micro_batch):
                e = min(images.shape[0], s + micro_batch)
                xp = self.add_fn(images[s:e], delta)              # :132
                g = self.grad_fn(xp, targets[s:e], None if noises is None else noises[s:e])
                total += self.sum_fn(g)
        if self.world > 1:
            ev = None
            if self.comm_events is not None and total.is_cuda:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            dist.all_reduce(total, op=dist.ReduceOp.SUM)          # the one exchange step
            if ev is not None:
                ev[1].record()
                self.comm_events.append(ev)
        total /= float(n_global * self.cfg.grad_reps)             # :166 mean over reps (and images)
        delta = self.step_fn(delta, total)                        # :169-180
        if self.cfg.apply_image_pertubation and self._bounds is not None:
            delta = self.project_fn(delta, self._bounds)          # :182-185
        return delta

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
