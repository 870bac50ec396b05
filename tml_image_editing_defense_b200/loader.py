"""Host pipeline of the PGD loop: a rank's shard of the dataset, batch by batch, staged in pinned host memory and
copied to the GPU one batch ahead of the consumer.

The reference feeds its loop from ``DataLoader(ImagePromptDataset(...), batch_size=1, shuffle=True, num_workers=0)``
(old/train_noise.py:100-101) or from a hand-split list of file names (run_all.py:14-21).  Here rank r owns the images
``r::world`` (``dataset.shard_indices``), a batch is assembled in a pinned buffer (two buffers, reused) and its
host-to-device copy is issued on a side stream while the previous batch is still being attacked; the consumer's
stream waits on the copy's event, never on the host.  On a machine without CUDA (tests) the same iteration runs
without pinning and without streams.
"""
from __future__ import annotations

from typing import Callable, Iterator, List, Optional, Sequence, Tuple

import torch

from .dataset import shard_indices


class ShardedImageLoader:
    def __init__(self, dataset, batch_size: int, rank: int = 0, world_size: int = 1, device: str = "cuda:0",
                 shuffle: bool = False, seed: int = 0, fetch: Optional[Callable[[Sequence[int]], torch.Tensor]] = None):
        """dataset[i] -> (image [3,H,W] in [-1,1], prompt) as ImagePromptDataset / SyntheticImageDataset give it.
        ``fetch(indices) -> [b,3,H,W]`` overrides the per-item path (e.g. SyntheticImageDataset.batch)."""
        if batch_size < 1:
            raise ValueError("batch_size must be >= 1")
        self.dataset, self.batch_size = dataset, batch_size
        self.rank, self.world = rank, world_size
        self.device = torch.device(device)
        self.indices: List[int] = shard_indices(len(dataset), rank, world_size)
        if shuffle:   # same permutation on every rank would break the partition: shuffle inside the shard only
            g = torch.Generator().manual_seed(seed * 1_000_003 + rank)
            self.indices = [self.indices[i] for i in torch.randperm(len(self.indices), generator=g).tolist()]
        self.fetch = fetch
        self.cuda = self.device.type == "cuda" and torch.cuda.is_available()
        self._pinned: List[Optional[torch.Tensor]] = [None, None]
        self._stream = torch.cuda.Stream(device=self.device) if self.cuda else None

    def __len__(self) -> int:
        return (len(self.indices) + self.batch_size - 1) // self.batch_size

    def batches(self) -> List[List[int]]:
        return [self.indices[s:s + self.batch_size] for s in range(0, len(self.indices), self.batch_size)]

    def _assemble(self, idx: Sequence[int], slot: int) -> Tuple[torch.Tensor, List[str]]:
        if self.fetch is not None:
            imgs, prompts = self.fetch(idx), [getattr(self.dataset, "default_prompt", "")] * len(idx)
        else:
            items = [self.dataset[i] for i in idx]
            imgs, prompts = torch.stack([it[0] for it in items]), [it[1] for it in items]
        imgs = imgs.to(torch.float32)
        if not self.cuda:
            return imgs, prompts
        buf = self._pinned[slot]
        if buf is None or buf.shape[1:] != imgs.shape[1:] or buf.shape[0] < imgs.shape[0]:
            buf = torch.empty((self.batch_size,) + tuple(imgs.shape[1:]), dtype=torch.float32).pin_memory()
            self._pinned[slot] = buf
        buf[: imgs.shape[0]].copy_(imgs)
        return buf[: imgs.shape[0]], prompts

    def _upload(self, host: torch.Tensor):
        if not self.cuda:
            return host.to(self.device), None
        with torch.cuda.stream(self._stream):
            dev = host.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._stream)
        return dev, ev

    def __iter__(self) -> Iterator[Tuple[List[int], torch.Tensor, List[str]]]:
        """Yields (dataset indices, images on the device, prompts); batch k+1 is in flight while batch k is used."""
        bl = self.batches()
        nxt = None
        free_ev: List[Optional[object]] = [None, None]     # "the upload that read pinned slot s has finished"
        for k in range(len(bl) + 1):
            cur = nxt
            if k < len(bl):
                slot = k & 1
                if free_ev[slot] is not None:
                    free_ev[slot].synchronize()             # the pinned buffer is reused: its last copy must be done
                host, prompts = self._assemble(bl[k], slot)
                dev, ev = self._upload(host)
                free_ev[slot] = ev
                nxt = (bl[k], dev, prompts, ev)
            if cur is not None:
                idx, dev, prompts, ev = cur
                if ev is not None:
                    torch.cuda.current_stream(self.device).wait_event(ev)
                    dev.record_stream(torch.cuda.current_stream(self.device))
                yield idx, dev, prompts
