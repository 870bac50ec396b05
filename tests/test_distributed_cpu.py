"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: image sharding with no data-path
collective, the one all-reduce of the universal-perturbation mode, identical replicas.  The compute
callables are the oracle's (the B200 kernels need a GPU; the -m gpu tests cover those)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tiny_model():
    from oracle.encoder_oracle import EncoderConfig, make_oracle
    return make_oracle(0, EncoderConfig(block_out_channels=(32, 32), layers_per_block=1, norm_num_groups=8))


def _grad_fn(model, kind=0):
    from oracle.encoder_oracle import encoder_attack_grad
    return lambda x, t, n: encoder_attack_grad(model, x, t, n, kind)[0]


def _worker_universal(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle.pgd_oracle import universal_project, universal_update
    from tml_image_editing_defense_b200.configs import UniversalConfig
    from tml_image_editing_defense_b200.dataset import SyntheticImageDataset
    from tml_image_editing_defense_b200.universal import UniversalTrainer
    # eps large enough that the image-range re-projection (apply_image_pertubation, default True) binds
    cfg = UniversalConfig(grad_reps=1, eps=0.5, step_size=0.4, resolution=16)
    assert cfg.apply_image_pertubation
    model = _tiny_model()
    ut = UniversalTrainer(cfg, _grad_fn(model), lambda g: g.sum(0, keepdim=True), lambda x, d: x + d,
                          lambda d, g: universal_update(d, g, None, cfg.eps, cfg.step_size), universal_project)
    n = 5
    ds = SyntheticImageDataset(n, resolution=16, seed=3)
    idx = ut.local_indices(n)
    imgs = ds.batch(idx)
    tg = torch.stack([torch.randn((4, 8, 8), generator=torch.Generator().manual_seed(50 + i)) for i in idx])
    delta = torch.zeros(1, 3, 16, 16)
    for _ in range(3):
        delta = ut.step(delta, imgs, tg, None, n_global=n, micro_batch=2)
        assert ut.check_replicas_identical(delta)
    if rank == 0:
        ret["delta"] = delta.clone()
    dist.destroy_process_group()


def _single_process_universal(project=True):
    from oracle.pgd_oracle import universal_project, universal_update
    from tml_image_editing_defense_b200.dataset import SyntheticImageDataset
    model = _tiny_model()
    gf = _grad_fn(model)
    n = 5
    ds = SyntheticImageDataset(n, resolution=16, seed=3)
    imgs = ds.batch(list(range(n)))
    tg = torch.stack([torch.randn((4, 8, 8), generator=torch.Generator().manual_seed(50 + i)) for i in range(n)])
    delta = torch.zeros(1, 3, 16, 16)
    for _ in range(3):
        g = gf(imgs + delta, tg, None).sum(0, keepdim=True) / n
        delta = universal_update(delta, g, None, 0.5, 0.4)
        if project:
            delta = universal_project(delta, imgs)     # the reference statement, image after image (:183-185)
    if project:
        assert float((imgs + delta).abs().max()) <= 1.0 + 1e-6
    return delta


def test_universal_allreduce_two_ranks_matches_single_process():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_universal, args=(2, port, ret), nprocs=2, join=True)
        d2 = ret["delta"]
    d1 = _single_process_universal()
    # the all-reduce changes the summation order: equal up to fp32 rounding of the gradient sum
    torch.testing.assert_close(d2, d1, rtol=0, atol=2e-6)
    assert float(d2.abs().max()) > 0
    assert not torch.equal(d1, _single_process_universal(project=False))   # the re-projection did bind somewhere


def _worker_sharded(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle.pgd_oracle import encoder_attack
    from tml_image_editing_defense_b200.dataset import SyntheticImageDataset
    from tml_image_editing_defense_b200.universal import ShardedPGD
    model = _tiny_model()
    n = 5
    ds = SyntheticImageDataset(n, resolution=16, seed=7)

    def targets(idx):
        return torch.stack([torch.randn((4, 8, 8), generator=torch.Generator().manual_seed(90 + i)) for i in idx])

    sp = ShardedPGD(lambda x, t: encoder_attack(model, x, t, None, 3, 0.1, 0.03, -1.0, 1.0))
    idx, out = sp.run(ds.batch, targets, n)
    full = sp.gather(idx, out, n, (3, 16, 16), "cpu")
    if rank == 0:
        ret["full"] = full.clone()
        ret["idx"] = list(idx)
    dist.destroy_process_group()


def test_sharded_pgd_two_ranks_bit_identical_to_single_process():
    """Per-image PGD has no data-path collective: the sharded result must equal the 1-process result bit for bit."""
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_sharded, args=(2, port, ret), nprocs=2, join=True)
        full = ret["full"]
        assert ret["idx"] == [0, 2, 4]
    from oracle.pgd_oracle import encoder_attack
    from tml_image_editing_defense_b200.dataset import SyntheticImageDataset
    model = _tiny_model()
    ds = SyntheticImageDataset(5, resolution=16, seed=7)
    for i in range(5):
        t = torch.randn((4, 8, 8), generator=torch.Generator().manual_seed(90 + i))[None]
        ref = encoder_attack(model, ds.batch([i]), t, None, 3, 0.1, 0.03, -1.0, 1.0)
        assert torch.equal(full[i], ref[0]), f"image {i} differs"
