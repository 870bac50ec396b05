"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the
weight packing + tap tables reproduce conv forward / input-gradient semantics, configs, sharding."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import emulate_gemm, pack_conv3x3, rel_err

ROOT = Path(__file__).resolve().parents[1]


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def test_library_exports_every_declared_symbol():
    from tml_image_editing_defense_b200 import _lib
    lib = _lib.load()
    header = (ROOT / "include" / "tml_b200.h").read_text()
    declared = set(re.findall(r"\b(tml_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tml_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.tml_version() == 100


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tml_image_editing_defense_b200 import _lib
    from tml_image_editing_defense_b200.vae import AutoencoderKL
    with pytest.raises(_lib.TmlError):
        AutoencoderKL()
    from tml_image_editing_defense_b200 import ops
    with pytest.raises(_lib.TmlError):
        ops.pgd_step_linf_(torch.zeros(4), torch.zeros(4), torch.zeros(4), 0.1, 0.01, -1, 1)


def _conv_desc(B, H, W, Cin, N, gn, taps=9):
    """A dense 3x3 (or 1x1) stride-1 convolution descriptor with dummy (never dereferenced) pointers."""
    from tml_image_editing_defense_b200 import _lib
    d = _lib.TmlGemmDesc()
    d.A = 256; d.A_C = Cin; d.A_W = W; d.A_H = H; d.A_B = B
    d.A_sW = Cin; d.A_sH = W * Cin; d.A_sB = H * W * Cin
    d.stride = 1; d.ntaps = taps
    for t in range(taps):
        d.dh[t] = (t // 3 - 1) if taps == 9 else 0
        d.dw[t] = (t % 3 - 1) if taps == 9 else 0
    d.OW = W; d.OH = H
    d.Bm = 256; d.N = N; d.B_sN = taps * Cin; d.alpha = 1.0
    d.D = 256; d.D_sW = N; d.D_sH = W * N; d.D_sB = H * W * N; d.D_sN = 1
    d.R_sW = N; d.R_sH = W * N; d.R_sB = H * W * N
    d.gn_mode = gn
    if gn:
        d.gn_partial = 256
    if gn == 2:
        d.gn_x = 256; d.gn_ss = 256; d.gn_mr = 256; d.gn_gamma = 256
    return d


def test_partial_sum_geometry_of_the_fused_reductions():
    """Host-side tiling logic (no GPU needed): how many GroupNorm partial entries per image each kernel writes.
    Operand-swapped kernel (rows of >= 128 pixels, 128/256/512 channels): one entry per 128 pixels for the statistics
    (8 epilogue warps), per 64 pixels for the backward sums (16 warps); pixel-major kernel: one per 128-row tile; 3x3
    convolutions on rows of 64 pixels: one per 128 slots of the pitch-66 halo space (64 rows x 66 / 128 = 33)."""
    import ctypes as C
    from tml_image_editing_defense_b200 import _lib
    lib = _lib.load()
    n = lambda d: lib.tml_debug_gn_chunks_per_image(C.byref(d))
    assert n(_conv_desc(2, 8, 512, 128, 128, 1)) == 8 * 512 // 128          # swapped, single CTA
    assert n(_conv_desc(2, 8, 512, 128, 128, 2)) == 8 * 512 // 64
    assert n(_conv_desc(2, 8, 128, 512, 512, 1)) == 8 * 128 // 128          # swapped, CTA pairs (rows of 128 pixels)
    assert n(_conv_desc(2, 8, 128, 512, 512, 2)) == 8 * 128 // 64
    assert lib.tml_debug_gn_tiles_per_image(64, 64) == 32
    assert n(_conv_desc(2, 64, 64, 512, 512, 1)) == 33                      # 3x3 on rows of 64 pixels: pitch-66 halo tiles
    assert n(_conv_desc(2, 64, 64, 512, 512, 2)) == 33
    assert n(_conv_desc(2, 64, 64, 512, 512, 1, taps=1)) == 32              # 1x1 at 64^2: plain pixel-major tiles
    assert n(_conv_desc(2, 32, 32, 512, 512, 1)) == lib.tml_debug_gn_tiles_per_image(32, 32) == 8   # rows of 32 pixels
    assert n(_conv_desc(2, 7, 128, 512, 512, 1)) == lib.tml_debug_gn_tiles_per_image(7, 128)         # odd rows: no pairs
    assert n(_conv_desc(2, 16, 16, 128, 256, 1, taps=1)) == lib.tml_debug_gn_tiles_per_image(16, 16)  # 1x1: pixel-major


def test_unet_groupnorm_chunking_is_batch_independent_and_fine_grained():
    """The UNet's general GroupNorm kernels (csrc/unet_kernels.cu: gng_*) split an image into pixel chunks that depend on
    (HW, C) only -- the entry point takes no batch -- with 4 ... 16 pixels per thread: at least 64 blocks per image at the
    SD-1.5 levels that have that many pixels (a coarser split left the 8x8 / 16x16 levels at 48-176 blocks per launch)."""
    from tml_image_editing_defense_b200 import _lib
    lib = _lib.load()
    n = lib.tml_debug_unet_gn_chunks_per_image
    for hw, c in [(4096, 320), (4096, 640), (4096, 960), (1024, 640), (1024, 1280), (1024, 1920), (256, 1280), (256, 2560)]:
        pl = 256 // min(c // 8, 256)                       # pixel lanes of a 256-thread block
        chunks = n(hw, c)
        ppc = -(-hw // chunks)                             # ceil: pixels per chunk (the last chunk may be ragged)
        assert chunks >= 64, (hw, c, chunks)
        assert 4 * pl <= ppc <= 16 * pl, (hw, c, chunks, ppc, pl)
    assert n(64, 1280) == 16 and n(64, 2560) == 16         # 8x8 level: 4 pixels per thread
    assert n(0, 320) == -1 and n(64, 100) == -1            # bad shapes are rejected, not guessed


@pytest.mark.parametrize("ci,co", [(8, 16), (16, 8)])
def test_pack_forward_s1_matches_conv2d(ci, co):
    g = torch.Generator().manual_seed(0)
    w = bf16_round(torch.randn(co, ci, 3, 3, generator=g))
    x = bf16_round(torch.randn(2, ci, 8, 16, generator=g))
    mat, dh, dw = pack_conv3x3(w, 0)
    y = emulate_gemm(x.permute(0, 2, 3, 1), mat, dh, dw, 1, 8, 16).permute(0, 3, 1, 2)
    ref = F.conv2d(x.double(), w.double(), padding=1)
    assert rel_err(y, ref) < 1e-12


def test_pack_forward_s2_matches_padded_conv2d():
    g = torch.Generator().manual_seed(1)
    w = bf16_round(torch.randn(8, 8, 3, 3, generator=g))
    x = bf16_round(torch.randn(2, 8, 16, 16, generator=g))
    mat, dh, dw = pack_conv3x3(w, 2)
    y = emulate_gemm(x.permute(0, 2, 3, 1), mat, dh, dw, 2, 8, 8).permute(0, 3, 1, 2)
    ref = F.conv2d(F.pad(x.double(), (0, 1, 0, 1)), w.double(), stride=2)
    assert rel_err(y, ref) < 1e-12


def test_pack_dgrad_s1_matches_autograd():
    g = torch.Generator().manual_seed(2)
    ci, co = 8, 16
    w = bf16_round(torch.randn(co, ci, 3, 3, generator=g))
    x = torch.randn(2, ci, 8, 8, generator=g, dtype=torch.float64, requires_grad=True)
    dy = bf16_round(torch.randn(2, co, 8, 8, generator=g))
    (ref,) = torch.autograd.grad(F.conv2d(x, w.double(), padding=1), x, dy.double())
    mat, dh, dw = pack_conv3x3(w, 1)
    dx = emulate_gemm(dy.permute(0, 2, 3, 1), mat, dh, dw, 1, 8, 8).permute(0, 3, 1, 2)
    assert rel_err(dx, ref) < 1e-12


def test_pack_dgrad_s2_parity_matches_autograd():
    g = torch.Generator().manual_seed(3)
    c = 8
    w = bf16_round(torch.randn(c, c, 3, 3, generator=g))
    x = torch.randn(2, c, 16, 16, generator=g, dtype=torch.float64, requires_grad=True)
    dy = bf16_round(torch.randn(2, c, 8, 8, generator=g))
    (ref,) = torch.autograd.grad(F.conv2d(F.pad(x, (0, 1, 0, 1)), w.double(), stride=2), x, dy.double())
    dx = torch.zeros(2, 16, 16, c, dtype=torch.float64)
    for q in range(4):
        ph, pw = q >> 1, q & 1
        mat, dh, dw = pack_conv3x3(w, 3 + q)
        assert len(dh) == (2 - ph) * (2 - pw)
        dx[:, ph::2, pw::2] = emulate_gemm(dy.permute(0, 2, 3, 1), mat, dh, dw, 1, 8, 8)
    assert rel_err(dx.permute(0, 3, 1, 2), ref) < 1e-12


def test_trainconfig_mirrors_reference_override():
    from tml_image_editing_defense_b200.configs import TrainConfig
    c = TrainConfig(norm_type="l2", eps=1.0, step_size=1.0, grad_reps=1)
    assert (c.eps, c.step_size, c.grad_reps) == (32, 7.5, 10)          # configs.py:152-155
    c = TrainConfig(norm_type="linf")
    assert (c.eps, c.step_size, c.grad_reps) == (0.1, 0.006, 5)        # configs.py:156-159
    c = TrainConfig(norm_type="linf", eps=32 / 255, step_size=4 / 255, grad_reps=1, override_from_norm_type=False)
    assert c.eps == 32 / 255 and c.grad_reps == 1
    with pytest.raises(ValueError):
        TrainConfig(apply_loss_on_images=False, apply_loss_on_latents=False)
    # defaults are the reference's (configs.py:103-111): losses on decoded images, perturbation loss on
    d = TrainConfig()
    assert (d.apply_loss_on_images, d.apply_loss_on_latents, d.perturbation_loss_lambda) == (True, False, 1.0)
    assert (d.rec_loss_lambda, d.n_optimization_steps, d.n_denoising_steps_per_iteration, d.seed) == (1.0, 200, 4, 42)
    assert (d.guidance_scale, d.eta, d.use_fixed_noise, d.n_noise, d.norm_type) == (3.0, 0.9, True, 1, "l2")
    e = TrainConfig.encoder_attack(norm_type="linf")
    assert (e.apply_loss_on_images, e.apply_loss_on_latents, e.perturbation_loss_lambda) == (False, True, 0.0)


def test_shard_indices_partition():
    from tml_image_editing_defense_b200.dataset import shard_indices, shard_counts, SyntheticImageDataset
    for n in (0, 1, 7, 64):
        for w in (1, 2, 4, 8):
            parts = [shard_indices(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert [len(p) for p in parts] == shard_counts(n, w)
    ds = SyntheticImageDataset(5, resolution=8, seed=3)
    a = ds.batch([1, 3])
    assert torch.equal(a[1], ds.image(3)) and a.min() >= -1 and a.max() < 1


def test_losses_mirror_reference(golden_dir):
    from tml_image_editing_defense_b200 import losses
    d = np.load(golden_dir / "losses.npz")
    a, b = torch.from_numpy(d["a"]), torch.from_numpy(d["b"])
    assert torch.equal(losses.perturbation_loss(a, b), torch.from_numpy(d["perturbation_loss"]))
    assert torch.equal(losses.LpDistance(2)(a, b), torch.from_numpy(d["l2_distance"]))
    assert torch.equal(losses.LpDistance(float("inf"))(a, b), torch.from_numpy(d["linf_distance"]))
    assert torch.equal(losses.LpRegularization(2)([a, b]), torch.from_numpy(d["l2_regularization"]))
    assert torch.equal(losses.CosineSimilarity()(a, b), torch.from_numpy(d["cosine"]))


def test_parser_flag_names():
    from tml_image_editing_defense_b200.parser import parse_args
    a = parse_args(["--resolution", "1024", "--train_batch_size", "4", "--seed", "7", "--mixed_precision", "bf16"])
    assert a.resolution == 1024 and a.train_batch_size == 4 and a.seed == 7


def test_image_prompt_dataset_from_jpegs(tmp_path):
    """data/dataset.py:7-43: rglob('*.jpg') -> Resize(bilinear) -> CenterCrop -> ToTensor -> Normalize(.5,.5)."""
    from PIL import Image
    from torchvision import transforms
    from tml_image_editing_defense_b200.dataset import ImagePromptDataset
    rng = np.random.default_rng(0)
    sub = tmp_path / "nested"
    sub.mkdir()
    paths = [tmp_path / "b.jpg", sub / "a.jpg"]
    for p, (h, w) in zip(paths, [(96, 160), (200, 120)]):            # landscape and portrait, neither square
        Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)).save(p, quality=95)
    (tmp_path / "ignored.png").write_bytes(b"")                        # only *.jpg is collected (dataset.py:13)
    ds = ImagePromptDataset(str(tmp_path), "a photo", resolution=64)
    assert len(ds) == 2
    ref_tf = transforms.Compose([transforms.Resize(64, interpolation=transforms.InterpolationMode.BILINEAR),
                                 transforms.CenterCrop(64), transforms.ToTensor(), transforms.Normalize([0.5], [0.5])])
    for i, p in enumerate(sorted(paths)):                              # this build sorts the paths (deterministic shards)
        img, prompt = ds[i]
        assert prompt == "a photo"
        assert img.shape == (3, 64, 64) and img.dtype == torch.float32
        assert float(img.min()) >= -1.0 and float(img.max()) <= 1.0
        assert torch.equal(img, ref_tf(Image.open(p).convert("RGB")))
    raw = ImagePromptDataset.get_image_transform_no_normalization(64)(Image.open(paths[0]).convert("RGB"))
    assert float(raw.min()) >= 0.0 and float(raw.max()) <= 1.0
    assert len(ImagePromptDataset(str(sub), "", resolution=64)) == 1
    assert len(ImagePromptDataset(str(tmp_path / "nested"), "", resolution=32)[0][0][0]) == 32


def test_load_checkpoint_round_trip(tmp_path):
    from tml_image_editing_defense_b200.vae import SD15_VAE
    from tml_image_editing_defense_b200.weights import encoder_param_shapes, load_checkpoint, random_init_state_dict
    sd = random_init_state_dict(SD15_VAE, seed=3, include_decoder=True)
    assert list(sd) == [n for n, _ in encoder_param_shapes(SD15_VAE, include_decoder=True)]
    assert sum(v.numel() for v in sd.values()) == 83_653_863          # the SD VAE (encoder 34 163 664 + decoder)
    torch.save(sd, tmp_path / "vae.pt")
    back = load_checkpoint(str(tmp_path / "vae.pt"))
    assert list(back) == list(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    torch.save({"state_dict": {k: v.to(torch.bfloat16) for k, v in sd.items()}}, tmp_path / "wrapped.pt")
    wrapped = load_checkpoint(str(tmp_path / "wrapped.pt"))             # Lightning-style wrapper, bf16 payload
    assert list(wrapped) == list(sd) and wrapped["encoder.conv_in.weight"].dtype == torch.bfloat16
    torch.testing.assert_close(wrapped["quant_conv.weight"].float(), sd["quant_conv.weight"], rtol=1e-2, atol=1e-3)


def test_parser_keeps_reference_flag_names():
    from tml_image_editing_defense_b200.parser import parse_args
    a = parse_args(["--gradient_checkpointing", "--allow_tf32", "--resolution", "256", "--train_batch_size", "4",
                    "--seed", "7", "--mixed_precision", "bf16", "--max_train_steps", "3", "--output_dir", "o",
                    "--local_rank", "0"])
    assert a.gradient_checkpointing and a.allow_tf32 and a.resolution == 256 and a.train_batch_size == 4
    assert a.seed == 7 and a.max_train_steps == 3 and a.output_dir == "o"
    b = parse_args([])
    assert not b.gradient_checkpointing and not b.allow_tf32


def test_sharded_image_loader_partitions_and_orders():
    """The host pipeline: shards r::world cover the dataset exactly once, batches keep the shard order, the last batch
    may be short, and shuffling permutes inside a shard only."""
    from tml_image_editing_defense_b200.dataset import SyntheticImageDataset
    from tml_image_editing_defense_b200.loader import ShardedImageLoader
    ds = SyntheticImageDataset(11, resolution=8, seed=5)
    seen = []
    for r in range(3):
        ld = ShardedImageLoader(ds, batch_size=2, rank=r, world_size=3, device="cpu", fetch=ds.batch)
        assert len(ld) == (len(range(r, 11, 3)) + 1) // 2
        got = []
        for idx, imgs, prompts in ld:
            assert imgs.shape == (len(idx), 3, 8, 8) and len(prompts) == len(idx)
            for j, i in enumerate(idx):
                assert torch.equal(imgs[j], ds.image(i))
            got += idx
        assert got == list(range(r, 11, 3))
        seen += got
    assert sorted(seen) == list(range(11))
    a = ShardedImageLoader(ds, 4, rank=1, world_size=2, device="cpu", shuffle=True, seed=3)
    b = ShardedImageLoader(ds, 4, rank=1, world_size=2, device="cpu", shuffle=True, seed=3)
    assert a.indices == b.indices and sorted(a.indices) == list(range(1, 11, 2)) and a.indices != list(range(1, 11, 2))
    per_item = ShardedImageLoader(ds, 3, device="cpu")              # the dataset[i] -> (image, prompt) path
    assert [i for idx, _, _ in per_item for i in idx] == list(range(11))
    assert list(ShardedImageLoader(SyntheticImageDataset(0, 8), 2, device="cpu")) == []   # empty shard
