"""GPU parity tests of the decoder path (vae.decode + image-space losses, main.py:156-171)
against the oracle, through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.helpers import cosine, rel_err  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def models(dev):
    from oracle.decoder_oracle import make_vae_oracle
    from oracle.encoder_oracle import perturb_affine_params
    from tml_image_editing_defense_b200.vae import AutoencoderKL
    oracle = make_vae_oracle(0)
    perturb_affine_params(oracle, 1234)
    vae = AutoencoderKL(device=str(dev)).load_state_dict(oracle.state_dict())
    assert vae.has_decoder
    return oracle, vae


def test_decoder_layerwise_parity_64(dev):
    from tests.gpu_check_decoder import run_decoder
    assert run_decoder(dev, 64, 2)


@pytest.mark.parametrize("lat", [(16, 16), (64, 64), (40, 56)])
def test_decode_and_dz_vs_oracle(dev, models, lat):
    """16^2 latents (128^2 images), the benchmark size 64^2 (512^2 images: every fast geometry incl. the pitch-66 halo
    mode and the attention epilogues at 4096 tokens), and a size off the fast paths (40 x 56 -> 320 x 448)."""
    oracle, vae = models
    lh, lw = lat
    g = torch.Generator().manual_seed(3)
    z = torch.randn((2, 4, lh, lw), generator=g)
    dimg = torch.randn((2, 3, 8 * lh, 8 * lw), generator=g)
    oracle.to(dev)
    z, dimg = z.to(dev), dimg.to(dev)
    with torch.enable_grad():
        zz = z.clone().requires_grad_(True)
        ref = oracle.decode(zz)
        ref.backward(dimg)
    oracle.to("cpu")
    ref, zgrad_ref = ref.detach(), zz.grad.detach()
    del zz
    torch.cuda.empty_cache()
    zc = z.clone().requires_grad_(True)
    img = vae.decode(zc).sample                     # autograd seam (main.py:156)
    assert rel_err(img.detach(), ref) < 4e-2
    img.backward(dimg)
    assert cosine(zc.grad, zgrad_ref) >= 0.998


def test_image_loss_kernel_vs_torch(dev):
    from oracle.decoder_oracle import image_losses
    from tml_image_editing_defense_b200 import ops
    g = torch.Generator().manual_seed(5)
    out = torch.randn((3, 3, 32, 48), generator=g)
    tgt = torch.randn((3, 3, 32, 48), generator=g)
    src = torch.randn((3, 3, 32, 48), generator=g)
    o = out.clone().requires_grad_(True)
    loss, rec_ref, pert_ref = image_losses(o, tgt, src, 0.7, 1.3)
    (d_ref,) = torch.autograd.grad(loss.sum(), o)
    rec, pert, dout = ops.image_loss(out.to(dev), tgt.to(dev), src.to(dev), 0.7, 1.3)
    torch.testing.assert_close(rec.cpu(), rec_ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(pert.cpu(), pert_ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dout.cpu(), d_ref, rtol=1e-4, atol=1e-8)
    rec2, pert2, d2 = ops.image_loss(out.to(dev), tgt.to(dev), None, 1.0, 0.0)
    assert float(pert2.abs().max()) == 0.0
    torch.testing.assert_close(d2.cpu(), (out - tgt) / (out - tgt).reshape(3, -1).norm(dim=1).view(3, 1, 1, 1), rtol=1e-4, atol=1e-8)


def test_autoencoder_attack_grad_vs_oracle(dev, models):
    """compute_grad with the reference's default image-space losses (UNet removed): gradient through decoder
    and encoder vs the fp32 oracle."""
    from oracle.decoder_oracle import autoencoder_attack_grad
    oracle, vae = models
    g = torch.Generator().manual_seed(8)
    x = torch.rand((2, 3, 64, 64), generator=g) * 2 - 1
    tgt = torch.rand((2, 3, 64, 64), generator=g) * 2 - 1
    noise = torch.randn((2, 4, 8, 8), generator=g)
    od = oracle.to(dev)
    g_ref, l_ref, out_ref = autoencoder_attack_grad(od, x.to(dev), tgt.to(dev), x.to(dev), noise.to(dev), 1.0, 1.0)
    oracle.to("cpu")
    gg, rec, pert, img = vae.attack_grad_images(x.to(dev), tgt.to(dev), x.to(dev), noise.to(dev), 1.0, 1.0)
    c = cosine(gg, g_ref)
    print("autoencoder attack gradient cosine", c)
    assert c >= 0.998
    torch.testing.assert_close(rec + pert, l_ref, rtol=3e-2, atol=0)
    assert rel_err(img, out_ref) < 5e-2


def test_trainer_image_loss_mode(dev, models):
    from tml_image_editing_defense_b200.configs import TrainConfig
    from tml_image_editing_defense_b200.trainer import Trainer
    _, vae = models
    g = torch.Generator().manual_seed(9)
    x = (torch.rand((2, 3, 64, 64), generator=g) * 2 - 1).to(dev)
    tgt = (torch.rand((1, 3, 64, 64), generator=g) * 2 - 1).to(dev)
    cfg = TrainConfig(norm_type="linf", eps=0.1, step_size=0.01, grad_reps=1, override_from_norm_type=False,
                      n_optimization_steps=10, apply_loss_on_images=True, apply_loss_on_latents=False,
                      perturbation_loss_lambda=1.0, device=str(dev))
    tr = Trainer(cfg, vae, micro_batch=1)
    xa = tr.run(x, target_image=tgt)
    assert float((xa - x).abs().max()) <= 0.1 + 1e-6
    assert tr.loss_history[-1] < tr.loss_history[0]


def test_decoder_vs_golden_fixture(golden_dir, dev, models):
    _, vae = models
    d = np.load(golden_dir / "decoder_64.npz")
    img = vae.decode(torch.from_numpy(d["z"]).to(dev)).sample
    assert rel_err(img.cpu(), torch.from_numpy(d["image"])) < 4e-2
    x = torch.from_numpy(d["x"]).to(dev)
    gg, rec, pert, out = vae.attack_grad_images(x, torch.from_numpy(d["target_image"]).to(dev), x,
                                                torch.from_numpy(d["noise"]).to(dev), 1.0, 1.0)
    assert cosine(gg.cpu(), torch.from_numpy(d["grad"])) >= 0.998
    np.testing.assert_allclose((rec + pert).cpu().numpy(), d["loss"], rtol=3e-2)


def test_diffusion_attack_hybrid_vs_oracle(dev, models):
    """encode (our kernels) -> add_noise -> UNet steps with CFG (PyTorch library module, bf16,
    checkpointed) -> decode (our kernels) -> image losses -> gradient, against the all-fp32 oracle pipeline."""
    from tml_image_editing_defense_b200.configs import TrainConfig
    from tml_image_editing_defense_b200.diffusion import DiffusionAttack
    from tml_image_editing_defense_b200.schedulers import DDIMScheduler
    from tml_image_editing_defense_b200.unet_torch import UNet2DConditionModel, tiny_unet_config
    oracle, vae = models
    torch.manual_seed(11)
    unet = UNet2DConditionModel(tiny_unet_config()).requires_grad_(False).to(dev)
    cfg = TrainConfig(norm_type="linf", override_from_norm_type=False, device=str(dev), apply_loss_on_images=True,
                      apply_loss_on_latents=False, perturbation_loss_lambda=1.0, n_denoising_steps_per_iteration=4)
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).to(dev)
    tgt = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).to(dev)
    pe = torch.randn(2, 7, 32, generator=g).to(dev)
    nz = [torch.randn(2, 4, 8, 8, generator=g).to(dev)]
    od = oracle.to(dev)
    ref = DiffusionAttack(cfg, od, unet, DDIMScheduler(), use_checkpointing=False, unet_dtype=torch.float32)
    g_ref, l_ref, img_ref, _ = ref.compute_grad(x, pe, x, tgt, None, nz)
    oracle.to("cpu")
    ours = DiffusionAttack(cfg, vae, unet, DDIMScheduler(), use_checkpointing=True, unet_dtype=torch.bfloat16)
    gg, l, img, _ = ours.compute_grad(x, pe, x, tgt, None, nz)
    c = cosine(gg, g_ref)
    print("diffusion attack gradient cosine", c)
    assert c >= 0.99
    assert abs(float(l) - float(l_ref)) / float(l_ref) < 0.03
    assert rel_err(img, img_ref) < 0.08
