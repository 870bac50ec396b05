"""CPU tests of the diffusion-attack host logic: scheduler formulas, the UNet
restatement's size / key names, and the full compute_grad flow with the oracle VAE and a tiny UNet."""
import math

import numpy as np
import pytest
import torch

from tml_image_editing_defense_b200.schedulers import DDIMScheduler, LCMScheduler
from tml_image_editing_defense_b200.unet_torch import UNet2DConditionModel, tiny_unet_config


def _alphas():
    betas = np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=np.float64) ** 2
    return np.cumprod(1 - betas)


def test_add_noise_and_ddim_step_formulas():
    ac = _alphas()
    s = DDIMScheduler()
    s.set_timesteps(4)
    assert s.timesteps.tolist() == [751, 501, 251, 1]
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    e = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    xt = s.add_noise(x0, e, [751])
    np.testing.assert_allclose(xt.numpy(), math.sqrt(ac[751]) * x0.numpy() + math.sqrt(1 - ac[751]) * e.numpy(), rtol=1e-12)
    eta = 0.9
    vn = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    out = s.step(e, 751, xt, eta=eta, variance_noise=vn)
    a_t, a_p = ac[751], ac[501]
    x0h = (xt.numpy() - math.sqrt(1 - a_t) * e.numpy()) / math.sqrt(a_t)
    sig = eta * math.sqrt((1 - a_p) / (1 - a_t)) * math.sqrt(1 - a_t / a_p)
    ref = math.sqrt(a_p) * x0h + math.sqrt(1 - a_p - sig ** 2) * e.numpy() + sig * vn.numpy()
    np.testing.assert_allclose(out.numpy(), ref, rtol=1e-10)
    last = s.step(e, 1, xt, eta=0.0)          # prev_t < 0 -> final_alpha_cumprod = alphas_cumprod[0]
    assert torch.isfinite(last).all()


def test_lcm_step_formulas():
    ac = _alphas()
    s = LCMScheduler()
    s.set_timesteps(4)
    assert s.timesteps.tolist() == [999, 759, 519, 279]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 4, 8, 8, generator=g, dtype=torch.float64)
    e = torch.randn(1, 4, 8, 8, generator=g, dtype=torch.float64)
    vn = torch.randn(1, 4, 8, 8, generator=g, dtype=torch.float64)
    t = 759
    sc = 10.0 * t
    c_skip, c_out = 0.25 / (sc ** 2 + 0.25), sc / math.sqrt(sc ** 2 + 0.25)
    x0h = (x.numpy() - math.sqrt(1 - ac[t]) * e.numpy()) / math.sqrt(ac[t])
    den = c_out * x0h + c_skip * x.numpy()
    ref = math.sqrt(ac[519]) * den + math.sqrt(1 - ac[519]) * vn.numpy()
    np.testing.assert_allclose(s.step(e, t, x, variance_noise=vn).numpy(), ref, rtol=1e-10)
    t = 279                                   # last step returns the denoised sample
    sc = 10.0 * t
    c_skip, c_out = 0.25 / (sc ** 2 + 0.25), sc / math.sqrt(sc ** 2 + 0.25)
    x0h = (x.numpy() - math.sqrt(1 - ac[t]) * e.numpy()) / math.sqrt(ac[t])
    np.testing.assert_allclose(s.step(e, t, x).numpy(), c_out * x0h + c_skip * x.numpy(), rtol=1e-10)


def test_unet_restatement_size_and_keys():
    with torch.device("meta"):
        m = UNet2DConditionModel()
    assert sum(p.numel() for p in m.parameters()) == 859_520_964          # the known SD-1.5 UNet size
    keys = set(m.state_dict().keys())
    for k in ["conv_in.weight", "time_embedding.linear_2.bias", "down_blocks.0.resnets.1.time_emb_proj.weight",
              "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight",
              "down_blocks.2.downsamplers.0.conv.weight", "mid_block.attentions.0.proj_in.weight",
              "up_blocks.1.attentions.2.transformer_blocks.0.ff.net.0.proj.weight",
              "up_blocks.3.resnets.2.conv_shortcut.weight", "up_blocks.0.upsamplers.0.conv.weight", "conv_out.bias"]:
        assert k in keys, k
    assert "down_blocks.3.attentions.0.norm.weight" not in keys and "up_blocks.0.attentions.0.norm.weight" not in keys


def test_diffusion_compute_grad_flow_with_oracle_vae():
    from oracle.decoder_oracle import make_vae_oracle
    from oracle.encoder_oracle import EncoderConfig
    from tml_image_editing_defense_b200.configs import TrainConfig
    from tml_image_editing_defense_b200.diffusion import DiffusionAttack
    vae = make_vae_oracle(0, EncoderConfig(block_out_channels=(32, 32, 32, 32), layers_per_block=1, norm_num_groups=8))
    torch.manual_seed(3)
    unet = UNet2DConditionModel(tiny_unet_config()).requires_grad_(False)
    cfg = TrainConfig(norm_type="linf", override_from_norm_type=False, device="cpu", apply_loss_on_images=True,
                      apply_loss_on_latents=False, perturbation_loss_lambda=1.0)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    tgt = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    pe = torch.randn(2, 7, 32, generator=g)
    nz = [torch.randn(1, 4, 8, 8, generator=g)]
    outs = []
    for ck in (False, True):
        da = DiffusionAttack(cfg, vae, unet, DDIMScheduler(), use_checkpointing=ck, unet_dtype=torch.float32)
        gr, loss, img, ld = da.compute_grad(x, pe, x, tgt, None, nz)
        outs.append((gr, float(loss)))
        assert gr.shape == x.shape and img.shape == x.shape and float(gr.abs().sum()) > 0
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-5, atol=1e-8)    # checkpointing does not change the gradient
    assert outs[0][1] == pytest.approx(outs[1][1], rel=1e-6)
