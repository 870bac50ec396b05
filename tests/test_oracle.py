"""CPU tests: the oracle against the golden vectors produced by the reference's own code
(oracle/gen_golden.py) and against an independent numpy restatement."""
import numpy as np
import pytest
import torch

from oracle import pgd_oracle as po
from oracle.encoder_oracle import make_oracle, EncoderConfig, OracleAutoencoderKL

LINF = ["refdefault", "northstar", "unit", "playground"]
L2 = ["refdefault", "small_eps", "masked"]


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_equal(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    nan = np.isnan(a) & np.isnan(b)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.array_equal(_bits(a)[~nan], _bits(b)[~nan])


@pytest.mark.parametrize("name", LINF)
def test_linf_oracle_matches_reference_bits(golden_dir, name):
    d = np.load(golden_dir / f"pgd_linf_{name}.npz")
    eps, step, lo, hi = d["params"]
    y = po.pgd_step_linf(torch.from_numpy(d["x_adv"]), torch.from_numpy(d["grad"]), torch.from_numpy(d["x"]),
                         eps, step, lo, hi)
    assert_bit_equal(y.numpy(), d["out"])
    y2 = po.pgd_step_linf_numpy(d["x_adv"], d["grad"], d["x"], eps, step, lo, hi)
    assert_bit_equal(y2, d["out"])


@pytest.mark.parametrize("name", L2)
def test_l2_oracle_matches_reference_bits(golden_dir, name):
    d = np.load(golden_dir / f"pgd_l2_{name}.npz")
    eps, step, lo, hi = d["params"]
    mask = torch.from_numpy(d["mask"]) if d["mask"].size else None
    y = po.pgd_step_l2(torch.from_numpy(d["x_adv"]), torch.from_numpy(d["grad"]), torch.from_numpy(d["x"]), mask,
                       eps, step, lo, hi)
    assert_bit_equal(y.numpy(), d["out"])


@pytest.mark.parametrize("name", ["ref", "tight"])
def test_universal_update_matches_reference_bits(golden_dir, name):
    d = np.load(golden_dir / f"universal_update_{name}.npz")
    eps, step = d["params"]
    y = po.universal_update(torch.from_numpy(d["delta"]), torch.from_numpy(d["grad"]), torch.from_numpy(d["source"]),
                            eps, step)
    assert_bit_equal(y.numpy(), d["out"])


def test_losses_match_reference(golden_dir):
    d = np.load(golden_dir / "losses.npz")
    a, b = torch.from_numpy(d["a"]), torch.from_numpy(d["b"])
    assert_bit_equal(po.perturbation_loss(a, b).numpy(), d["perturbation_loss"])
    assert_bit_equal(po.lp_distance(a, b, 2).numpy(), d["l2_distance"])
    assert_bit_equal(po.lp_distance(a, b, float("inf")).numpy(), d["linf_distance"])
    assert_bit_equal(po.lp_regularization([a, b], 2).numpy(), d["l2_regularization"])
    assert_bit_equal(po.cosine_similarity_plus_one(a, b).numpy(), d["cosine"])


def test_sign_special_values():
    g = torch.tensor([0.0, -0.0, 1e-30, -1e-30, float("nan"), float("inf")])
    assert torch.sign(g).tolist() == [0.0, 0.0, 1.0, -1.0, 0.0, 1.0]


def test_encoder_param_count_and_keys():
    m = OracleAutoencoderKL(EncoderConfig())
    assert sum(p.numel() for p in m.parameters()) == 34_163_664
    keys = set(m.state_dict().keys())
    for k in ["encoder.conv_in.weight", "encoder.down_blocks.0.resnets.0.norm1.weight",
              "encoder.down_blocks.1.resnets.0.conv_shortcut.weight", "encoder.down_blocks.2.downsamplers.0.conv.bias",
              "encoder.mid_block.attentions.0.group_norm.weight", "encoder.mid_block.attentions.0.to_q.weight",
              "encoder.mid_block.attentions.0.to_out.0.bias", "encoder.mid_block.resnets.1.conv2.weight",
              "encoder.conv_norm_out.bias", "encoder.conv_out.weight", "quant_conv.weight"]:
        assert k in keys, k
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in keys


def test_encoder_oracle_regression_64(golden_dir):
    from oracle.encoder_oracle import perturb_affine_params, encoder_attack_grad
    d = np.load(golden_dir / "encoder_64.npz")
    torch.set_num_threads(max(1, torch.get_num_threads()))
    m = make_oracle(0)
    perturb_affine_params(m, 1234)
    x = torch.from_numpy(d["x"])
    with torch.no_grad():
        mom = m.moments(x)
    np.testing.assert_allclose(mom.numpy(), d["moments"], rtol=1e-4, atol=1e-5)
    g, l, z = encoder_attack_grad(m, x, torch.from_numpy(d["target"]), torch.from_numpy(d["noise"]), 0)
    np.testing.assert_allclose(l.numpy(), d["loss_kind0"], rtol=1e-4)
    cos = float((g.flatten() @ torch.from_numpy(d["grad_kind0"]).flatten()) /
                (g.norm() * np.linalg.norm(d["grad_kind0"])))
    assert cos > 0.99999


def test_encoder_attack_loss_decreases():
    from oracle.encoder_oracle import perturb_affine_params
    m = make_oracle(0)
    g = torch.Generator().manual_seed(5)
    x = torch.rand((1, 3, 32, 32), generator=g) * 2 - 1
    tgt = torch.randn((1, 4, 4, 4), generator=g)
    noise = torch.randn((1, 4, 4, 4), generator=g)
    rec = []
    po.encoder_attack(m, x, tgt, noise, 6, 32 / 255, 4 / 255, -1, 1, kind=0,
                      record=lambda xa, gr, ls: rec.append(float(ls.sum())))
    assert rec[-1] < rec[0]


def test_decoder_param_count_and_keys():
    from oracle.decoder_oracle import OracleVAE
    m = OracleVAE(EncoderConfig())
    assert sum(p.numel() for p in m.parameters()) == 83_653_863          # the known SD VAE size
    assert sum(p.numel() for p in m.decoder.parameters()) + sum(p.numel() for p in m.post_quant_conv.parameters()) == 49_490_199
    keys = set(m.state_dict().keys())
    for k in ["post_quant_conv.weight", "decoder.conv_in.weight", "decoder.mid_block.attentions.0.to_out.0.bias",
              "decoder.up_blocks.0.resnets.2.conv2.weight", "decoder.up_blocks.0.upsamplers.0.conv.weight",
              "decoder.up_blocks.2.resnets.0.conv_shortcut.weight", "decoder.up_blocks.3.resnets.2.norm2.bias",
              "decoder.conv_norm_out.weight", "decoder.conv_out.bias"]:
        assert k in keys, k
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in keys


def test_decoder_oracle_regression_64(golden_dir):
    from oracle.decoder_oracle import make_vae_oracle
    from oracle.encoder_oracle import perturb_affine_params
    d = np.load(golden_dir / "decoder_64.npz")
    m = make_vae_oracle(0)
    perturb_affine_params(m, 1234)
    with torch.no_grad():
        img = m.decode(torch.from_numpy(d["z"]))
    np.testing.assert_allclose(img.numpy(), d["image"], rtol=1e-4, atol=1e-5)
