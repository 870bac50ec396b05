"""GPU parity tests of the individual kernels, called through the C ABI (run with -m gpu on a B200)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.test_oracle import assert_bit_equal, LINF, L2  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def lib():
    from tml_image_editing_defense_b200 import _lib
    return _lib.load()


# ------------------------------------------------------------------ K9: fused PGD update
@pytest.mark.parametrize("name", LINF)
def test_pgd_linf_bit_exact_vs_reference_golden(golden_dir, dev, name):
    from tml_image_editing_defense_b200 import ops
    d = np.load(golden_dir / f"pgd_linf_{name}.npz")
    eps, step, lo, hi = [float(v) for v in d["params"]]
    xa = torch.from_numpy(d["x_adv"]).to(dev)
    out = ops.pgd_step_linf_(xa, torch.from_numpy(d["grad"]).to(dev), torch.from_numpy(d["x"]).to(dev), eps, step,
                             lo, hi)
    assert_bit_equal(out.cpu().numpy(), d["out"])


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1023, 4 * 1000 * 1000 + 1])
def test_pgd_linf_bit_exact_vs_oracle_odd_sizes(dev, n):
    from oracle.pgd_oracle import pgd_step_linf
    from tml_image_editing_defense_b200 import ops
    g = torch.Generator().manual_seed(n)
    x = torch.rand(n, generator=g) * 2 - 1
    xa = (x + (torch.rand(n, generator=g) - 0.5) * 0.2).clamp(-1, 1)
    gr = torch.randn(n, generator=g)
    gr[::7] = 0.0
    ref = pgd_step_linf(xa, gr, x, 32 / 255, 4 / 255, -1, 1)
    out = ops.pgd_step_linf_(xa.to(dev), gr.to(dev), x.to(dev), 32 / 255, 4 / 255, -1.0, 1.0)
    assert_bit_equal(out.cpu().numpy(), ref.numpy())


def test_pgd_linf_full_size_properties(dev):
    """BASELINE cfg-2 size (64x3x512x512): result stays in the eps-ball and the clamp range, moves
    by exactly +-step or 0 before projection, and is idempotent under a zero gradient."""
    from tml_image_editing_defense_b200 import ops
    n = 64 * 3 * 512 * 512
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(n, generator=g, device=dev) * 2 - 1
    xa = x.clone()
    gr = torch.randn(n, generator=g, device=dev)
    eps, step = 32 / 255, 4 / 255
    for _ in range(3):
        ops.pgd_step_linf_(xa, gr, x, eps, step, -1.0, 1.0)
    assert float((xa - x).abs().max()) <= eps + 1e-7
    assert float(xa.min()) >= -1 and float(xa.max()) <= 1
    ref = torch.clamp(torch.minimum(torch.maximum(x - 3 * step * gr.sign(), x - eps), x + eps), -1, 1)
    assert float((xa - ref).abs().max()) < 1e-6
    before = xa.clone()
    ops.pgd_step_linf_(xa, torch.zeros_like(gr), x, eps, step, -1.0, 1.0)
    assert torch.equal(xa, before)


@pytest.mark.parametrize("name", L2)
def test_pgd_l2_vs_reference_golden(golden_dir, dev, name):
    from tml_image_editing_defense_b200 import ops
    d = np.load(golden_dir / f"pgd_l2_{name}.npz")
    eps, step, lo, hi = [float(v) for v in d["params"]]
    mask = torch.from_numpy(d["mask"]).to(dev) if d["mask"].size else None
    out = ops.pgd_step_l2_(torch.from_numpy(d["x_adv"]).to(dev), torch.from_numpy(d["grad"]).to(dev),
                           torch.from_numpy(d["x"]).to(dev), mask, eps, step, lo, hi)
    # fp32 norms are reduced in a different order than ATen's: tolerance 2e-6 absolute on values in [-1,1]
    np.testing.assert_allclose(out.cpu().numpy(), d["out"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", ["ref", "tight"])
def test_universal_step_vs_reference_golden(golden_dir, dev, name):
    from tml_image_editing_defense_b200 import ops
    d = np.load(golden_dir / f"universal_update_{name}.npz")
    eps, step = [float(v) for v in d["params"]]
    out = ops.universal_step_(torch.from_numpy(d["delta"]).to(dev), torch.from_numpy(d["grad"]).to(dev),
                              torch.from_numpy(d["source"]).to(dev), eps, step)
    np.testing.assert_allclose(out.cpu().numpy(), d["out"], rtol=0, atol=2e-6)


def test_universal_project_bit_exact_vs_oracle(dev):
    """old/train_noise.py:183-185 as its own entry point: one source (the reference statement, bit-exact) and the
    (min, max) pair a sharded step uses; +-0, NaN and on-boundary values included."""
    from oracle.pgd_oracle import universal_project
    from tml_image_editing_defense_b200 import ops
    g = torch.Generator().manual_seed(5)
    n = 3 * 37 * 41
    delta = (torch.rand((1, n), generator=g) - 0.5) * 0.8
    src = torch.rand((4, n), generator=g) * 2 - 1
    src[0, :8] = torch.tensor([1.0, -1.0, 0.0, -0.0, 1.0, -1.0, 0.999999, -0.999999])
    delta[0, :8] = torch.tensor([0.3, -0.3, 0.0, -0.0, -0.3, 0.3, float("nan"), 1e-8])
    for k in (1, 2, 4):
        out = ops.universal_project_(delta.clone().to(dev), src[:k].contiguous().to(dev))
        ref = universal_project(delta, src[:k])
        assert_bit_equal(out.cpu().numpy(), ref.numpy())          # NaN stays NaN (payload aside), everything else bit-equal
    lo, hi = src.amin(0), src.amax(0)
    out = ops.universal_project_(delta.clone().to(dev), torch.stack([lo, hi]).to(dev)).cpu()
    ok = ~torch.isnan(out)
    assert float(((src + out)[:, ok[0]]).abs().max()) <= 1.0 + 1e-6


def test_add_delta_and_batch_sum(dev):
    from tml_image_editing_defense_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 3, 16, 16, generator=g)
    dl = torch.randn(1, 3, 16, 16, generator=g)
    assert torch.equal(ops.add_delta(x.to(dev), dl.to(dev)).cpu(), x + dl)
    s = ops.batch_sum(x.to(dev), 0.5).cpu()
    ref = torch.zeros(1, 3, 16, 16)
    for b in range(5):
        ref += x[b]
    assert torch.equal(s, ref * 0.5)


# ------------------------------------------------------------------ K8: sample + loss + dmoments
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("use_noise", [True, False])
def test_latent_loss_vs_autograd(dev, kind, use_noise):
    from oracle.encoder_oracle import DiagonalGaussianDistribution, latent_loss
    from tml_image_editing_defense_b200 import ops
    g = torch.Generator().manual_seed(kind)
    m = torch.randn(3, 8, 8, 8, generator=g)
    m[:, 4:] *= 12            # exercise the logvar clamp on both sides
    m[0, 4, 0, 0], m[0, 4, 0, 1] = 25.0, -35.0
    t = torch.randn(3, 4, 8, 8, generator=g)
    noise = torch.randn(3, 4, 8, 8, generator=g) if use_noise else None
    mm = m.clone().requires_grad_(True)
    dist = DiagonalGaussianDistribution(mm)
    z_ref = dist.mean + dist.std * noise if use_noise else dist.mean
    l_ref = latent_loss(z_ref, t, kind)
    (dm_ref,) = torch.autograd.grad(l_ref.sum(), mm)
    z, l, dm = ops.latent_loss(m.to(dev), noise.to(dev) if use_noise else None, t.to(dev), kind)
    torch.testing.assert_close(z.cpu(), z_ref.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(l.cpu(), l_ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dm.cpu(), dm_ref, rtol=1e-4, atol=1e-7)


# ------------------------------------------------------------------ K1-K5: tcgen05 implicit GEMM
def test_tcgen05_gemm_suite(dev, lib):
    from tests.gpu_check import run_gemm_suite
    lib.tml_debug_set_gemm_impl(0)
    before = np.zeros(2, np.int64)
    lib.tml_launch_counts(before.ctypes.data_as(C.POINTER(C.c_int64)))
    assert run_gemm_suite(lib, dev)
    after = np.zeros(2, np.int64)
    lib.tml_launch_counts(after.ctypes.data_as(C.POINTER(C.c_int64)))
    assert after[0] - before[0] >= 32, "the tcgen05 kernel did not run"


def test_pgd_l2_full_size_properties(dev):
    """Size-independent properties at a large size: every image ends inside its L2 ball and the clamp
    range; a zero gradient leaves an in-ball iterate untouched."""
    from tml_image_editing_defense_b200 import ops
    B, C, H, W = 16, 3, 512, 512
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.rand((B, C, H, W), generator=g, device=dev) * 2 - 1
    xa = x.clone()
    gr = torch.randn((B, C, H, W), generator=g, device=dev)
    eps, step = 8.0, 3.0
    for _ in range(4):
        ops.pgd_step_l2_(xa, gr, x, None, eps, step, -1.0, 1.0)
    d = (xa - x).reshape(B, -1).norm(dim=1)
    assert float(d.max()) <= eps * (1 + 1e-5)
    assert float(d.min()) > 0.5 * eps            # 4 steps of 3.0 along a fixed direction saturate the ball
    assert float(xa.min()) >= -1 and float(xa.max()) <= 1
    before = xa.clone()
    ops.pgd_step_l2_(xa, torch.zeros_like(gr), x, None, eps * 2, step, -1.0, 1.0)
    assert torch.equal(xa, before)


def test_pgd_empty_and_unaligned(dev):
    from tml_image_editing_defense_b200 import _lib, ops
    e = torch.empty(0, device=dev)
    assert ops.pgd_step_linf_(e, e.clone(), e.clone(), 0.1, 0.01, -1.0, 1.0).numel() == 0
    base = torch.zeros(9, device=dev)
    with pytest.raises(_lib.TmlError):
        ops.pgd_step_linf_(base[1:], base[1:].clone(), torch.zeros(8, device=dev), 0.1, 0.01, -1.0, 1.0)


def _run_gemm_suite(extra_env):
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    env = dict(os.environ, **extra_env)
    r = subprocess.run([sys.executable, str(root / "tests" / "gpu_check.py"), "--gemm-only", "--impl", "tc"], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ALL OK" in r.stdout


def test_tcgen05_gemm_suite_cta_pairs(dev):
    """The cta_group::2 (CTA-pair) instantiation of the pixel-major kernel's halo mode, enabled with TML_PAIR=1 (by
    default pairs are used outside halo mode only).  A subprocess because the switch is read once."""
    _run_gemm_suite({"TML_PAIR": "1"})


def test_tcgen05_gemm_suite_fallback_paths(dev):
    """The same suite with the fast paths switched off (no operand-swapped kernel, no CTA pairs, register epilogue
    instead of TMA stores): the A/B switches documented in DESIGN.md must stay correct."""
    _run_gemm_suite({"TML_NO_SWAP": "1", "TML_PAIR": "0", "TML_NO_TMA_STORE": "1"})


def test_encoder_walk_with_fused_input_groupnorm(dev):
    """TML_FUSE_INGN=1 (off by default): the CTA-pair 3x3 convolutions take the RAW GroupNorm input and
    apply scale / shift / SiLU on their operand path.  The layer-by-layer comparison with the oracle (every saved
    activation, every backward stage) at 256^2 -- where the 128^2 stage runs on pairs -- must hold with it switched on."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, str(root / "tests" / "gpu_check.py"), "--skip-gemm", "--impl", "tc", "--res", "256",
                        "--batch", "2"], env=dict(os.environ, TML_FUSE_INGN="1"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ALL OK" in r.stdout
