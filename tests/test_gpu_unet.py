"""GPU parity tests of the native UNet (csrc/unet.cu) through the C ABI, against the fp32 PyTorch restatement of
diffusers' UNet2DConditionModel (``unet_torch.py``: the oracle of this path -- the module itself lives in the absent
``diffusers``, so like the VAE its parity is unpinned beyond the parameter count 859 520 964 and the key set).

Tolerances (bf16 activations, fp32 accumulation): stage outputs within 5e-2 relative L2, stage gradients within 1e-1,
final prediction cosine >= 0.999, sample-gradient cosine >= 0.999 (tiny) / 0.999 (SD-1.5)."""
import pytest
import torch

from tests.helpers import cosine, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_unet_tiny_stagewise_vs_oracle(dev):
    """Every resnet / transformer / sampler output and every backward stage of a 3-level UNet (64/128/128 channels,
    8 heads of 8 and 16 channels, 5 prompt tokens padded and masked to 64) at 32 x 32 latents."""
    from tests.gpu_check_unet import run_unet
    ok, cy, cdx = run_unet(dev, "tiny", batch=2, size=32, tokens=5, verbose=False)
    assert ok
    assert cy >= 0.999 and cdx >= 0.999


def test_unet_tiny_odd_shapes(dev):
    """Non-square latents, a batch of 3 and 77 prompt tokens (two key tiles of 64, 51 masked slots)."""
    from tests.gpu_check_unet import make_oracle, tiny_native_config
    from tml_image_editing_defense_b200.unet import UNet2DConditionModel as NativeUNet
    cfg = tiny_native_config()
    m = make_oracle(cfg, 3, 2.0)
    native = NativeUNet(cfg, device=str(dev)).load_state_dict(m.state_dict())
    md = m.to(dev)
    g = torch.Generator().manual_seed(5)
    x = torch.randn((3, 4, 32, 64), generator=g).to(dev)
    ctx = torch.randn((3, 77, cfg.cross_attention_dim), generator=g).to(dev)
    dout = (torch.randn((3, 4, 32, 64), generator=g) * 1e-2).to(dev)
    xr = x.clone().requires_grad_(True)
    y_ref = md(xr, torch.tensor(981.0, device=dev), ctx).sample
    (dx_ref,) = torch.autograd.grad((y_ref * dout).sum(), [xr])
    xn = x.clone().requires_grad_(True)
    y = native(xn, 981.0, ctx).sample           # autograd seam
    (dx,) = torch.autograd.grad((y * dout).sum(), [xn])
    assert cosine(y, y_ref) >= 0.999 and rel_err(y, y_ref) < 5e-2
    assert cosine(dx, dx_ref) >= 0.999 and rel_err(dx, dx_ref) < 1e-1


def test_unet_checkpointed_backward_is_bit_identical(dev):
    """keep_activations (the default: saved state kept from the forward) and the forward re-run in the backward give the
    same bits, and a sample is independent of the batch it shares (fixed-order reductions)."""
    from tests.gpu_check_unet import make_oracle, tiny_native_config
    from tml_image_editing_defense_b200.unet import UNet2DConditionModel as NativeUNet
    cfg = tiny_native_config()
    sd = make_oracle(cfg, 4, 2.0).state_dict()
    g = torch.Generator().manual_seed(9)
    x = torch.randn((2, 4, 32, 32), generator=g).to(dev)
    ctx = torch.randn((2, 7, cfg.cross_attention_dim), generator=g).to(dev)
    dout = torch.randn((2, 4, 32, 32), generator=g).to(dev)
    res = []
    for keep in (False, True):
        native = NativeUNet(cfg, device=str(dev), keep_activations=keep).load_state_dict(sd)
        xn = x.clone().requires_grad_(True)
        y = native(xn, 500, ctx).sample
        (dx,) = torch.autograd.grad((y * dout).sum(), [xn])
        res.append((y.detach().clone(), dx.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    native = NativeUNet(cfg, device=str(dev)).load_state_dict(sd)
    x1 = x[1:2].clone().requires_grad_(True)
    y1 = native(x1, 500, ctx[1:2]).sample
    (dx1,) = torch.autograd.grad((y1 * dout[1:2]).sum(), [x1])
    assert torch.equal(y1.detach(), res[0][0][1:2]) and torch.equal(dx1, res[0][1][1:2])


def test_unet_sd15_vs_oracle(dev):
    """The SD-1.5 topology itself (320/640/1280/1280 channels, 10..80 channels per GroupNorm group, heads of 40/80/160
    channels, 77 prompt tokens) at 64 x 64 latents (512^2 images), batch 2 = one image under classifier-free guidance."""
    from tests.gpu_check_unet import run_unet
    ok, cy, cdx = run_unet(dev, "sd15", batch=2, size=64, tokens=77, verbose=False)
    assert ok
    assert cy >= 0.999 and cdx >= 0.999


def test_diffusion_attack_native_unet_vs_oracle(dev):
    """The reference's compute_grad (main.py:144-246) with EVERY network on this repo's kernels -- encoder, 4 DDIM steps
    of the UNet under classifier-free guidance, decoder, image losses -- against the all-fp32 PyTorch pipeline."""
    from oracle.decoder_oracle import make_vae_oracle
    from tests.gpu_check_unet import make_oracle
    from tml_image_editing_defense_b200.configs import TrainConfig
    from tml_image_editing_defense_b200.diffusion import DiffusionAttack
    from tml_image_editing_defense_b200.schedulers import DDIMScheduler
    from tml_image_editing_defense_b200.unet import UNet2DConditionModel as NativeUNet
    from tml_image_editing_defense_b200.unet_torch import UNetConfig
    from tml_image_editing_defense_b200.vae import AutoencoderKL
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ucfg = UNetConfig(block_out_channels=(64, 128), cross_attention_dim=64, attention_head_dim=8, norm_num_groups=32,
                      down_has_attn=(True, False), up_has_attn=(False, True))
    um = make_oracle(ucfg, 21, 2.0)
    vm = make_vae_oracle(0)
    vae = AutoencoderKL(device=str(dev)).load_state_dict(vm.state_dict())
    native = NativeUNet(ucfg, device=str(dev)).load_state_dict(um.state_dict())
    cfg = TrainConfig(norm_type="linf", override_from_norm_type=False, device=str(dev), apply_loss_on_images=True,
                      apply_loss_on_latents=False, perturbation_loss_lambda=1.0, n_denoising_steps_per_iteration=4)
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(2, 3, 128, 128, generator=g) * 2 - 1).to(dev)
    tgt = (torch.rand(2, 3, 128, 128, generator=g) * 2 - 1).to(dev)
    pe = torch.randn(2, 7, 64, generator=g).to(dev)
    nz = [torch.randn(2, 4, 16, 16, generator=g).to(dev)]
    ref = DiffusionAttack(cfg, vm.to(dev), um.to(dev), DDIMScheduler(), use_checkpointing=False, unet_dtype=torch.float32)
    g_ref, l_ref, img_ref, _ = ref.compute_grad(x, pe, x, tgt, None, nz)
    ours = DiffusionAttack(cfg, vae, native, DDIMScheduler(), use_checkpointing=False, unet_dtype=torch.float32)
    gg, l, img, _ = ours.compute_grad(x, pe, x, tgt, None, nz)
    c = cosine(gg, g_ref)
    print("native diffusion attack gradient cosine", c)
    assert c >= 0.99
    assert abs(float(l) - float(l_ref)) / float(l_ref) < 0.03
    assert rel_err(img, img_ref) < 0.08


def test_unfused_attention_path_agrees(dev):
    """TML_NO_FUSED_ATTN / TML_NO_FUSED_ATTN_BWD route the self attention through the GEMM-epilogue path (logits, P~ and
    dS materialised; the verified baseline the fused kernels were developed against): same stage-wise parity."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    env = dict(os.environ, TML_NO_FUSED_ATTN="1", TML_NO_FUSED_ATTN_BWD="1")
    r = subprocess.run([sys.executable, str(root / "tests" / "gpu_check_unet.py"), "--which", "tiny"], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ALL OK" in r.stdout


def test_transposed_a_operand_gemm(dev):
    """GemmOp::a_trans: A stored [batch][k][m] and read as an MN-major UMMA operand (dK = dS^T Q, dV = P^T dO without a
    transposed copy) against a plain matmul, for full 128-row tiles and for a 64-row tail tile."""
    import ctypes as C
    from tml_image_editing_defense_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    for (nb, tq, tk, dp) in ((3, 256, 384, 64), (2, 64, 128, 128)):
        A = torch.randn((nb, tk, tq), generator=g).to(dev).bfloat16()          # [batch][k][m]
        Bt = torch.randn((nb, dp, tk), generator=g).to(dev).bfloat16()         # [batch][n][k]
        D = torch.zeros((nb, tq, dp), device=dev, dtype=torch.bfloat16)
        gw = 128 if tq % 128 == 0 else 64
        d = _lib.TmlGemmDesc()
        d.A, d.A_C, d.A_W, d.A_H, d.A_B = A.data_ptr(), tk, gw, tq // gw, nb
        d.A_sW, d.A_sH, d.A_sB = tk, gw * tk, tq * tk
        d.a_trans, d.A_sK = 1, tq
        d.stride, d.ntaps = 1, 1
        d.OW, d.OH = gw, tq // gw
        d.Bm, d.N, d.B_sN, d.B_sBatch = Bt.data_ptr(), dp, tk, dp * tk
        d.alpha = 1.0
        d.D, d.D_sW, d.D_sH, d.D_sB, d.D_sN = D.data_ptr(), dp, gw * dp, tq * dp, 1
        _lib.check(lib.tml_debug_gemm(C.byref(d), torch.cuda.current_stream(dev).cuda_stream))
        torch.cuda.synchronize()
        ref = torch.einsum("bkm,bnk->bmn", A.float(), Bt.float())
        assert rel_err(D.float(), ref) < 1e-2, (nb, tq, tk, dp, rel_err(D.float(), ref))


def _fused_attention(dev, q, k, v, dout, d, scale, cross=False):
    """q [nb,tq,dp], k / v [nb,tkv,dp] bf16 with channels >= d zero (v[:, :, d] = 1): runs the fused kernels through
    tml_debug_attention and returns (O, dQ, dK, dV) as fp32 restricted to the d real channels."""
    import ctypes as C
    from tml_image_editing_defense_b200 import _lib
    lib = _lib.load()
    nb, tq, dp = q.shape
    tkv = k.shape[1]
    O = torch.zeros_like(q)
    rmax = torch.zeros((nb, tq), device=dev)
    inv_l = torch.zeros((nb, tq), device=dev)
    dQ, dK, dV = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    ws = torch.zeros(nb * tq * (2 * dp + 12) + 1024, dtype=torch.uint8, device=dev)
    _lib.check(lib.tml_debug_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), nb, tq, tkv, dp, d, scale, O.data_ptr(),
                                       rmax.data_ptr(), inv_l.data_ptr(), dout.data_ptr(), dQ.data_ptr(),
                                       None if cross else dK.data_ptr(), None if cross else dV.data_ptr(), ws.data_ptr(),
                                       ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize()
    return O.float()[..., :d], dQ.float()[..., :d], dK.float()[..., :d], dV.float()[..., :d]


@pytest.mark.parametrize("case", ["random", "rising_maxima", "masked_keys"])
def test_fused_attention_kernels_vs_fp32(dev, case):
    """mh_attn_fwd_online_kernel / attn_bwd_dq_kernel / attn_bwd_dkv_kernel on their own against fp32 softmax attention and
    its autograd: random logits; logits that keep growing along the keys (the online softmax has to rescale its TMEM
    accumulator again and again); and the cross-attention layout (77 real keys in 128 slots, masked through the spare
    channel, no dK / dV).  Tolerances: bf16 operands and probabilities, fp32 accumulation."""
    g = torch.Generator().manual_seed(11)
    nb, tq, d, dp = 6, 256, 40, 64
    tkv = 128 if case == "masked_keys" else 512
    scale = d ** -0.5
    q = torch.zeros((nb, tq, dp))
    k = torch.zeros((nb, tkv, dp))
    v = torch.zeros((nb, tkv, dp))
    q[..., :d] = torch.randn((nb, tq, d), generator=g) * 1.5
    k[..., :d] = torch.randn((nb, tkv, d), generator=g) * 1.5
    v[..., :d] = torch.randn((nb, tkv, d), generator=g)
    if case == "rising_maxima":
        # logits grow by ~0.1 per key along a common direction (+9 in log2 units per 64-key tile: the exponent reference
        # moves on every tile); steeper ramps make the rows near one-hot, where the cancellation dP - D limits ANY bf16
        # backward (0.4 per key: forward still within 2e-2, dQ off by 30 %)
        u = torch.randn((d,), generator=g)
        u = u / u.norm()
        q[..., :d] = q[..., :d] * 0.3 + 4.0 * u
        k[..., :d] = k[..., :d] * 0.3 + (torch.arange(tkv).float()[None, :, None] * (0.1 / (4.0 * scale))) * u
    valid = tkv
    if case == "masked_keys":
        valid = 77
        k[:, valid:, :] = 0
        v[:, valid:, :] = 0
        q[..., d] = 1.0
        k[:, valid:, d] = -29952.0
    v[..., d] = 1.0
    dout = torch.zeros((nb, tq, dp))
    dout[..., :d] = torch.randn((nb, tq, d), generator=g)
    qb, kb, vb, db = (t.to(dev).bfloat16() for t in (q, k, v, dout))
    O, dQ, dK, dV = _fused_attention(dev, qb, kb, vb, db, d, scale, cross=case == "masked_keys")
    qr = qb.float()[..., :d].clone().requires_grad_(True)
    kr = kb.float()[:, :valid, :d].clone().requires_grad_(True)
    vr = vb.float()[:, :valid, :d].clone().requires_grad_(True)
    P = torch.softmax(torch.einsum("bqd,bkd->bqk", qr, kr) * scale, dim=-1)
    Oref = torch.einsum("bqk,bkd->bqd", P, vr)
    gq, gk, gv = torch.autograd.grad((Oref * db.float()[..., :d]).sum(), [qr, kr, vr])
    assert not torch.isnan(O).any()
    assert rel_err(O, Oref.detach()) < 2e-2, rel_err(O, Oref.detach())
    tol = 1e-1 if case == "rising_maxima" else 4e-2
    assert rel_err(dQ, gq) < tol, rel_err(dQ, gq)
    if case != "masked_keys":
        assert rel_err(dK, gk) < tol, rel_err(dK, gk)
        assert rel_err(dV, gv) < tol, rel_err(dV, gv)


def test_diffusion_attack_pgd_loop_native(dev):
    """DiffusionAttack.run = the reference's PGD loop (main.py:79-135) around the full attack on native kernels: the
    iterate stays in the eps-ball and the image range, the loss history is finite, and the first update is the fused
    linf step of the first gradient."""
    from oracle.decoder_oracle import make_vae_oracle
    from tests.gpu_check_unet import make_oracle
    from tml_image_editing_defense_b200 import ops
    from tml_image_editing_defense_b200.configs import TrainConfig
    from tml_image_editing_defense_b200.diffusion import DiffusionAttack
    from tml_image_editing_defense_b200.schedulers import DDIMScheduler
    from tml_image_editing_defense_b200.unet import UNet2DConditionModel as NativeUNet
    from tml_image_editing_defense_b200.unet_torch import UNetConfig
    from tml_image_editing_defense_b200.vae import AutoencoderKL
    ucfg = UNetConfig(block_out_channels=(64, 128), cross_attention_dim=64, attention_head_dim=8, norm_num_groups=32,
                      down_has_attn=(True, False), up_has_attn=(False, True))
    native = NativeUNet(ucfg, device=str(dev)).load_state_dict(make_oracle(ucfg, 21, 2.0).state_dict())
    vae = AutoencoderKL(device=str(dev)).load_state_dict(make_vae_oracle(0).state_dict())
    eps, step = 8 / 255, 2 / 255
    cfg = TrainConfig(norm_type="linf", eps=eps, step_size=step, grad_reps=1, override_from_norm_type=False,
                      n_optimization_steps=3, device=str(dev), apply_loss_on_images=True, apply_loss_on_latents=False,
                      perturbation_loss_lambda=1.0, n_denoising_steps_per_iteration=2, n_noise=1)
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(2, 3, 128, 128, generator=g) * 1.6 - 0.8).to(dev)
    tgt = (torch.rand(1, 3, 128, 128, generator=g) * 2 - 1).to(dev)
    pe = torch.randn(2, 7, 64, generator=g).to(dev)
    nz = [torch.randn(2, 4, 16, 16, generator=g).to(dev)]
    da = DiffusionAttack(cfg, vae, native, DDIMScheduler(), use_checkpointing=False, unet_dtype=torch.float32)
    x_adv = da.run(x, tgt, pe, noises=nz)
    assert len(da.loss_history) == 3 and all(l == l and abs(l) < 1e6 for l in da.loss_history)
    assert float((x_adv - x).abs().max()) <= eps + 1e-6
    assert float(x_adv.min()) >= -1.0 and float(x_adv.max()) <= 1.0
    assert float((x_adv - x).abs().max()) > 0.5 * step
    g0, _, _, _ = da.compute_grad(x.clone(), pe, x, tgt.expand(2, -1, -1, -1).contiguous(), None, nz)
    first = ops.pgd_step_linf_(x.clone(), g0.contiguous(), x, eps, step, -1.0, 1.0)
    import copy
    cfg1 = copy.copy(cfg)
    cfg1.n_optimization_steps = 1
    da1 = DiffusionAttack(cfg1, vae, native, DDIMScheduler(), use_checkpointing=False, unet_dtype=torch.float32)
    assert torch.equal(da1.run(x, tgt, pe, noises=nz), first)
