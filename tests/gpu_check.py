"""GPU diagnostic: layer-by-layer comparison of the CUDA path against the oracle (run on a B200).

    python tests/gpu_check.py [--res 64] [--batch 2] [--impl tc|simt|both]

Prints one line per check and never stops at the first failure; meant for gpurun round trips where
one call must localise a bug.  Not part of the product path."""
import argparse
import ctypes as C
import sys
import time
import traceback
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from tests.helpers import (BACKWARD_STAGE_NAMES, cosine, emulate_gemm, fetch_saved, oracle_trace, pack_conv3x3,  # noqa
                           rel_err)
from tml_image_editing_defense_b200 import _lib, ops  # noqa


IMPL = {"simt": False}


def gemm_case(name, lib, B, H, W, Cin, N, mode, dev, stride=1, bias=False, resid=False, alpha=1.0, seed=0, gn=0, xf=False):
    """conv3x3 in packing mode `mode` (0 fwd s1, 1 dgrad s1, 2 fwd s2, 3..6 dgrad s2 parity) or mode -1: 1x1."""
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16)
    if mode == -1:
        Wm = (torch.randn(N, Cin, generator=g) / Cin ** 0.5).to(torch.bfloat16).float()
        dh, dw = [0], [0]
        OH, OW = H, W
    else:
        co, ci = (N, Cin) if mode in (0, 2) else (Cin, N)
        w = (torch.randn(co, ci, 3, 3, generator=g) / (9 * Cin) ** 0.5).to(torch.bfloat16).float()
        Wm, dh, dw = pack_conv3x3(w, mode)
        OH, OW = (H // 2, W // 2) if mode == 2 else (H, W)
    A_ref = A.float()
    ss_in = None
    if xf:   # fused input normalisation: the kernel convolves bf16(silu(A * scale + shift)), A = the raw GroupNorm input
        ss_in = torch.stack([torch.rand(B, Cin, generator=g) + 0.5, torch.randn(B, Cin, generator=g) * 0.5], dim=-1).contiguous()
        u = A.float() * ss_in[:, None, None, :, 0] + ss_in[:, None, None, :, 1]
        A_ref = (u * torch.sigmoid(u)).to(torch.bfloat16).float()
    ref = emulate_gemm(A_ref, Wm, dh, dw, stride, OH, OW) * alpha
    bias_t = torch.randn(N, generator=g) if bias else None
    resid_t = torch.randn(B, OH, OW, N, generator=g).to(torch.bfloat16) if resid else None
    if bias:
        ref = ref + bias_t.double()
    if resid:
        ref = ref + resid_t.double()
    Ad, Wd = A.to(dev), Wm.to(torch.bfloat16).to(dev)
    # output with a guard band on both sides: a stray store (wrong tile decode, junk rows of the pitch-66 mode written)
    # shows up as a changed sentinel even when the values inside happen to be right
    GUARD = 8192
    Dbig = torch.full((GUARD + B * OH * OW * N + GUARD,), -777.0, dtype=torch.bfloat16, device=dev)
    D = Dbig[GUARD: GUARD + B * OH * OW * N].view(B, OH, OW, N)
    D.fill_(float("nan"))
    d = _lib.TmlGemmDesc()
    d.A = Ad.data_ptr(); d.A_C = Cin; d.A_W = W; d.A_H = H; d.A_B = B
    d.A_sW = Cin; d.A_sH = W * Cin; d.A_sB = H * W * Cin
    d.stride = stride; d.ntaps = len(dh)
    for t in range(len(dh)):
        d.dh[t] = dh[t]; d.dw[t] = dw[t]
    d.OW = OW; d.OH = OH
    d.Bm = Wd.data_ptr(); d.N = N; d.B_sN = Wm.shape[1]; d.B_sBatch = 0
    d.alpha = alpha
    bd = bias_t.to(dev) if bias else None
    rd = resid_t.to(dev) if resid else None
    d.bias = bd.data_ptr() if bias else None
    d.resid = rd.data_ptr() if resid else None
    d.R_sW = N; d.R_sH = OW * N; d.R_sB = OH * OW * N
    d.D = D.data_ptr(); d.out_fp32 = 0
    d.D_sW = N; d.D_sH = OW * N; d.D_sB = OH * OW * N; d.D_sN = 1; d.n_store = 0
    keep = []
    if xf:
        keep.append(ss_in.to(dev))
        d.in_gn_ss = keep[-1].data_ptr()
    if gn and IMPL["simt"]:
        gn = 0   # the SIMT debug kernel has no fused reductions
    if gn:
        d.gn_mode = gn
        if gn == 2:
            xg = torch.randn(B, OH, OW, N, generator=g).to(torch.bfloat16)
            ss = torch.stack([torch.rand(B, N, generator=g) + 0.5, torch.randn(B, N, generator=g) * 0.3], dim=-1)
            mr = torch.stack([torch.randn(B, 32, generator=g) * 0.2, torch.rand(B, 32, generator=g) + 0.5], dim=-1)
            gam = torch.randn(N, generator=g)
            keep = [xg.to(dev), ss.contiguous().to(dev), mr.contiguous().to(dev), gam.to(dev)]
            d.gn_x, d.gn_ss, d.gn_mr, d.gn_gamma = [t.data_ptr() for t in keep]
            d.gn_silu = 1
        ntile = lib.tml_debug_gn_chunks_per_image(C.byref(d))
        part = torch.full((B, ntile, 32, 2), float("nan"), dtype=torch.float32, device=dev)
        d.gn_partial = part.data_ptr()
    rc = lib.tml_debug_gemm(C.byref(d), torch.cuda.current_stream().cuda_stream)
    if rc:
        print(f"[gemm] {name}: launch error {lib.tml_last_error().decode()}")
        return False
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print(f"[gemm] {name}: kernel failed ({e}); last hang id = {lib.tml_debug_last_hang()}")
        raise
    out = D.float().cpu()
    err = rel_err(out, ref)
    nan = int(torch.isnan(out).sum())
    guard_ok = bool((Dbig[:GUARD] == -777.0).all()) and bool((Dbig[-GUARD:] == -777.0).all())
    if not guard_ok:
        print(f"        {name}: stores outside the output tensor (guard band overwritten)")
    ok = err < 1e-2 and nan == 0 and guard_ok
    if gn:
        # reference reductions over the bf16 values the kernel stored
        o64 = out.double().view(B, OH * OW, 32, N // 32)
        got = part.double().sum(dim=1).cpu()       # [B, 32, 2]
        if gn == 1:
            want = torch.stack([o64.sum(dim=(1, 3)), (o64 * o64).sum(dim=(1, 3))], dim=-1)
        else:
            x64 = xg.double().view(B, OH * OW, 32, N // 32)
            sc = ss[..., 0].double().view(B, 1, 32, N // 32)
            sh = ss[..., 1].double().view(B, 1, 32, N // 32)
            u = x64 * sc + sh
            sg = torch.sigmoid(u)
            dxh = o64 * (sg * (1 + u * (1 - sg))) * gam.double().view(1, 1, 32, N // 32)
            xh = (x64 - mr[..., 0].double().view(B, 1, 32, 1)) * mr[..., 1].double().view(B, 1, 32, 1)
            want = torch.stack([dxh.sum(dim=(1, 3)), (dxh * xh).sum(dim=(1, 3))], dim=-1)
        gerr = float((got - want).abs().max() / (want.abs().max() + 1e-30))
        print(f"        fused GN mode {gn}: max rel dev of group sums = {gerr:.3e} {'OK' if gerr < 2e-3 else 'FAIL'}")
        ok = ok and gerr < 2e-3
    print(f"[gemm] {name:34s} B={B} {H}x{W} C={Cin} N={N} taps={len(dh)} s={stride}: rel_err={err:.3e} nan={nan} "
          f"{'OK' if ok else 'FAIL'}")
    if not ok:
        diff = (out.double() - ref).abs().view(-1, N)
        rows = diff.max(dim=1).values
        bad = torch.nonzero(rows > 0.05 * ref.abs().max()).flatten()
        print(f"        bad rows {bad.numel()}/{rows.numel()} first {bad[:16].tolist()}  max|ref|={float(ref.abs().max()):.3f}")
        cols = diff.max(dim=0).values
        badc = torch.nonzero(cols > 0.05 * ref.abs().max()).flatten()
        print(f"        bad cols {badc.numel()}/{N} first {badc[:16].tolist()}")
    return ok


def run_gemm_suite(lib, dev):
    ok = True
    cases = [
        ("1x1 M128 K64", dict(B=1, H=8, W=16, Cin=64, N=64, mode=-1)),
        ("1x1 M128 K128 N128", dict(B=1, H=8, W=16, Cin=128, N=128, mode=-1)),
        ("1x1 K512 N256 bias", dict(B=2, H=8, W=16, Cin=512, N=256, mode=-1, bias=True)),
        ("1x1 N512 (2 n-tiles) resid", dict(B=2, H=16, W=16, Cin=256, N=512, mode=-1, resid=True, bias=True)),
        ("1x1 N16", dict(B=1, H=8, W=8, Cin=64, N=16, mode=-1)),
        ("1x1 N1536", dict(B=1, H=8, W=8, Cin=128, N=1536, mode=-1)),
        ("1x1 alpha", dict(B=1, H=8, W=8, Cin=128, N=64, mode=-1, alpha=0.125)),
        # lean per-warp TMA-store epilogue: multi-wave, bias + residual, narrow tiles (TW < 32), odd column-chunk counts
        ("lean 1x1 128->256 64x64 B=8 bias resid", dict(B=8, H=64, W=64, Cin=128, N=256, mode=-1, bias=True, resid=True)),
        ("lean 1x1 256->128 64x64 B=12 resid (mt=2)", dict(B=12, H=64, W=64, Cin=256, N=128, mode=-1, resid=True)),
        ("lean 1x1 W=8 bias", dict(B=3, H=8, W=8, Cin=128, N=128, mode=-1, bias=True)),
        ("lean 1x1 W=16 N=96 alpha", dict(B=2, H=16, W=16, Cin=64, N=96, mode=-1, bias=True, alpha=0.5)),
        ("lean 1x1 W=24 (register fallback)", dict(B=2, H=8, W=24, Cin=64, N=64, mode=-1, bias=True)),
        ("conv3x3 fwd 128->128", dict(B=2, H=16, W=16, Cin=128, N=128, mode=0, bias=True)),
        ("conv3x3 fwd 128->256 W=128", dict(B=1, H=4, W=128, Cin=128, N=256, mode=0)),
        ("conv3x3 fwd 512->512 8x8", dict(B=2, H=8, W=8, Cin=512, N=512, mode=0, resid=True)),
        ("conv3x3 dgrad 256->128", dict(B=2, H=16, W=16, Cin=256, N=128, mode=1)),
        ("conv3x3 s2 fwd 128", dict(B=2, H=32, W=32, Cin=128, N=128, mode=2, stride=2, bias=True)),
        ("conv3x3 s2 fwd 256 64->32", dict(B=1, H=64, W=64, Cin=256, N=256, mode=2, stride=2)),
        ("s2 dgrad parity 00", dict(B=2, H=16, W=16, Cin=128, N=128, mode=3)),
        ("s2 dgrad parity 01", dict(B=2, H=16, W=16, Cin=128, N=128, mode=4)),
        ("s2 dgrad parity 10", dict(B=2, H=16, W=16, Cin=128, N=128, mode=5)),
        ("s2 dgrad parity 11", dict(B=2, H=16, W=16, Cin=128, N=128, mode=6)),
        ("multi-wave 128->128 64x64 B=8", dict(B=8, H=64, W=64, Cin=128, N=128, mode=0)),
        ("mt=2 128->128 64x64 B=12 +stats", dict(B=12, H=64, W=64, Cin=128, N=128, mode=0, bias=True, gn=1)),
        ("mt=2 s2 128 128->64 B=12", dict(B=12, H=128, W=128, Cin=128, N=128, mode=2, stride=2, bias=True)),
        ("mt=2 dgrad 256->128 B=12 +gnbwd", dict(B=12, H=64, W=64, Cin=256, N=128, mode=1, gn=2)),
        ("mt=2 parity 11 B=12", dict(B=12, H=64, W=64, Cin=128, N=128, mode=6)),
        ("stats N=256 resid", dict(B=2, H=32, W=32, Cin=128, N=256, mode=0, resid=True, gn=1)),
        ("stats N=512 2 n-tiles", dict(B=2, H=16, W=16, Cin=256, N=512, mode=0, bias=True, gn=1)),
        ("gnbwd N=512 8x8", dict(B=2, H=8, W=8, Cin=512, N=512, mode=1, gn=2)),
        ("gnbwd N=256", dict(B=3, H=32, W=32, Cin=256, N=256, mode=1, gn=2)),
        ("halo mt2 128->128 8x128 +stats", dict(B=2, H=8, W=128, Cin=128, N=128, mode=0, bias=True, gn=1)),
        ("halo mt2 dgrad 256->128 4x256 +gnbwd", dict(B=2, H=4, W=256, Cin=256, N=128, mode=1, gn=2)),
        ("halo mt1 512->512 3x128 resid", dict(B=1, H=3, W=128, Cin=512, N=512, mode=0, resid=True)),
        ("halo mt1 256->256 5x128 +stats", dict(B=2, H=5, W=128, Cin=256, N=256, mode=0, bias=True, gn=1)),
        ("halo N=16 dgrad 4x128", dict(B=1, H=4, W=128, Cin=128, N=16, mode=1)),
        ("halo pair 512->512 4x128 2 n-tiles", dict(B=2, H=4, W=128, Cin=512, N=512, mode=0, resid=True, bias=True, gn=1)),
        ("halo pair N=16 dgrad 8x256", dict(B=2, H=8, W=256, Cin=128, N=16, mode=1)),
        ("halo multi-wave 128->128 128x128 B=4", dict(B=4, H=128, W=128, Cin=128, N=128, mode=0, resid=True)),
        # rows of 64 pixels: pitch-66 halo tile, 33 M tiles of 128 slots per 64 rows (pairs = same tile of two images)
        ("h66 pair 512->512 64x64 B=2 all", dict(B=2, H=64, W=64, Cin=512, N=512, mode=0, bias=True, resid=True, gn=1)),
        ("h66 pair dgrad 512->512 64x64 B=4 +gnbwd", dict(B=4, H=64, W=64, Cin=512, N=512, mode=1, gn=2)),
        ("h66 single 256->256 64x64 B=3 +stats", dict(B=3, H=64, W=64, Cin=256, N=256, mode=0, bias=True, gn=1)),
        ("h66 single 128->128 128x64 B=1 resid", dict(B=1, H=128, W=64, Cin=128, N=128, mode=0, resid=True)),
        ("h66 pair 64->256 64x64 B=2", dict(B=2, H=64, W=64, Cin=64, N=256, mode=0)),
        # operand-swapped kernel (N = 128 output channels, rows of 256 pixels)
        ("swap 128->128 4x256 bias +stats", dict(B=2, H=4, W=256, Cin=128, N=128, mode=0, bias=True, gn=1)),
        ("swap dgrad 128->128 3x512 resid", dict(B=1, H=3, W=512, Cin=128, N=128, mode=1, resid=True)),
        ("swap dgrad 256->128 2x256", dict(B=2, H=2, W=256, Cin=256, N=128, mode=1)),
        ("swap 64->128 1x256", dict(B=1, H=1, W=256, Cin=64, N=128, mode=0, bias=True)),
        ("swap 128->256 3x256 bias +stats", dict(B=2, H=3, W=256, Cin=128, N=256, mode=0, bias=True, gn=1)),
        ("swap 256->256 2x256 resid +stats", dict(B=1, H=2, W=256, Cin=256, N=256, mode=0, resid=True, gn=1)),
        ("swap dgrad 256->256 2x512 +gnbwd", dict(B=2, H=2, W=512, Cin=256, N=256, mode=1, gn=2)),
        ("swap dgrad 128->128 4x256 +gnbwd", dict(B=2, H=4, W=256, Cin=128, N=128, mode=1, gn=2)),
        ("swap 128->512 2x256 +stats", dict(B=1, H=2, W=256, Cin=128, N=512, mode=0, bias=True, gn=1)),
        ("swap dgrad 512->512 1x256 +gnbwd", dict(B=2, H=1, W=256, Cin=512, N=512, mode=1, gn=2)),
        # CTA pairs with the GroupNorm + SiLU of the input applied on the operand path (raw input, per-image scale / shift)
        ("swpair xf 256->256 4x128 bias +stats", dict(B=2, H=4, W=128, Cin=256, N=256, mode=0, bias=True, gn=1, xf=True)),
        ("swpair xf 128->256 6x256 (two segments)", dict(B=2, H=6, W=256, Cin=128, N=256, mode=0, bias=True, xf=True)),
        ("swpair xf 512->512 2x128 resid", dict(B=1, H=2, W=128, Cin=512, N=512, mode=0, resid=True, xf=True)),
        ("swpair xf multi-wave 256->512 64x128 B=4 all", dict(B=4, H=64, W=128, Cin=256, N=512, mode=0, bias=True, resid=True, gn=1, xf=True)),
        # the same kernel on CTA pairs (rows of 128 pixels: M = 256 channels, N = 2 rows x 128 pixels)
        ("swpair 256->256 4x128 bias +stats", dict(B=2, H=4, W=128, Cin=256, N=256, mode=0, bias=True, gn=1)),
        ("swpair 512->512 2x128 resid", dict(B=1, H=2, W=128, Cin=512, N=512, mode=0, resid=True)),
        ("swpair dgrad 512->512 4x128 +gnbwd", dict(B=2, H=4, W=128, Cin=512, N=512, mode=1, gn=2)),
        ("swpair dgrad 512->256 2x128", dict(B=2, H=2, W=128, Cin=512, N=256, mode=1)),
        ("swpair multi-wave 256->512 64x128 B=4 all", dict(B=4, H=64, W=128, Cin=256, N=512, mode=0, bias=True, resid=True, gn=1)),
        ("swap multi-wave 128->128 64x256 B=4 all", dict(B=4, H=64, W=256, Cin=128, N=128, mode=0, bias=True, resid=True, gn=1)),
    ]
    import os
    no_pairs = any(os.environ.get(k) == v for k, v in (("TML_NO_SWAP", "1"), ("TML_NO_SWAP_PAIR", "1"),
                                                        ("TML_SWAP_PREFER_PAIR", "0")))
    for name, kw in cases:
        if kw.get("xf") and (no_pairs or IMPL["simt"]):
            continue   # the fused input normalisation exists on the CTA-pair form of the swapped kernel only
        try:
            ok &= gemm_case(name, lib, dev=dev, **kw)
        except Exception:
            ok = False
            print(f"[gemm] {name}: EXCEPTION")
            traceback.print_exc()
    return ok


def run_encoder(dev, res, batch, kind, label):
    from oracle.encoder_oracle import make_oracle, perturb_affine_params
    from tml_image_editing_defense_b200.vae import AutoencoderKL
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = make_oracle(0)
    perturb_affine_params(model, 1234)
    g = torch.Generator().manual_seed(100 + res)
    x = torch.rand((batch, 3, res, res), generator=g) * 2 - 1
    tgt = torch.randn((batch, 4, res // 8, res // 8), generator=g)
    noise = torch.randn((batch, 4, res // 8, res // 8), generator=g)
    model_d = model.to(dev)
    acts, grads, mom_ref, gx_ref, loss_ref = oracle_trace(model_d, x.to(dev), tgt.to(dev), noise.to(dev), kind)

    vae = AutoencoderKL(device=str(dev)).load_state_dict(model.state_dict())
    xd = x.to(dev)
    mom, saved = vae._forward_raw(xd, keep=True)
    torch.cuda.synchronize()
    ok = True

    def report(name, a, b, tol):
        nonlocal ok
        e, c = rel_err(a, b), cosine(a, b)
        good = e < tol and not torch.isnan(a).any()
        ok &= bool(good)
        print(f"[{label}] {name:12s} rel_err={e:.3e} cos={c:.6f} {'OK' if good else 'FAIL'}")

    report("conv_in", fetch_saved(vae, saved, "conv_in"), acts["conv_in"], 2e-2)
    for i in range(10):
        report(f"res{i}_h1", fetch_saved(vae, saved, "resnet_h1", i), acts[f"res{i}_h1"], 3e-2)
        report(f"res{i}_out", fetch_saved(vae, saved, "resnet_out", i), acts[f"res{i}_out"], 3e-2)
        if i in (1, 3, 5):
            report(f"down{i // 2}_out", fetch_saved(vae, saved, "down_out", i // 2), acts[f"down{i // 2}_out"], 3e-2)
        if i == 8:
            report("attn_out", fetch_saved(vae, saved, "attn_out"), acts["attn_out"], 3e-2)
    report("moments", mom, mom_ref, 3e-2)

    # backward with stage dumps
    z, loss, dm = ops.latent_loss(mom, noise.to(dev), tgt.to(dev), kind)
    report("loss", loss, loss_ref, 2e-2)
    slot = batch * res * res * 128 * 2
    dump = torch.zeros(len(BACKWARD_STAGE_NAMES) * slot, dtype=torch.uint8, device=dev)
    vae._lib.tml_debug_set_grad_dump(dump.data_ptr(), slot, len(BACKWARD_STAGE_NAMES))
    gx = vae._backward_raw(dm, saved, tuple(x.shape))
    torch.cuda.synchronize()
    vae._lib.tml_debug_set_grad_dump(None, 0, 0)
    for k, name in enumerate(BACKWARD_STAGE_NAMES):
        ref = grads[name]
        Bn, Cc, Hh, Ww = ref.shape
        n = Bn * Cc * Hh * Ww
        t = dump[k * slot: k * slot + 2 * n].view(torch.bfloat16).view(Bn, Hh, Ww, Cc).float().permute(0, 3, 1, 2)
        report("d_" + name, t, ref, 6e-2)
    report("grad_x", gx, gx_ref, 8e-2)
    print(f"[{label}] grad_x cosine = {cosine(gx, gx_ref):.6f} (gate >= 0.999)")
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=64)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--impl", default="both")
    ap.add_argument("--skip-gemm", action="store_true")
    ap.add_argument("--gemm-only", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    impls = {"tc": [0], "simt": [1], "both": [1, 0]}[args.impl]
    allok = True
    for impl in impls:
        label = "tc" if impl == 0 else "simt"
        lib.tml_debug_set_gemm_impl(impl)
        IMPL["simt"] = impl == 1
        print(f"===== GEMM impl = {label} =====", flush=True)
        try:
            if not args.skip_gemm:
                allok &= run_gemm_suite(lib, dev)
            sys.stdout.flush()
            if not args.gemm_only:
                allok &= run_encoder(dev, args.res, args.batch, 0, label)
        except Exception:
            allok = False
            traceback.print_exc()
        sys.stdout.flush()
    lib.tml_debug_set_gemm_impl(0)
    print("ALL OK" if allok else "SOME CHECKS FAILED")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
