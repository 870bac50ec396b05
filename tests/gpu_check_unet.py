"""GPU diagnostic for the native UNet (csrc/unet.cu): stage-by-stage forward and backward comparison with the fp32
PyTorch restatement of diffusers' UNet2DConditionModel (``unet_torch.py`` -- the oracle of this path; diffusers itself
is absent, so parity of the module is unpinned like the VAE's: parameter count 859 520 964 and the key set only)."""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from tests.helpers import cosine, rel_err  # noqa: E402


def tiny_native_config():
    from tml_image_editing_defense_b200.unet_torch import UNetConfig
    # the native kernels need channel counts that are multiples of 64 and 32 GroupNorm groups
    return UNetConfig(block_out_channels=(64, 128, 128), cross_attention_dim=64, attention_head_dim=8, norm_num_groups=32,
                      down_has_attn=(True, True, False), up_has_attn=(False, True, True))


def make_oracle(cfg, seed=0, sharpen=4.0):
    """Random-init oracle with perturbed norm affine parameters and sharper attention logits (default init gives
    near-uniform softmax rows, which would hide errors in the attention products)."""
    from tml_image_editing_defense_b200.unet_torch import UNet2DConditionModel
    torch.manual_seed(seed)
    m = UNet2DConditionModel(cfg).float().eval().requires_grad_(False)
    g = torch.Generator().manual_seed(seed + 1)
    for name, p in m.named_parameters():
        if "norm" in name and name.endswith(".weight"):
            p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
        elif "norm" in name and name.endswith(".bias"):
            p.copy_(0.1 * torch.randn(p.shape, generator=g))
        elif name.endswith(("to_q.weight", "to_k.weight")):
            p.mul_(sharpen)
    return m


def module_sequence(m):
    """Modules in this repo's forward order: (kind, module) with kind in res / tf / down / up."""
    seq = []
    for blk in m.down_blocks:
        for i, r in enumerate(blk.resnets):
            seq.append(("res", r))
            if blk.attentions is not None:
                seq.append(("tf", blk.attentions[i]))
        if blk.downsamplers is not None:
            seq.append(("down", blk.downsamplers[0]))
    seq += [("res", m.mid_block.resnets[0]), ("tf", m.mid_block.attentions[0]), ("res", m.mid_block.resnets[1])]
    for blk in m.up_blocks:
        for i, r in enumerate(blk.resnets):
            seq.append(("res", r))
            if blk.attentions is not None:
                seq.append(("tf", blk.attentions[i]))
        if blk.upsamplers is not None:
            seq.append(("up", blk.upsamplers[0]))
    return seq


def oracle_trace(m, x, t, ctx, dout):
    """Forward outputs and per-module input gradients of the oracle, keyed by module."""
    outs, gins = {}, {}
    hooks = []
    seq = module_sequence(m)
    for kind, mod in seq + [("gn", m.conv_norm_out), ("conv_in", m.conv_in)]:
        hooks.append(mod.register_forward_hook(lambda mod_, inp, out, k=mod: outs.__setitem__(k, out.detach())))
        hooks.append(mod.register_full_backward_hook(lambda mod_, gi, go, k=mod: gins.__setitem__(k, gi[0].detach() if gi[0] is not None else None)))
    xr = x.clone().requires_grad_(True)
    with torch.enable_grad():
        y = m(xr, t, ctx).sample
        (dx,) = torch.autograd.grad((y * dout).sum(), [xr])
    for h in hooks:
        h.remove()
    return seq, outs, gins, y.detach(), dx.detach()


def run_unet(dev, which="tiny", batch=2, size=32, tokens=5, timestep=417.0, verbose=True, sharpen=2.0):
    from tml_image_editing_defense_b200.unet import UNet2DConditionModel as NativeUNet
    from tml_image_editing_defense_b200.unet_torch import UNetConfig
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = tiny_native_config() if which == "tiny" else UNetConfig()
    m = make_oracle(cfg, 0, sharpen)
    g = torch.Generator().manual_seed(7)
    x = torch.randn((batch, cfg.in_channels, size, size), generator=g)
    ctx = torch.randn((batch, tokens, cfg.cross_attention_dim), generator=g)
    dout = torch.randn((batch, cfg.out_channels, size, size), generator=g) * 1e-2
    native = NativeUNet(cfg, device=str(dev), keep_activations=True).load_state_dict(m.state_dict())
    md = m.to(dev)
    seq, outs, gins, y_ref, dx_ref = oracle_trace(md, x.to(dev), torch.tensor(timestep, device=dev), ctx.to(dev), dout.to(dev))
    y, saved = native._forward_raw(x.to(dev), timestep, ctx.to(dev), keep=True)
    torch.cuda.synchronize()
    ok = True

    def report(name, a, b, tol):
        nonlocal ok
        e, c = rel_err(a, b), cosine(a, b)
        good = e < tol and not torch.isnan(a.float()).any()
        ok &= bool(good)
        if verbose or not good:
            print(f"[unet] {name:16s} rel_err={e:.3e} cos={c:.6f} {'OK' if good else 'FAIL'}", flush=True)

    def nhwc(tn):
        return tn.float().permute(0, 3, 1, 2)

    report("conv_in", nhwc(native.saved_tensor(saved, "conv_in")), outs[md.conv_in], 2e-2)
    idx = {"res": 0, "tf": 0, "down": 0, "up": 0}
    names = {"res": "resnet_out", "tf": "tf_out", "down": "down_out", "up": "up_out"}
    for kind, mod in seq:
        k = idx[kind]
        idx[kind] += 1
        report(f"{kind}{k}_out", nhwc(native.saved_tensor(saved, names[kind], k)), outs[mod], 5e-2)
    report("out", y, y_ref, 5e-2)

    # backward: every stage dumps its output gradient in walk order
    order = [("gn", md.conv_norm_out)]
    ups = [b for b in md.up_blocks]
    for blk in reversed(ups):
        if blk.upsamplers is not None:
            order.append(("up", blk.upsamplers[0]))
        for i in reversed(range(len(blk.resnets))):
            if blk.attentions is not None:
                order.append(("tf", blk.attentions[i]))
            order.append(("res", blk.resnets[i]))
    order += [("res", md.mid_block.resnets[1]), ("tf", md.mid_block.attentions[0]), ("res", md.mid_block.resnets[0])]
    for blk in reversed(list(md.down_blocks)):
        if blk.downsamplers is not None:
            order.append(("down", blk.downsamplers[0]))
        for i in reversed(range(len(blk.resnets))):
            if blk.attentions is not None:
                order.append(("tf", blk.attentions[i]))
            order.append(("res", blk.resnets[i]))
    slot = max(int(gins[mod].numel()) for _, mod in order) * 2
    slot = (slot + 255) // 256 * 256
    dump = torch.zeros(len(order) * slot, dtype=torch.uint8, device=dev)
    native._lib.tml_debug_set_grad_dump(dump.data_ptr(), slot, len(order))
    dx = native._backward_raw(dout.to(dev), saved, tuple(x.shape), tokens)
    torch.cuda.synchronize()
    native._lib.tml_debug_set_grad_dump(None, 0, 0)
    for k, (kind, mod) in enumerate(order):
        ref = gins[mod]
        Bn, Cc, Hh, Ww = ref.shape
        n = Bn * Cc * Hh * Ww
        tn = dump[k * slot: k * slot + 2 * n].view(torch.bfloat16).view(Bn, Hh, Ww, Cc)
        report(f"d_{kind}@{k}", nhwc(tn), ref, 1e-1)
    report("dx", dx, dx_ref, 1e-1)
    cos_y, cos_dx = cosine(y, y_ref), cosine(dx, dx_ref)
    print(f"[unet] {which}: out cosine = {cos_y:.6f}, dsample cosine = {cos_dx:.6f}")
    return ok, cos_y, cos_dx


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="tiny", choices=["tiny", "sd15"])
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--tokens", type=int, default=5)
    ap.add_argument("--sharpen", type=float, default=2.0)
    a = ap.parse_args()
    ok, _, _ = run_unet(torch.device("cuda:0"), a.which, a.batch, a.size, a.tokens, sharpen=a.sharpen)
    print("ALL OK" if ok else "SOME CHECKS FAILED")
    sys.exit(0 if ok else 1)
