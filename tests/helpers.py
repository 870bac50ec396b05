"""Shared test helpers: a torch emulation of the implicit-GEMM operation the CUDA kernels execute
(same tap lists, same packed operands), used on CPU to validate packing and on GPU as the checker."""
import ctypes as C

import numpy as np
import torch


def bf16_bits_to_float(u16: np.ndarray) -> torch.Tensor:
    return torch.from_numpy((u16.astype(np.uint32) << 16).view(np.float32).copy())


def pack_conv3x3(w: torch.Tensor, mode: int):
    """Call the library's host-side packer (no CUDA).  Returns (matrix [N, ntaps*K] fp32, dh, dw)."""
    from tml_image_editing_defense_b200 import _lib
    lib = _lib.load()
    Co, Ci = w.shape[0], w.shape[1]
    wc = w.detach().float().contiguous()
    out = np.zeros(Co * Ci * 9, dtype=np.uint16)
    nt = C.c_int()
    dh = (C.c_int * 9)()
    dw = (C.c_int * 9)()
    _lib.check(lib.tml_debug_pack_conv3x3(wc.data_ptr(), Co, Ci, mode, out.ctypes.data, C.byref(nt), dh, dw))
    nt = nt.value
    N = Co if mode in (0, 2) else Ci
    K = Ci if mode in (0, 2) else Co
    mat = bf16_bits_to_float(out[: N * nt * K]).view(N, nt * K)
    return mat, list(dh[:nt]), list(dw[:nt])


def emulate_gemm(A: torch.Tensor, Bm: torch.Tensor, dh, dw, stride: int, OH: int, OW: int) -> torch.Tensor:
    """A: [B,H,W,C] ; Bm: [N, ntaps*C] ; returns [B,OH,OW,N] (float64 accumulate)."""
    B, H, W, Cc = A.shape
    N = Bm.shape[0]
    out = torch.zeros(B, OH, OW, N, dtype=torch.float64)
    A = A.double()
    Bm = Bm.double()
    for t, (a, b) in enumerate(zip(dh, dw)):
        ih = torch.arange(OH) * stride + a
        iw = torch.arange(OW) * stride + b
        vh = (ih >= 0) & (ih < H)
        vw = (iw >= 0) & (iw < W)
        g = torch.zeros(B, OH, OW, Cc, dtype=torch.float64)
        sub = A[:, ih[vh]][:, :, iw[vw]]
        idx_h = torch.nonzero(vh).flatten()
        idx_w = torch.nonzero(vw).flatten()
        g[:, idx_h[:, None], idx_w[None, :]] = sub
        out += g @ Bm[:, t * Cc:(t + 1) * Cc].T
    return out


def cosine(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


# ------------------------------------------------------------------------------------------------
# oracle tracing: intermediate activations and their gradients, by the names the CUDA path saves
# ------------------------------------------------------------------------------------------------
BACKWARD_STAGE_NAMES = ["res9_out", "attn_out", "res8_out", "res7_out", "res6_out", "down2_out", "res5_out",
                        "res4_out", "down1_out", "res3_out", "res2_out", "down0_out", "res1_out", "res0_out",
                        "conv_in"]


def oracle_resnets(model):
    enc = model.encoder
    out = []
    for blk in enc.down_blocks:
        out.extend(list(blk.resnets))
    out.extend(list(enc.mid_block.resnets))
    return out


def oracle_trace(model, x, target, noise, kind=0):
    """Run the oracle with hooks.  Returns (acts, grads, moments, grad_x, losses): dicts of NCHW tensors."""
    from oracle.encoder_oracle import latent_loss, DiagonalGaussianDistribution
    acts = {}
    handles = []

    def keep(name):
        def hook(mod, inp, out):
            out.retain_grad()
            acts[name] = out
        return hook

    enc = model.encoder
    handles.append(enc.conv_in.register_forward_hook(keep("conv_in")))
    for i, r in enumerate(oracle_resnets(model)):
        handles.append(r.conv1.register_forward_hook(keep(f"res{i}_h1")))
        handles.append(r.register_forward_hook(keep(f"res{i}_out")))
    for i, blk in enumerate(enc.down_blocks):
        if blk.downsamplers is not None:
            handles.append(blk.downsamplers[0].register_forward_hook(keep(f"down{i}_out")))
    if enc.mid_block.attentions is not None:
        handles.append(enc.mid_block.attentions[0].register_forward_hook(keep("attn_out")))
    with torch.enable_grad():
        xx = x.clone().requires_grad_(True)
        moments = model.moments(xx)
        dist = DiagonalGaussianDistribution(moments)
        z = dist.mean if noise is None else dist.mean + dist.std * noise
        losses = latent_loss(z, target, kind)
        losses.sum().backward()
    for h in handles:
        h.remove()
    grads = {k: v.grad.detach() for k, v in acts.items() if v.grad is not None}
    acts = {k: v.detach() for k, v in acts.items()}
    return acts, grads, moments.detach(), xx.grad.detach(), losses.detach()


def fetch_saved(vae, saved, name, index=0):
    """bf16 NHWC activation kept by the CUDA forward -> fp32 NCHW torch tensor."""
    from tml_image_editing_defense_b200 import _lib
    off = C.c_size_t()
    dims = (C.c_int * 4)()
    _lib.check(vae._lib.tml_debug_saved_tensor(vae._h, name.encode(), index, C.byref(off), dims))
    B, H, W, Cc = list(dims)
    n = B * H * W * Cc
    t = saved[off.value: off.value + 2 * n].view(torch.bfloat16).view(B, H, W, Cc)
    return t.float().permute(0, 3, 1, 2).contiguous()


# ------------------------------------------------------------------------------------------------
# decoder tracing
# ------------------------------------------------------------------------------------------------
DEC_BACKWARD_STAGE_NAMES = ["res13_out", "res12_out", "res11_out", "up2_out", "res10_out", "res9_out", "res8_out",
                            "up1_out", "res7_out", "res6_out", "res5_out", "up0_out", "res4_out", "res3_out",
                            "res2_out", "res1_out", "attn_out", "res0_out", "conv_in"]


def oracle_decoder_resnets(model):
    dec = model.decoder
    out = list(dec.mid_block.resnets)
    for blk in dec.up_blocks:
        out.extend(list(blk.resnets))
    return out


def oracle_decoder_trace(model, z, dimage):
    """Decoder activations and their gradients for a given upstream gradient d(image)."""
    acts, handles = {}, []

    def keep(name):
        def hook(mod, inp, out):
            out.retain_grad()
            acts[name] = out
        return hook

    dec = model.decoder
    handles.append(dec.conv_in.register_forward_hook(keep("conv_in")))
    for i, r in enumerate(oracle_decoder_resnets(model)):
        handles.append(r.register_forward_hook(keep(f"res{i}_out")))
    for i, blk in enumerate(dec.up_blocks):
        if blk.upsamplers is not None:
            handles.append(blk.upsamplers[0].register_forward_hook(keep(f"up{i}_out")))
    if dec.mid_block.attentions is not None:
        handles.append(dec.mid_block.attentions[0].register_forward_hook(keep("attn_out")))
    with torch.enable_grad():
        zz = z.clone().requires_grad_(True)
        img = model.decode(zz)
        img.backward(dimage)
    for h in handles:
        h.remove()
    grads = {k: v.grad.detach() for k, v in acts.items() if v.grad is not None}
    acts = {k: v.detach() for k, v in acts.items()}
    return acts, grads, img.detach(), zz.grad.detach()


def fetch_decoder_saved(vae, saved, name, index=0):
    from tml_image_editing_defense_b200 import _lib
    off = C.c_size_t()
    dims = (C.c_int * 4)()
    _lib.check(vae._lib.tml_debug_decoder_saved_tensor(vae._h, name.encode(), index, C.byref(off), dims))
    B, H, W, Cc = list(dims)
    n = B * H * W * Cc
    t = saved[off.value: off.value + 2 * n].view(torch.bfloat16).view(B, H, W, Cc)
    return t.float().permute(0, 3, 1, 2).contiguous()
