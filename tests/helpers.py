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
