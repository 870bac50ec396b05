"""Micro-benchmark of ONE 3x3 convolution shape through tml_debug_gemm (CUDA events, after warm-up).

    python tools/gemm_micro.py --B 16 --H 128 --W 128 --Cin 512 --N 512 [--resid] [--bias] [--gn 1|2] [--iters 20]

Environment switches of the kernels (TML_NO_SWAP, TML_NO_SWAP_PAIR, TML_DBG_MMA_ONLY, TML_DBG_NO_EPI ...) are read
once per process, so A/B comparisons run this script once per setting.  Results are NOT checked here (tests do)."""
import argparse
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tml_image_editing_defense_b200 import _lib  # noqa


def main():
    ap = argparse.ArgumentParser()
    for k, v in dict(B=16, H=128, W=128, Cin=512, N=512, iters=20, gn=0).items():
        ap.add_argument(f"--{k}", type=int, default=v)
    ap.add_argument("--k1", action="store_true", help="1x1 convolution instead of 3x3")
    ap.add_argument("--resid", action="store_true")
    ap.add_argument("--bias", action="store_true")
    ap.add_argument("--xf", action="store_true", help="fused input GroupNorm + SiLU (CTA-pair 3x3 kernel): A is the raw input")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    B, H, W, Cin, N = a.B, a.H, a.W, a.Cin, a.N
    A = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    T = 1 if a.k1 else 9
    Wm = (torch.randn(N, T * Cin, device=dev) / (T * Cin) ** 0.5).to(torch.bfloat16)
    D = torch.empty(B, H, W, N, dtype=torch.bfloat16, device=dev)
    d = _lib.TmlGemmDesc()
    d.A = A.data_ptr(); d.A_C = Cin; d.A_W = W; d.A_H = H; d.A_B = B
    d.A_sW = Cin; d.A_sH = W * Cin; d.A_sB = H * W * Cin
    d.stride = 1; d.ntaps = T
    for t in range(T):
        d.dh[t] = (t // 3 - 1) if T == 9 else 0; d.dw[t] = (t % 3 - 1) if T == 9 else 0
    d.OW = W; d.OH = H
    d.Bm = Wm.data_ptr(); d.N = N; d.B_sN = T * Cin; d.B_sBatch = 0; d.alpha = 1.0
    keep = []
    if a.bias:
        keep.append(torch.randn(N, device=dev)); d.bias = keep[-1].data_ptr()
    if a.resid:
        keep.append(torch.randn(B, H, W, N, device=dev).to(torch.bfloat16)); d.resid = keep[-1].data_ptr()
    d.R_sW = N; d.R_sH = W * N; d.R_sB = H * W * N
    d.D = D.data_ptr(); d.D_sW = N; d.D_sH = W * N; d.D_sB = H * W * N; d.D_sN = 1
    if a.xf:
        keep.append(torch.stack([torch.rand(B, Cin, device=dev) + 0.5, torch.randn(B, Cin, device=dev) * 0.5], dim=-1).contiguous())
        d.in_gn_ss = keep[-1].data_ptr()
    if a.gn:
        d.gn_mode = a.gn
        if a.gn == 2:
            keep += [torch.randn(B, H, W, N, device=dev).to(torch.bfloat16), torch.rand(B, N, 2, device=dev) + 0.5,
                     torch.rand(B, 32, 2, device=dev) + 0.5, torch.randn(N, device=dev)]
            d.gn_x, d.gn_ss, d.gn_mr, d.gn_gamma = [t.data_ptr() for t in keep[-4:]]
            d.gn_silu = 1
        n = lib.tml_debug_gn_chunks_per_image(C.byref(d))
        keep.append(torch.empty(B, n, 32, 2, device=dev)); d.gn_partial = keep[-1].data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        rc = lib.tml_debug_gemm(C.byref(d), st)
        if rc:
            print("launch error", lib.tml_last_error().decode()); return
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        lib.tml_debug_gemm(C.byref(d), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    fl = 2.0 * B * H * W * N * T * Cin
    print(f"{a.tag:28s} B={B} {H}x{W} {Cin}->{N} res={int(a.resid)} gn={a.gn}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  out {B * H * W * N * 2 / ms / 1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
