"""Hardware experiment: does a row-shifted (non-1024-byte-aligned) start address work in a K-major
SWIZZLE_128B UMMA shared-memory descriptor, with / without the base_offset field?  (Needed for
reusing one halo tile for the three horizontal taps of a 3x3 convolution.)"""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tml_image_editing_defense_b200 import _lib  # noqa


def run(shift, bo, dev, lib):
    g = torch.Generator().manual_seed(0)
    B, H, W, Cin, N = 2, 1, 64, 128, 64
    A = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16)
    Wm = (torch.randn(N, Cin, generator=g) / Cin ** 0.5).to(torch.bfloat16)
    ref = A.float().view(-1, Cin) @ Wm.float().T
    Ad, Wd = A.to(dev), Wm.to(dev)
    D = torch.full((B, H, W, N), float("nan"), dtype=torch.bfloat16, device=dev)
    d = _lib.TmlGemmDesc()
    d.A = Ad.data_ptr(); d.A_C = Cin; d.A_W = W; d.A_H = H; d.A_B = B
    d.A_sW = Cin; d.A_sH = W * Cin; d.A_sB = H * W * Cin
    d.stride = 1; d.ntaps = 1; d.OW = W; d.OH = H
    d.Bm = Wd.data_ptr(); d.N = N; d.B_sN = Cin; d.alpha = 1.0
    d.D = D.data_ptr(); d.D_sW = N; d.D_sH = W * N; d.D_sB = H * W * N; d.D_sN = 1
    d.dbg_shift = shift; d.dbg_bo = bo
    rc = lib.tml_debug_gemm(C.byref(d), torch.cuda.current_stream().cuda_stream)
    if rc:
        return f"launch error {lib.tml_last_error().decode()}"
    torch.cuda.synchronize()
    out = D.float().cpu().view(-1, N)
    err = float((out - ref).norm() / ref.norm())
    rows_bad = int(((out - ref).abs().max(dim=1).values > 0.05 * ref.abs().max()).sum())
    return f"rel_err={err:.3e} bad_rows={rows_bad}/{out.shape[0]} {'OK' if err < 1e-2 else 'FAIL'}"


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    lib = _lib.load()
    for bo in (0, 1):
        for shift in (0, 1, 2, 3, 4, 7, 8):
            try:
                print(f"shift={shift} base_offset_field={'set' if bo else '0'}: {run(shift, bo, dev, lib)}", flush=True)
            except Exception as e:  # a trap poisons the context: stop
                print(f"shift={shift} bo={bo}: EXCEPTION {e}")
                sys.exit(0)
