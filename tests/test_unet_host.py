"""CPU tests of the native UNet's host side (no GPU): weight packing of the pad-1 stride-2 convolution and its parity
dgrads, and a dry run of both walks (layout, scratch size, planning of every GEMM) through the C ABI."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from tests.helpers import emulate_gemm, pack_conv3x3


def test_downsample_pad1_packing_matches_conv2d():
    """mode 7 = Conv2d(stride 2, padding 1) forward; modes 8..11 = its input gradient per output parity class."""
    g = torch.Generator().manual_seed(0)
    Ci, Co, H, W = 64, 64, 8, 16
    w = torch.randn((Co, Ci, 3, 3), generator=g).bfloat16().float()
    x = torch.randn((2, Ci, H, W), generator=g).bfloat16().float()
    mat, dh, dw = pack_conv3x3(w, 7)
    y = emulate_gemm(x.permute(0, 2, 3, 1), mat, dh, dw, 2, H // 2, W // 2).permute(0, 3, 1, 2)
    ref = F.conv2d(x.double(), w.double(), stride=2, padding=1)
    assert torch.allclose(y, ref, atol=1e-9)
    dy = torch.randn((2, Co, H // 2, W // 2), generator=g).bfloat16().float()
    xr = x.double().requires_grad_(True)
    (dx_ref,) = torch.autograd.grad((F.conv2d(xr, w.double(), stride=2, padding=1) * dy.double()).sum(), [xr])
    dx = torch.zeros_like(dx_ref)
    for q in range(4):
        ph, pw = q >> 1, q & 1
        mat, dh, dw = pack_conv3x3(w, 8 + q)
        part = emulate_gemm(dy.permute(0, 2, 3, 1), mat, dh, dw, 1, H // 2, W // 2).permute(0, 3, 1, 2)
        dx[:, :, ph::2, pw::2] = part
    assert torch.allclose(dx, dx_ref, atol=1e-9)


def _dry_run(cfgp, B, h, w, T):
    from tml_image_editing_defense_b200 import _lib
    from tml_image_editing_defense_b200.unet_torch import UNet2DConditionModel
    lib = _lib.load()
    lib.tml_debug_set_host_only(1)
    try:
        torch.manual_seed(0)
        m = UNet2DConditionModel(cfgp)
        cfg = _lib.TmlUnetCfg()
        cfg.in_channels, cfg.out_channels = cfgp.in_channels, cfgp.out_channels
        cfg.num_blocks = len(cfgp.block_out_channels)
        for i, ch in enumerate(cfgp.block_out_channels):
            cfg.block_out_channels[i] = ch
            cfg.down_has_attn[i] = int(cfgp.down_has_attn[i])
            cfg.up_has_attn[i] = int(cfgp.up_has_attn[i])
        cfg.layers_per_block = cfgp.layers_per_block
        cfg.cross_attention_dim = cfgp.cross_attention_dim
        cfg.num_heads = cfgp.attention_head_dim
        cfg.norm_num_groups = cfgp.norm_num_groups
        hnd = C.c_void_p()
        _lib.check(lib.tml_unet_create(C.byref(cfg), 0, C.byref(hnd)))
        try:
            for k, v in m.state_dict().items():
                t = v.detach().float().contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(lib.tml_unet_set_weight(hnd, k.encode(), t.data_ptr(), 0, shape, t.dim()))
            _lib.check(lib.tml_unet_finalize(hnd, None))
            ws, sv = C.c_size_t(), C.c_size_t()
            rc = lib.tml_unet_query(hnd, B, h, w, T, C.byref(ws), C.byref(sv))
            msg = lib.tml_last_error().decode()
            n_res, n_tf = C.c_size_t(), C.c_size_t()
            dims = (C.c_int * 4)()
            if rc == 0:
                lib.tml_debug_unet_saved_tensor(hnd, b"count_resnets", 0, C.byref(n_res), dims)
                lib.tml_debug_unet_saved_tensor(hnd, b"count_tf", 0, C.byref(n_tf), dims)
            return rc, msg, ws.value, sv.value, n_res.value, n_tf.value
        finally:
            lib.tml_unet_destroy(hnd)
    finally:
        lib.tml_debug_set_host_only(0)


def test_unet_dry_run_plans_every_gemm_of_a_three_level_net():
    from tests.gpu_check_unet import tiny_native_config
    rc, msg, ws, sv, n_res, n_tf = _dry_run(tiny_native_config(), 2, 32, 32, 5)
    assert rc == 0, msg
    assert n_res == 2 * 3 + 2 + 3 * 3 and n_tf == 4 + 1 + 6   # down 2x3, mid 2, up 3x3; attn: down (2+2), mid, up (3+3)
    assert ws > 0 and sv > 0
    rc2, _, ws2, sv2, _, _ = _dry_run(tiny_native_config(), 4, 32, 32, 5)
    assert rc2 == 0 and sv2 > 1.9 * sv      # saved state scales with the batch


def test_unet_dry_run_rejects_shapes_the_kernels_cannot_tile():
    from tests.gpu_check_unet import tiny_native_config
    rc, msg, *_ = _dry_run(tiny_native_config(), 2, 16, 16, 5)    # 16 -> 8 -> 4: rows of 4 pixels have no tile
    assert rc != 0 and "tile" in msg
