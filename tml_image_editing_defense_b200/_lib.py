"""ctypes binding of the C ABI in ``include/tml_b200.h`` (the library is ``csrc/libtml_b200.so``).

There is deliberately no fallback: if the shared library is missing or cannot be loaded the
import of any compute path raises, and on a machine without an sm_100a GPU
``tml_encoder_create`` returns an error that is raised as ``TmlError``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# TML_LIB_PATH selects another build of the same sources (A/B timing of compile-time tuning constants); the default
# is the in-tree library every test and benchmark uses.
_LIB_PATH = Path(os.environ["TML_LIB_PATH"]).resolve() if os.environ.get("TML_LIB_PATH") else \
    Path(__file__).resolve().parent / "csrc" / "libtml_b200.so"


class TmlError(RuntimeError):
    pass


class TmlEncoderCfg(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int), ("latent_channels", C.c_int), ("num_blocks", C.c_int),
        ("block_out_channels", C.c_int * 8), ("layers_per_block", C.c_int), ("norm_num_groups", C.c_int),
        ("norm_eps", C.c_float), ("mid_block_add_attention", C.c_int),
    ]


class TmlUnetCfg(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int), ("out_channels", C.c_int), ("num_blocks", C.c_int),
        ("block_out_channels", C.c_int * 8), ("layers_per_block", C.c_int), ("cross_attention_dim", C.c_int),
        ("num_heads", C.c_int), ("norm_num_groups", C.c_int), ("down_has_attn", C.c_int * 8),
        ("up_has_attn", C.c_int * 8),
    ]


class TmlGemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("A_C", C.c_int), ("A_W", C.c_int), ("A_H", C.c_int), ("A_B", C.c_int),
        ("A_sW", C.c_int64), ("A_sH", C.c_int64), ("A_sB", C.c_int64),
        ("stride", C.c_int), ("ntaps", C.c_int), ("dh", C.c_int * 9), ("dw", C.c_int * 9),
        ("OW", C.c_int), ("OH", C.c_int),
        ("Bm", C.c_void_p), ("N", C.c_int), ("B_sN", C.c_int64), ("B_sBatch", C.c_int64),
        ("alpha", C.c_float), ("bias", C.c_void_p), ("resid", C.c_void_p),
        ("R_sB", C.c_int64), ("R_sH", C.c_int64), ("R_sW", C.c_int64),
        ("D", C.c_void_p), ("out_fp32", C.c_int),
        ("D_sB", C.c_int64), ("D_sH", C.c_int64), ("D_sW", C.c_int64), ("D_sN", C.c_int64), ("n_store", C.c_int), ("beta", C.c_float),
        ("gn_mode", C.c_int), ("gn_partial", C.c_void_p), ("gn_x", C.c_void_p), ("gn_ss", C.c_void_p),
        ("gn_mr", C.c_void_p), ("gn_gamma", C.c_void_p), ("gn_silu", C.c_int),
        ("dbg_shift", C.c_int), ("dbg_bo", C.c_int),
        ("in_gn_ss", C.c_void_p),
        ("a_trans", C.c_int), ("A_sK", C.c_int64),
    ]


# name -> (restype, argtypes); every symbol include/tml_b200.h declares
SIGNATURES = {
    "tml_last_error": (C.c_char_p, []),
    "tml_version": (C.c_int, []),
    "tml_encoder_create": (C.c_int, [C.POINTER(TmlEncoderCfg), C.c_int, C.POINTER(C.c_void_p)]),
    "tml_encoder_destroy": (None, [C.c_void_p]),
    "tml_encoder_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_int]),
    "tml_encoder_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tml_encoder_query": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "tml_encoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "tml_encoder_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_float, C.c_void_p, C.c_void_p]),
    "tml_decoder_query": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "tml_decoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "tml_decoder_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "tml_image_loss_workspace": (C.c_size_t, [C.c_int]),
    "tml_image_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_float,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tml_posterior_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tml_posterior_sample_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                C.c_int, C.c_void_p]),
    "tml_debug_decoder_saved_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_size_t),
                                                 C.POINTER(C.c_int)]),
    "tml_latent_loss": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tml_pgd_step_linf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.c_int64, C.c_void_p]),
    "tml_pgd_l2_workspace": (C.c_size_t, [C.c_int]),
    "tml_pgd_step_l2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "tml_add_delta": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "tml_batch_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_void_p]),
    "tml_universal_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_int64, C.c_void_p, C.c_void_p]),
    "tml_universal_project": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int64, C.c_void_p]),
    "tml_unet_create": (C.c_int, [C.POINTER(TmlUnetCfg), C.c_int, C.POINTER(C.c_void_p)]),
    "tml_unet_destroy": (None, [C.c_void_p]),
    "tml_unet_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_int]),
    "tml_unet_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tml_unet_query": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t),
                                 C.POINTER(C.c_size_t)]),
    "tml_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tml_unet_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "tml_debug_unet_saved_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_size_t),
                                              C.POINTER(C.c_int)]),
    "tml_debug_set_host_only": (None, [C.c_int]),
    "tml_debug_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tml_launch_counts": (None, [C.POINTER(C.c_int64)]),
    "tml_gemm_timing_enable": (None, [C.c_int]),
    "tml_gemm_timing_collect": (None, [C.POINTER(C.c_double)]),
    "tml_gemm_timing_report": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "tml_debug_last_hang": (C.c_int, []),
    "tml_debug_set_gemm_impl": (None, [C.c_int]),
    "tml_debug_gemm": (C.c_int, [C.POINTER(TmlGemmDesc), C.c_void_p]),
    "tml_debug_gn_tiles_per_image": (C.c_int, [C.c_int, C.c_int]),
    "tml_debug_gn_chunks_per_image": (C.c_int, [C.c_void_p]),
    "tml_debug_unet_gn_chunks_per_image": (C.c_int, [C.c_int, C.c_int]),
    "tml_debug_saved_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "tml_debug_set_grad_dump": (None, [C.c_void_p, C.c_size_t, C.c_int]),
    "tml_debug_pack_conv3x3": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int),
                                         C.POINTER(C.c_int), C.POINTER(C.c_int)]),
}

_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        # not built yet (fresh checkout): compile the CUDA sources in tree; still no non-CUDA fallback
        try:
            from .build import build_extension
            build_extension()
        except Exception as e:  # noqa: BLE001
            raise TmlError(f"{_LIB_PATH} is missing and building it failed: {e}") from e
    if not _LIB_PATH.exists():
        raise TmlError(
            f"{_LIB_PATH} not found: build it with `python -m tml_image_editing_defense_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().tml_last_error()
        raise TmlError(f"tml_b200 error {rc}: {msg.decode() if msg else '?'}")


def launch_counts():
    out = (C.c_int64 * 2)()
    load().tml_launch_counts(out)
    return int(out[0]), int(out[1])
