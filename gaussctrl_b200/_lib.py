"""ctypes binding of libgaussctrl_b200.so (the C ABI declared in include/gaussctrl_b200.h).

The product path has no fallback: if the shared library cannot be loaded (and cannot be built with nvcc),
importing this module raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_longlong, c_size_t, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCB_LIB_PATH") or os.path.join(HERE, "libgaussctrl_b200.so")  # override: A/B builds

GCB_ACT_NONE, GCB_ACT_SILU, GCB_ACT_GEGLU = 0, 1, 2
GCB_GEMM_TCGEN05, GCB_GEMM_MMA_SYNC, GCB_GEMM_TCGEN05_DIRECT = 0, 1, 2
GCB_GEMM_TCGEN05_PERSISTENT, GCB_GEMM_TCGEN05_ONE_TILE = 3, 4
GCB_ATTN_AUTO, GCB_ATTN_TCGEN05, GCB_ATTN_MMA_SYNC = 0, 1, 2

_P = c_void_p
_FP = POINTER(c_float)
_IP = POINTER(c_int32)

# name -> (restype, argtypes); mirrors include/gaussctrl_b200.h one to one (tests check the symbol list)
SIGNATURES = {
    "gcb_version": (c_int, []),
    "gcb_last_error": (c_char_p, []),
    "gcb_device_info": (c_int, [POINTER(c_int)] * 3),
    "gcb_conv2d_nhwc_fwd": (c_int, [_P, _P, _P, _P, c_int, _P, _P] + [c_int] * 8 + [_P]),
    "gcb_conv2d_direct_nhwc_fwd": (c_int, [_P, _P, _P, _P, _P] + [c_int] * 10 + [_P]),
    "gcb_im2col3x3_s2_nhwc": (c_int, [_P, _P] + [c_int] * 6 + [_P]),
    "gcb_im2col3x3_c4_nhwc": (c_int, [_P, _P] + [c_int] * 3 + [_P]),
    "gcb_geglu_tile_n": (c_int, [c_int]),
    "gcb_geglu_pack_rows": (c_int, [c_int, _IP]),
    "gcb_groupnorm_workspace_bytes": (c_size_t, [c_int, c_int]),
    "gcb_groupnorm_nhwc_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_int, _P,
                                       c_size_t, _P]),
    "gcb_layernorm_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_float, _P]),
    "gcb_attn_multi_fwd": (c_int, [_P, c_int, _P, _P, c_int, _P, _P, c_int, _P, c_int] + [c_int] * 7 +
                           [_P, _FP, c_float, c_int, _P]),
    "gcb_softmax_rows_fwd": (c_int, [_P, _P, c_int, c_int, c_float, _P]),
    "gcb_silu_fwd": (c_int, [_P, _P, c_longlong, _P]),
    "gcb_add_fwd": (c_int, [_P, _P, _P, c_longlong, c_float, c_float, _P]),
    "gcb_geglu_fwd": (c_int, [_P, _P, c_int, c_int, _P]),
    "gcb_upsample_nearest2x_nhwc": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "gcb_timestep_embedding": (c_int, [_P, c_int, c_int, _P, _P]),
    "gcb_nchw_to_nhwc_f16": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "gcb_nhwc_to_nchw_f16": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "gcb_transpose_f16": (c_int, [_P, _P, c_int, c_int, c_int, _P]),
    "gcb_cfg_ddim_step": (c_int, [_P, _P, _P, _P, c_longlong, c_float, _P, _P]),
    "gcb_postprocess_composite": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "gcb_depth_to_disparity": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P]),
    "gcb_project_gaussians_fwd": (c_int, [_P, _P, c_float, _P, _FP, _FP, c_float, c_float, c_float, c_float, c_int,
                                          c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "gcb_project_sh_fused_fwd": (c_int, [_P] * 6 + [_FP, _FP, _FP, c_float, c_float, c_float, c_float] + [c_int] * 6 +
                                 [_P] * 7 + [_P]),
    "gcb_sh_fwd": (c_int, [c_int, c_int, _P, _P, _P, c_int, _P]),
    "gcb_scan_workspace_bytes": (c_size_t, [c_int]),
    "gcb_cumsum_i32": (c_int, [_P, _P, c_int, _P, c_size_t, _P]),
    "gcb_bin_gaussians_workspace_bytes": (c_size_t, [c_int, c_longlong, c_int, c_int]),
    "gcb_bin_gaussians": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_longlong, _P, _P, _P, _P, _P, c_size_t, _P]),
    "gcb_rasterize_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _FP, _P, _P, _P, _P]),
    "gcb_rasterize_rgbd_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P]),
    "gcb_render_eval_batch_workspace_bytes": (c_size_t, [c_int, c_longlong, c_int, c_int]),
    "gcb_render_eval_batch": (c_int, [_P] * 6 + [c_int, c_int, c_int, _FP, _FP, _FP, _FP, c_int, c_int, _P, c_longlong,
                                      _P, _P, _P, _P, _P, c_size_t, _P]),
    "gcb_handle_control_bytes": (c_size_t, []),
    "gcb_handle_create": (c_int, [c_int, c_int, c_size_t, POINTER(_P)]),
    "gcb_handle_destroy": (c_int, [_P]),
    "gcb_handle_arena": (_P, [_P]),
    "gcb_handle_ipc_export": (c_int, [_P, ctypes.c_char_p]),
    "gcb_handle_ipc_open": (c_int, [_P, c_int, ctypes.c_char_p]),
    "gcb_allgather_ref_kv": (c_int, [_P, c_size_t, _P, c_size_t, c_int, _P]),
    "gcb_linear_allgather_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_size_t, c_int, _P]),
    "gcb_peer_barrier": (c_int, [_P, c_int, _P]),
    "gcb_handle_error": (c_int, [_P, POINTER(c_int)]),
    "gcb_rasterize_bwd": (c_int, [_P] * 6 + [c_int, c_int, c_int, _FP] + [_P] * 8 + [_P]),
    "gcb_project_gaussians_bwd": (c_int, [_P, _P, c_float, _P, _FP, _FP, c_float, c_float, c_float, c_float, c_int, c_int,
                                          _P, _P, _P, _P, c_int, _P, _P, _P, _P]),
    "gcb_sh_bwd": (c_int, [c_int, c_int, _P, _P, _P, c_int, _P]),
    "gcb_raster_finalize": (c_int, [_P, _P, _P, _P, _P, c_int, _P]),
    "gcb_l1_ssim_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gcb_l1_ssim_loss_fwd_bwd": (c_int, [_P, _P, c_int, c_int, c_int, c_float, _P, _P, _P, c_size_t, _P]),
    "gcb_embed_tokens_f16": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "gcb_quick_gelu_fwd": (c_int, [_P, _P, c_longlong, _P]),
    "gcb_attn_causal_fwd": (c_int, [_P, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_float, _P]),
    "gcb_adam_step": (c_int, [c_int, POINTER(_P), POINTER(_P), POINTER(_P), POINTER(_P), POINTER(c_longlong),
                              POINTER(c_double), c_double, c_double, c_double, c_int, _P]),
}


class GcbError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        # the library is built in-tree by __graft_entry__.build(); try once here so a fresh checkout works
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise GcbError(f"gaussctrl_b200 error {rc}: {lib.gcb_last_error().decode()}")
