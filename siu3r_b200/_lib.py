"""ctypes binding of libsiu3r_b200.so (the C ABI declared in include/siu3r_b200.h).

The product path has NO fallback: if the CUDA library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsiu3r_b200.so")

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_f = C.c_float

# name -> (restype, argtypes); mirrors include/siu3r_b200.h one to one
SIGNATURES = {
    "siu3r_abi_version": (_i, []),
    "siu3r_note_launch": (None, [_i]),
    "siu3r_launch_count": (C.c_longlong, []),
    "siu3r_reset_launch_count": (None, []),
    "siu3r_pdl_enabled": (_i, []),
    "siu3r_set_pdl": (None, [_i]),
    "siu3r_raster_workspace_bytes": (_l, [_i, _i, _i, _l]),
    "siu3r_raster_forward": (_i, [_i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _l, _l,
                                  C.POINTER(C.c_int64), _p, _p, _p, _p, _p, _p]),
    "siu3r_raster_forward_nosync": (_i, [_i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _l, _l, _p, _p]),
    "siu3r_raster_set_regsort": (None, [_i]),
    "siu3r_raster_set_culling": (None, [_i]),
    "siu3r_raster_set_binning": (None, [_i]),
    "siu3r_raster_features_forward_nosync": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _f, _f, _f, _f, _f, _f, _p, _p, _p, _p, _l, _l, _p, _p]),
    "siu3r_raster_features_forward": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _l, _l, C.POINTER(C.c_int64), _p]),
    "siu3r_gemm_tc": (_i, [_i, _i, _i, _p, _p, _l, _p, _p, _l, _p, _l, _p, _p, _l, _i, _f, _i, _p]),
    "siu3r_conv2d_tc": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _l, _p, _p, _l, _i, _i, _p]),
    "siu3r_gemm_simt": (_i, [_i, _i, _i, _p, _l, _p, _l, _p, _l, _p, _p, _l, _i, _f, _p]),
    "siu3r_gemm_skinny": (_i, [_i, _i, _i, _p, _l, _p, _l, _p, _l, _p, _p, _l, _i, _f, _p]),
    "siu3r_split_tf32": (_i, [_p, _p, _p, _l, _p]),
    "siu3r_gemm_debug_set": (None, [_p]),
    "siu3r_gemm_force": (None, [_i]),
    "siu3r_gemm_plan": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "siu3r_conv_rows_up2x_tc": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _l, _i, _p]),
    "siu3r_ply_record_words": (_i, [_i, _i, _i, _i]),
    "siu3r_labels_from_qc_logits": (_i, [_p, _i, _i, _i, _i, _i, _l, _l, _l, _l, _l, _f, _p, _p, _i, _p, _p, _p, _p, _p]),
    "siu3r_resize_lanczos_u8": (_i, [_p, _i, _i, _l, _p, _p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "siu3r_ply_pack": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _l, _i, _i, _i, _p, _p]),
    "siu3r_gemm_tc_group2": (_i, [_p, _i, _i, _p, _l, _p, _l, _p, _l, _p, _p, _l, _i, _f, _p, _p, _i, _p, _p, _l, _i, _p]),
    "siu3r_gemm_tc_rope_vt": (_i, [_i, _i, _i, _p, _l, _p, _l, _p, _l, _p, _i, _p, _p, _i, _p, _l, _i, _p]),
    "siu3r_gemm_tc_rope": (_i, [_i, _i, _i, _p, _p, _l, _p, _p, _l, _p, _l, _p, _i, _i, _p, _p, _i, _p]),
    "siu3r_rope2d_table": (_i, [_p, _i, _i, _f, _f, _p]),
    "siu3r_rope2d": (_i, [_p, _p, _i, _i, _i, _i, _l, _l, _f, _f, _i, _l, _i, _p]),
    "siu3r_transpose_v": (_i, [_p, _l, _l, _i, _i, _i, _p, _l, _p]),
    "siu3r_flash_attn_tc": (_i, [_p, _l, _l, _i, _i, _p, _l, _l, _i, _i, _p, _l, _l, _p, _l, _l, _i, _i, _i, _i, _f, _i, _p]),
    "siu3r_layernorm_group2": (_i, [_p, _p, _l, _p, _p, _p, _p, _p, _p, _l, _i, _i, _i, _f, _i, _p]),
    "siu3r_layernorm": (_i, [_p, _l, _p, _p, _p, _l, _i, _i, _f, _p, _l, _i, _p]),
    "siu3r_flash_attn_d64": (_i, [_p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _l, _l, _i, _i, _i, _i, _f, _i, _i, _p]),
    "siu3r_attn_small_d32_ws_bytes": (_l, [_i, _i, _i, _i]),
    "siu3r_attn_small_d32": (_i, [_p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _l, _l, _p, _i, _i, _i, _i, _f, _i, _p, _l, _p]),
    "siu3r_msdeform_attn": (_i, [_p, _l, _i, _p, _l, _p, C.POINTER(C.c_int), _i, _i, _i, _i, _i, _i, _p, _l, _i, _p]),
    "siu3r_eltwise": (_i, [_i, _p, _p, _p, _l, _p]),
    "siu3r_scale": (_i, [_p, _f, _p, _l, _p]),
    "siu3r_rows_affine": (_i, [_p, _l, _p, _p, _p, _l, _p, _l, _l, _i, _i, _p]),
    "siu3r_resize_bilinear_nhwc": (_i, [_p, _i, _i, _i, _i, _l, _p, _i, _i, _l, _i, _i, _p]),
    "siu3r_pixel_shuffle_nhwc": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "siu3r_im2col_nhwc": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _l, _i, _p]),
    "siu3r_nchw_to_nhwc": (_i, [_p, _p, _i, _i, _i, _l, _p]),
    "siu3r_nhwc_to_nchw": (_i, [_p, _l, _p, _i, _i, _i, _p]),
    "siu3r_maxpool3x3s2_nhwc": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "siu3r_dwconv3x3_nhwc": (_i, [_p, _l, _l, _i, _i, _i, _i, _p, _p, _p, _l, _l, _i, _p]),
    "siu3r_groupnorm_nhwc": (_i, [_p, _i, _i, _i, _i, _p, _p, _f, _i, _p, _p, _p]),
    "siu3r_depth_exp": (_i, [_p, _l, _p, _l, _p]),
    "siu3r_gaussian_adapter": (_i, [_p, _l, _p, _p, _p, _p, _p, _p]),
    "siu3r_attn_mask_from_logits": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "siu3r_resize_select": (_i, [_p, _i, _i, _i, _i, _p, _i, _p, _i, _i, _p]),
    "siu3r_argmax_area": (_i, [_p, _l, _i, _p, _f, _p, _p, _p, _p]),
    "siu3r_label_lut": (_i, [_p, _l, _p, _p, _p, _p, _p, _p]),
    "siu3r_qc_logits": (_i, [_p, _l, _i, _p, _i, _p, _i, _p, _p]),
    "siu3r_render_record_pack": (_i, [_p, _p, _p, _p, _l, _p, _p]),
    "siu3r_render_record_unpack": (_i, [_p, _l, _p, _p, _p, _p, _p]),
    # h3 mode (fp32-grade results on the fp16 tensor-core path)
    "siu3r_gemm_h3_force": (None, [_i]),
    "siu3r_gemm_h3_set_mhalf": (None, [_i]),
    "siu3r_gemm_h3_set_remainder_tiles": (None, [_i]),
    "siu3r_gemm_h3_cluster_cap": (None, [_i]),
    "siu3r_gemm_h3_order": (None, [_i]),
    "siu3r_gemm_h3_debug": (None, [_i]),
    "siu3r_gemm_h3_debug_ts": (None, [_p]),
    "siu3r_gemm_h3_plan": (_i, [_i, _i, _i, _i, _p, _p, _p]),
    "siu3r_split_h3": (_i, [_p, _l, _l, _i, _p, _l, _l, _i, _p]),
    "siu3r_merge_h3": (_i, [_p, _l, _l, _l, _i, _p, _l, _p]),
    "siu3r_gemm_h3": (_i, [_i, _p, _i, _i, _p, _l, _l, _p, _l, _l, _p, _l, _p, _l, _l, _p, _p, _l, _i, _f, _p, _p, _i, _p, _p, _l, _l, _i, _i, _p]),
    "siu3r_gemm_h3_ln": (_i, [_i, _p, _i, _i, _p, _l, _l, _p, _l, _l, _p, _l, _p, _l, _l, _p, _p, _l, _i, _f, _p, _p, _i, _p, _p, _l, _l, _i, _i, _p, _p,
                              _f, _p, _p]),
    "siu3r_conv2d_h3": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _l, _l, _p, _l, _l, _p, _l, _p, _l, _l, _p, _p, _l, _i, _p]),
    "siu3r_flash_h3_debug_swap": (None, [_i]),
    "siu3r_flash_attn_h3": (_i, [_p, _l, _l, _l, _i, _i, _p, _l, _l, _l, _i, _i, _p, _l, _l, _l, _i, _l, _p, _l, _l, _p, _l, _l, _l, _i, _i, _i, _i,
                                 _f, _p]),
    "siu3r_transpose_v_h3": (_i, [_p, _l, _l, _i, _i, _i, _p, _l, _l, _p]),
    "siu3r_layernorm_h3": (_i, [_p, _p, _l, _p, _p, _p, _p, _p, _p, _l, _p, _p, _l, _l, _i, _i, _i, _f, _p]),
    "siu3r_eltwise_h3": (_i, [_i, _p, _p, _p, _l, _l, _p]),
    "siu3r_resize_bilinear_nhwc_h3": (_i, [_p, _i, _i, _i, _i, _l, _p, _i, _i, _l, _l, _i, _p]),
    "siu3r_im2col_nhwc_h3": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _l, _l, _p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; fails loudly (no CPU / library fallback exists for the product path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"siu3r_b200: CUDA extension not built ({LIB_PATH} missing). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python siu3r_b200/build.py`. There is no fallback path."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


ERRORS = {-1: "invalid argument", -2: "capacity / workspace too small", -3: "CUDA error", -4: "unsupported configuration"}


def check(code: int, what: str):
    if code != 0:
        raise RuntimeError(f"siu3r_b200.{what} failed: {ERRORS.get(code, code)} (see stderr)")
