"""ctypes binding of ``libgandanet_sm100.so`` (C ABI declared in ``include/gandanet.h``).

There is no CPU fallback: if the library is missing or the device is not sm_100 every product call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgandanet_sm100.so")

ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
PREC_FP32, PREC_BF16, PREC_FP16, PREC_BF16X3, PREC_FP16X3 = 0, 1, 2, 3, 4

_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t


class ConvArgs(C.Structure):
    _fields_ = [("x", _vp), ("x_pitch", _i), ("x_c0", _i),
                ("w", _vp), ("w_k_pitch", _i), ("w_c0", _i), ("w_group_stride", _ll), ("groups", _i),
                ("y", _vp), ("y_pitch", _i), ("y_c0", _i),
                ("bias", _vp), ("alpha_ptr", _vp),
                ("res", _vp), ("res_pitch", _i), ("res_c0", _i),
                ("B", _i), ("Hi", _i), ("Wi", _i), ("Cin", _i), ("Ho", _i), ("Wo", _i), ("Cout", _i),
                ("kh", _i), ("kw", _i), ("stride", _i), ("pad", _i), ("transposed", _i),
                ("act", _i), ("slope", _f),
                ("splits", _i), ("ws", _vp), ("ws_bytes", _sz)]


class WgradArgs(C.Structure):
    _fields_ = [("dy", _vp), ("dy_pitch", _i), ("dy_c0", _i),
                ("x", _vp), ("x_pitch", _i), ("x_c0", _i),
                ("out", _vp), ("layout", _i), ("out_cin_total", _i), ("out_c0", _i), ("accumulate", _i),
                ("scale_ptr", _vp), ("scale", _f),
                ("B", _i), ("Hi", _i), ("Wi", _i), ("Cin", _i), ("Ho", _i), ("Wo", _i), ("Cout", _i),
                ("kh", _i), ("kw", _i), ("stride", _i), ("pad", _i), ("groups", _i),
                ("splits", _i), ("ws", _vp), ("ws_bytes", _sz)]


class ConvTcArgs(C.Structure):
    _fields_ = [("x_hi", _vp), ("x_lo", _vp), ("w_hi", _vp), ("w_lo", _vp),
                ("y", _vp), ("y_pitch", _i), ("y_c0", _i),
                ("bias", _vp),
                ("res", _vp), ("res_pitch", _i), ("res_c0", _i),
                ("B", _i), ("Hi", _i), ("Wi", _i), ("Cin", _i), ("Ho", _i), ("Wo", _i), ("Cout", _i),
                ("kh", _i), ("kw", _i), ("stride", _i), ("pad", _i), ("transposed", _i),
                ("act", _i), ("slope", _f), ("precision", _i), ("alpha_ptr", _vp), ("groups", _i), ("y16", _vp), ("y16_pitch", _i)]


class WgradTcArgs(C.Structure):
    _fields_ = [("dy_hi", _vp), ("dy_lo", _vp), ("x_hi", _vp), ("x_lo", _vp),
                ("out", _vp), ("out_cin_total", _i), ("out_c0", _i), ("accumulate", _i), ("scale", _f),
                ("B", _i), ("Hi", _i), ("Wi", _i), ("Cin", _i), ("Ho", _i), ("Wo", _i), ("Cout", _i),
                ("kh", _i), ("kw", _i), ("stride", _i), ("pad", _i),
                ("precision", _i), ("ws", _vp), ("ws_bytes", _sz), ("groups", _i), ("scale_ptr", _vp)]


class PamFwdArgs(C.Structure):
    _fields_ = [("q", _vp), ("k", _vp), ("qk_pitch", _i), ("d", _i),
                ("v", _vp), ("v_pitch", _i),
                ("x", _vp), ("x_pitch", _i), ("gamma", _vp),
                ("o", _vp), ("y", _vp), ("y_pitch", _i), ("lse", _vp),
                ("B", _i), ("N", _i), ("C", _i), ("precision", _i), ("chunk", _i),
                ("ws", _vp), ("ws_bytes", _sz), ("v16", _vp), ("y16", _vp), ("y16_pitch", _i)]


class PamBwdArgs(C.Structure):
    _fields_ = [("q", _vp), ("k", _vp), ("qk_pitch", _i), ("d", _i),
                ("v", _vp), ("v_pitch", _i),
                ("o", _vp), ("lse", _vp), ("gamma", _vp),
                ("dy", _vp), ("dy_pitch", _i),
                ("dq", _vp), ("dk", _vp), ("dv", _vp), ("rowdot", _vp),
                ("B", _i), ("N", _i), ("C", _i), ("precision", _i), ("chunk", _i),
                ("ws", _vp), ("ws_bytes", _sz)]


# name -> (restype, argtypes).  Every symbol include/gandanet.h declares is listed here (tests check the header against it).
SIGNATURES = {
    "gdn_version": (_i, []),
    "gdn_last_error": (C.c_char_p, []),
    "gdn_init": (_i, [_i]),
    "gdn_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "gdn_nhwc_to_nchw": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _vp]),
    "gdn_weight_oihw_to_ohwi": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "gdn_weight_oihw_to_ihwo": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "gdn_conv2d": (_i, [C.POINTER(ConvArgs), _vp]),
    "gdn_conv2d_wgrad": (_i, [C.POINTER(WgradArgs), _vp]),
    "gdn_conv2d_suggest_splits": (_i, [C.POINTER(ConvArgs)]),
    "gdn_wgrad_suggest_splits": (_i, [C.POINTER(WgradArgs)]),
    "gdn_pack_act_bf16": (_i, [_vp, _i, _i, _ll, _i, _vp, _vp, _vp, _vp, _i, _f, _vp]),
    "gdn_pack_actgrad_bf16": (_i, [_vp, _i, _vp, _i, _ll, _i, _vp, _vp, _i, _f, _vp]),
    "gdn_pack_actgrad_bf16g": (_i, [_vp, _i, _vp, _i, _ll, _i, _vp, _vp, _i, _f, _vp]),
    "gdn_pack_weight_bf16_elems": (_sz, [_i, _i, _i, _i, _i]),
    "gdn_pack_weight_bf16": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "gdn_conv2d_tc": (_i, [C.POINTER(ConvTcArgs), _vp]),
    "gdn_conv_tc_set_halo": (_i, [_i]),
    "gdn_conv_tc_set_wgrad_swap": (_i, [_i]),
    "gdn_conv_tc_set_wgrad_col": (_i, [_i]),
    "gdn_conv_tc_set_col": (_i, [_i]),
    "gdn_conv_tc_set_wgrad_col_row": (_i, [_i]),
    "gdn_conv2d_wgrad_tc_ws_bytes": (_sz, [C.POINTER(WgradTcArgs)]),
    "gdn_conv2d_wgrad_tc": (_i, [C.POINTER(WgradTcArgs), _vp]),
    "gdn_linear_tc_supported": (_i, [_i, _i, _i]),
    "gdn_linear_tc_fwd_ws_bytes": (_sz, [_i, _i, _i]),
    "gdn_linear_tc_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _sz, _vp]),
    "gdn_linear_tc_dgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "gdn_linear_tc_wgrad_ws_bytes": (_sz, [_i, _i, _i]),
    "gdn_linear_tc_wgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "gdn_thin_conv_supported": (_i, [_i, _i, _i]),
    "gdn_thin_conv_expand": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "gdn_thin_conv_expand_p": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp]),
    "gdn_thin_conv_reduce": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "gdn_thin_conv_reduce_gated": (_i, [_vp, _i, _vp, _i, _f, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "gdn_thin_conv_tap_supported": (_i, [_i, _i, _i]),
    "gdn_thin_conv_tap_mask_bytes": (_sz, [_i, _i, _i, _i]),
    "gdn_thin_conv_tap_pair": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gdn_thin_conv_tap_dgrad": (_i, [_vp, _i, _vp, _vp, _f, _vp, _vp, _i, _i, _i, _i, _vp]),
    "gdn_thin_conv_wgrad_ws_bytes": (_sz, [_i, _i, _i, _i]),
    "gdn_thin_conv_wgrad": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "gdn_colstats_ws_bytes": (_sz, [_ll, _i]),
    "gdn_colstats": (_i, [_vp, _i, _i, _ll, _i, _vp, _vp, _vp]),
    "gdn_bn_finalize": (_i, [_vp, _ll, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gdn_bn_eval_coeffs": (_i, [_vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp]),
    "gdn_affine_act": (_i, [_vp, _i, _i, _vp, _i, _i, _ll, _i, _vp, _vp, _i, _f, _vp]),
    "gdn_stat_fused_ws_bytes": (_sz, [_ll, _i]),
    "gdn_stat_fused_counters": (_i, []),
    "gdn_bn_stats": (_i, [_vp, _i, _i, _ll, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gdn_colsums_f": (_i, [_vp, _i, _i, _ll, _i, _vp, _vp, _vp, _vp, _vp]),
    "gdn_bn_bwd_reduce_f": (_i, [_vp, _i, _i, _vp, _i, _i, _ll, _i, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "gdn_bn_bwd_reduce": (_i, [_vp, _i, _i, _vp, _i, _i, _ll, _i, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp]),
    "gdn_bn_bwd_apply": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _i, _ll, _i, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp]),
    "gdn_bn_bwd_apply16": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i, _ll, _i, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp]),
    "gdn_act_bwd": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _ll, _i, _i, _f, _vp]),
    "gdn_axpy": (_i, [_vp, _i, _i, _vp, _i, _i, _ll, _i, _f, _i, _vp]),
    "gdn_scale_dev": (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    "gdn_sums_to_float": (_i, [_vp, _vp, _i, _f, _vp]),
    "gdn_bicubic_up2_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "gdn_bicubic_up2_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "gdn_bicubic_up2_bilinear_add_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "gdn_tap_shift_sum": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "gdn_tap_shift_expand": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "gdn_narrow_conv1x1_fwd": (_i, [_vp, _i, _vp, _i, _vp, _ll, _vp]),
    "gdn_narrow_conv1x1_dgrad": (_i, [_vp, _i, _vp, _i, _vp, _ll, _i, _vp]),
    "gdn_narrow_conv1x1_wgrad_ws_bytes": (_sz, [_ll, _i, _i]),
    "gdn_narrow_conv1x1_wgrad": (_i, [_vp, _i, _vp, _i, _ll, _vp, _i, _vp, _sz, _vp]),
    "gdn_bilinear_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "gdn_bilinear_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "gdn_bicubic_down_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "gdn_bicubic_down_nchw_to_nhwc_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "gdn_maxpool2_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "gdn_maxpool2_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "gdn_maxpool2_bwd_relu_pack16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "gdn_maxpool2_fwd_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "gdn_maxpool2_bwd_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "gdn_pam_fwd_ws_bytes": (_sz, [C.POINTER(PamFwdArgs)]),
    "gdn_pam_fwd": (_i, [C.POINTER(PamFwdArgs), _vp]),
    "gdn_pam_bwd_ws_bytes": (_sz, [C.POINTER(PamBwdArgs)]),
    "gdn_pam_bwd": (_i, [C.POINTER(PamBwdArgs), _vp]),
    "gdn_cam_fwd": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "gdn_cam_bwd_ws_bytes": (_sz, [_i, _i, _i]),
    "gdn_cam_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _vp, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "gdn_cam_tc_ws_bytes": (_sz, [_i, _i, _i]),
    "gdn_cam_fwd_tc": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "gdn_cam_fwd_tc16": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "gdn_cam_bwd_tc": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _vp, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "gdn_row_softmax": (_i, [_vp, _vp, _ll, _i, _i, _vp, _vp]),
    "gdn_cam_softmax": (_i, [_vp, _vp, _i, _i, _vp]),
    "gdn_gamma_residual": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _ll, _i, _vp]),
    "gdn_rowdot": (_i, [_vp, _i, _vp, _i, _ll, _i, _vp, _vp]),
    "gdn_dot_ws_bytes": (_sz, [_ll]),
    "gdn_dot": (_i, [_vp, _i, _i, _vp, _i, _i, _ll, _i, _vp, _vp, _vp]),
    "gdn_mse": (_i, [_vp, _vp, _ll, _vp, _vp, _f, _i, _vp, _vp]),
    "gdn_l1": (_i, [_vp, _vp, _ll, _vp, _i, _vp, _f, _i, _vp, _vp]),
    "gdn_tv": (_i, [_vp, _i, _i, _i, _f, _vp, _vp, _f, _i, _vp, _vp]),
    "gdn_bce_logits": (_i, [_vp, _i, _vp, _f, _vp, _vp, _f, _vp]),
    "gdn_ssim_ws_bytes": (_sz, [_i, _i, _i]),
    "gdn_ssim": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "gdn_adamw": (_i, [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _f, _vp]),
    "gdn_adamw_multi": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _i, _f, _vp]),
    "gdn_adamw_dyn": (_i, [_vp, _vp, _vp, _vp, _ll, _vp, _f, _f, _f, _f, _f, _vp]),
    "gdn_adamw_multi_dyn": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _vp]),
    "gdn_fill": (_i, [_vp, _ll, _f, _vp]),
    "gdn_destandardise": (_i, [_vp, _vp, _vp, _vp, _ll, _ll, _f, _f, _vp]),
    "gdn_masked_spatial_mean_ws_bytes": (_sz, [_ll, _ll]),
    "gdn_masked_spatial_mean": (_i, [_vp, _vp, _ll, _ll, _f, _f, _vp, _vp, _sz, _vp]),
    "gdn_ensemble_stats": (_i, [_vp, _ll, _i, _ll, _f, _f, _vp, _vp, _vp]),
    "gdn_hist_match": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "gdn_sort_rows_ws_bytes": (_sz, [_i, _i]),
    "gdn_sort_rows": (_i, [_vp, _vp, _i, _i, _vp, _sz, _vp]),
    "gdn_bicubic_resize": (_i, [_vp, _vp, _ll, _i, _i, _i, _i, _f, _f, _vp]),
    "gdn_blend_region": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _i, _vp]),
}

_lock = threading.Lock()
_lib = None
_inited_devices = set()
launch_count = 0          # number of C-ABI compute calls made (each launches >= 1 kernel); read by bench.py


class GdnError(RuntimeError):
    pass


def load(path: str = LIB_PATH) -> C.CDLL:
    """dlopen the library and bind every declared symbol (no GPU needed)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(path):
                raise GdnError(f"{path} not found: build it with `python -m gan_danet_b200.build` (there is no CPU fallback)")
            lib = C.CDLL(path)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def lib_for_device(index: int) -> C.CDLL:
    lib = load()
    if index not in _inited_devices:
        rc = lib.gdn_init(index)
        if rc != 0:
            raise GdnError(f"gdn_init({index}) failed ({rc}): {lib.gdn_last_error().decode()}")
        _inited_devices.add(index)
    return lib


def check(rc: int, what: str = "") -> None:
    global launch_count
    launch_count += 1
    if rc != 0:
        raise GdnError(f"{what} failed ({rc}): {load().gdn_last_error().decode()}")
