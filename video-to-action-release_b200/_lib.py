"""ctypes binding of libv2a_b200.so (the C ABI declared in include/v2a_b200.h).

There is deliberately no fallback: if the shared library is missing or a call
fails, a RuntimeError is raised.  Nothing on the product path imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

V2A_MAX_SRC = 2
V2A_MAX_TAPS = 16


class IgemmSrc(C.Structure):
    _fields_ = [("hi", C.c_void_p), ("lo", C.c_void_p), ("channels", C.c_int), ("dims", C.c_int * 4)]


class IgemmTap(C.Structure):
    _fields_ = [("src", C.c_int), ("d", C.c_int * 4), ("nchunks", C.c_int)]


class IgemmDesc(C.Structure):
    _fields_ = [
        ("src", IgemmSrc * V2A_MAX_SRC),
        ("nsrc", C.c_int),
        ("taps", IgemmTap * V2A_MAX_TAPS),
        ("ntaps", C.c_int),
        ("w_hi", C.c_void_p),
        ("w_lo", C.c_void_p),
        ("wrows", C.c_int),
        ("ktot", C.c_int),
        ("out_dims", C.c_int * 4),
        ("tile_log2", C.c_int * 4),
        ("block_n", C.c_int),
        ("passes", C.c_int),
        ("cout", C.c_int),
        ("ldc", C.c_int),
        ("out_f32", C.c_void_p),
        ("out_hi", C.c_void_p),
        ("out_lo", C.c_void_p),
        ("bias", C.c_void_p),
        ("rowvec", C.c_void_p),
        ("ld_rowvec", C.c_int),
        ("rowvec_mul", C.c_int * 4),
        ("residual", C.c_void_p),
        ("ld_res", C.c_int),
        ("stats", C.c_void_p),
        ("stats_mul", C.c_int * 4),
        ("stats_ld", C.c_int),
        ("stats_replicas", C.c_int),
        ("stats_rep_stride", C.c_int64),
        ("a_fp16", C.c_int),
        ("b_fp16", C.c_int),
        ("out_pix_mul", C.c_int64 * 4),
        ("out_pix_off", C.c_int64),
        ("split_stride", C.c_int64),
    ]


V2A_WGRAD_MAX_UNITS = 80


class WgradUnit(C.Structure):
    _fields_ = [("src", C.c_int), ("d", C.c_int * 4), ("chunk", C.c_int)]


class WgradDesc(C.Structure):
    _fields_ = [
        ("src", IgemmSrc * V2A_MAX_SRC),
        ("nsrc", C.c_int),
        ("units", WgradUnit * V2A_WGRAD_MAX_UNITS),
        ("nunits", C.c_int),
        ("dy", IgemmSrc),
        ("box_log2", C.c_int * 4),
        ("cout", C.c_int),
        ("passes", C.c_int),
        ("out", C.c_void_p),
        ("ld_out", C.c_int),
        ("x_fp16", C.c_int),
    ]


class EncPrepDesc(C.Structure):
    _fields_ = [
        ("xa", C.c_void_p), ("mean_rstd_a", C.c_void_p), ("gamma_a", C.c_void_p), ("beta_a", C.c_void_p),
        ("xb", C.c_void_p), ("mean_rstd_b", C.c_void_p), ("gamma_b", C.c_void_p), ("beta_b", C.c_void_p),
        ("idn", C.c_void_p),
        ("groups", C.c_int), ("C", C.c_int), ("H", C.c_int), ("W", C.c_int), ("images", C.c_int),
        ("relu", C.c_int), ("phase_split", C.c_int), ("plane_fmt", C.c_int),
        ("out_f32", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
        ("out2_hi", C.c_void_p), ("out2_lo", C.c_void_p),
    ]


class EncGnBwdDesc(C.Structure):
    _fields_ = [
        ("dout", C.c_void_p), ("outv", C.c_void_p), ("raw", C.c_void_p), ("mean_rstd", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("mask_mode", C.c_int), ("groups", C.c_int), ("C", C.c_int), ("HW", C.c_int), ("images", C.c_int),
        ("sums", C.c_void_p), ("coef", C.c_void_p), ("d_hi", C.c_void_p), ("d_lo", C.c_void_p),
        ("g_out", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
    ]


class PrepDesc(C.Structure):
    _fields_ = [
        ("x0", C.c_void_p),
        ("x1", C.c_void_p),
        ("C0", C.c_int),
        ("C1", C.c_int),
        ("stats0", C.c_void_p),
        ("stats1", C.c_void_p),
        ("pixels_per_inst", C.c_int64),
        ("inst_per_group", C.c_int),
        ("groups", C.c_int),
        ("eps", C.c_float),
        ("gn_scratch", C.c_void_p),
        ("gamma", C.c_void_p),
        ("beta", C.c_void_p),
        ("act", C.c_int),
        ("film", C.c_void_p),
        ("pixels_per_film", C.c_int64),
        ("mode", C.c_int),
        ("H", C.c_int),
        ("W", C.c_int),
        ("P", C.c_int64),
        ("out_hi", C.c_void_p),
        ("out_lo", C.c_void_p),
        ("out_f32", C.c_void_p),
        ("raw_hi", C.c_void_p),
        ("raw_lo", C.c_void_p),
        ("stats_rep0", C.c_int),
        ("stats_rep1", C.c_int),
        ("stats_rep_stride0", C.c_int64),
        ("stats_rep_stride1", C.c_int64),
    ]


class PolicyGnDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("T", C.c_int), ("C", C.c_int), ("groups", C.c_int),
        ("eps", C.c_float),
        ("y", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("film", C.c_void_p), ("ld_film", C.c_int),
        ("addend", C.c_void_p), ("ld_add", C.c_int),
        ("out_f32", C.c_void_p), ("ld_out", C.c_int),
        ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("ld_hl", C.c_int),
        ("mean_rstd", C.c_void_p),
        ("dout", C.c_void_p), ("ld_dout", C.c_int),
        ("dy_hi", C.c_void_p), ("dy_lo", C.c_void_p),
        ("dyT_hi", C.c_void_p), ("dyT_lo", C.c_void_p),
        ("dy_f32", C.c_void_p),
        ("dbias", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
        ("dfilm", C.c_void_p), ("ld_dfilm", C.c_int),
        ("ld_T", C.c_int64),
        ("partials", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/v2a_b200.h declares
_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
SIGNATURES = {
    "v2a_last_error": (C.c_char_p, []),
    "v2a_version": (_i, []),
    "v2a_launch_count": (_i64, []),
    "v2a_igemm_plan_create": (_i, [C.POINTER(IgemmDesc), C.POINTER(_vp)]),
    "v2a_igemm_plan_run": (_i, [_vp, _vp]),
    "v2a_igemm_plan_destroy": (None, [_vp]),
    "v2a_igemm_plan_k_splits": (_i, [_vp]),
    "v2a_igemm_dual_plan_create": (_i, [C.POINTER(IgemmDesc), C.POINTER(IgemmDesc), _i, _i, _vp, C.POINTER(_vp)]),
    "v2a_igemm_dual_plan_run": (_i, [_vp, _vp]),
    "v2a_igemm_dual_plan_destroy": (None, [_vp]),
    "v2a_wgrad_plan_create": (_i, [C.POINTER(WgradDesc), C.POINTER(_vp)]),
    "v2a_wgrad_plan_run": (_i, [_vp, _vp]),
    "v2a_wgrad_plan_k_splits": (_i, [_vp]),
    "v2a_wgrad_plan_destroy": (None, [_vp]),
    "v2a_wgrad_scatter": (_i, [_vp, _i, _i, _i, _i, _vp, _i64, _i, _vp]),
    "v2a_gn_finalize": (_i, [_vp, _i, _i64, _i, _i, _i, _i64, _f, _vp, _vp]),
    "v2a_enc_stem_pack": (_i, [_vp, _f, _f, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "v2a_enc_gn_relu_maxpool": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "v2a_gather_split_fmt": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _i, _vp]),
    "v2a_enc_maxpool_relu_bwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "v2a_enc_prep": (_i, [C.POINTER(EncPrepDesc), _vp]),
    "v2a_enc_gn_bwd": (_i, [C.POINTER(EncGnBwdDesc), _vp]),
    "v2a_enc_unblock_add": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "v2a_enc_spatial_softmax_fwd": (_i, [_vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "v2a_enc_spatial_softmax_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "v2a_enc_linear_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "v2a_channel_stats": (_i, [_vp, _i64, _i64, _i, _vp, _vp]),
    "v2a_prep": (_i, [C.POINTER(PrepDesc), _vp]),
    "v2a_attention": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "v2a_linear": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "v2a_timestep_embedding": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "v2a_unet_input_pack": (_i, [_vp, C.POINTER(_i64), _vp, C.POINTER(_i64), _i, _i, _i, _i, _vp, _vp, _vp]),
    "v2a_stencil9": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "v2a_unet_output_head": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, C.POINTER(_i64), _vp]),
    "v2a_ddpm_step": (_i, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "v2a_ddim_step": (_i, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "v2a_cfg_step": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "v2a_unnormalize_clamp": (_i, [_vp, _vp, _i64, _vp]),
    "v2a_split_hl": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp]),
    "v2a_sum_slices_hl": (_i, [_vp, _i, _i64, _i64, _i, _vp, _vp, _vp]),
    "v2a_params_fingerprint": (_i, [_vp, _i, _vp, _vp]),
    "v2a_gather_split": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "v2a_policy_gn_act_fwd": (_i, [C.POINTER(PolicyGnDesc), _vp]),
    "v2a_policy_gn_act_bwd": (_i, [C.POINTER(PolicyGnDesc), _vp]),
    "v2a_policy_im2col_t": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(_i), _vp, _vp, _i64, _vp]),
    "v2a_grad_prep": (_i, [_vp, _i64, _i, _i, _vp, _vp, _i, _vp, _vp, _i64, _vp, _vp]),
    "v2a_policy_ddim_step": (_i, [_vp, _i, _vp, _i, _i64, _i, _f, _f, _f, _f, _i, _i, _vp]),
    "v2a_act_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "v2a_scatter_rows": (_i, [_vp, _i, _i64, _i, _vp, _vp, _vp]),
    "v2a_add_strided": (_i, [_vp, _i, _vp, _i, _i64, _i, _i, _vp]),
    "v2a_grad_sumsq": (_i, [_vp, _i64, _vp, _vp]),
    "v2a_adamw_ema_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _f, _f, _f, _i, _f, _vp]),
    "v2a_pr_layernorm": (_i, [_vp, _i64, _i, _vp, _i, _i, _i, _i, _vp, _vp, _f, _vp, _i64, _i, _vp]),
    "v2a_pr_l2norm_scale": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "v2a_pr_attention": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _i, _vp]),
    "v2a_pr_token_mean": (_i, [_vp, _i64, _i, _i, _i, _i, _vp, _i, _vp]),
    "v2a_pr_broadcast_rows": (_i, [_vp, _i, _i, _i, _vp, _i64, _i, _vp]),
    "v2a_replay_gather_images": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "v2a_replay_gather_actions": (_i, [_vp, _i, _i, _i, _vp, _vp]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the .so is absent) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        _build.build_extension()
    if not os.path.exists(path):
        raise RuntimeError(f"v2a_b200: CUDA extension {path} is missing and could not be built")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().v2a_last_error()
        raise RuntimeError(f"v2a_b200 {what} failed (rc={rc}): {msg.decode() if msg else '?'}")
