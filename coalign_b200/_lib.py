"""ctypes binding of libcoalign_b200.so (include/coalign_b200.h).  No fallback: a missing library or a
non-B200 device raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcoalign_b200.so")

CB_MAX_KSTEPS = 168
CB_MAX_AGENTS = 64
CB_OUT_PF, CB_OUT_PS, CB_OUT_UPSAMPLE, CB_OUT_HEADS = 0, 1, 2, 3
(CB_OPT_NO_PDL, CB_OPT_EPI_DIRECT, CB_OPT_TMA_STORE, CB_OPT_HALO_BO, CB_OPT_FUSE_VERSION, CB_OPT_FUSE_BLEND_FP32,
 CB_OPT_FUSE_OCC3, CB_OPT_CONV_DEBUG) = range(8)


class KStep(C.Structure):
    _fields_ = [("row_off", C.c_int32), ("w_k", C.c_int32), ("col", C.c_uint16), ("a_sel", C.c_uint16)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("a_ptr", C.c_void_p * 2), ("a_rows", C.c_int64 * 2), ("a_pitch", C.c_int32 * 2),
        ("w_ptr", C.c_void_p), ("w_rows", C.c_int32), ("w_k_total", C.c_int32),
        ("n_img", C.c_int32), ("Hp", C.c_int32), ("Wp", C.c_int32),
        ("n_total", C.c_int32), ("block_n", C.c_int32),
        ("bias", C.c_void_p), ("cout_mod", C.c_int32), ("relu", C.c_int32),
        ("residual", C.c_void_p), ("res_pitch", C.c_int32), ("res_lo_off", C.c_int64),
        ("out", C.c_void_p), ("out_pitch", C.c_int32), ("out_ch_off", C.c_int32), ("out_lo_off", C.c_int64),
        ("out_mode", C.c_int32), ("up_k", C.c_int32), ("out_Hp", C.c_int32), ("out_Wp", C.c_int32),
        ("out_plane_rows", C.c_int64),
        ("head_out", C.c_void_p * 4), ("head_c0", C.c_int32 * 4), ("head_cn", C.c_int32 * 4), ("n_heads", C.c_int32),
        ("n_ksteps", C.c_int32), ("ksteps", KStep * CB_MAX_KSTEPS),
        ("in_pad", C.c_int32),
    ]


CB_WGRAD_MAX_BOXES, CB_WGRAD_MAX_UNITS = 4, 28


class WgradBox(C.Structure):
    _fields_ = [("row_off", C.c_int32), ("out_ld", C.c_int32), ("out_off", C.c_int64), ("col", C.c_uint16),
                ("x_sel", C.c_uint16), ("pad_", C.c_uint32)]


class WgradUnit(C.Structure):
    _fields_ = [("m0", C.c_int32), ("a_row_off", C.c_int32), ("m_valid", C.c_int32), ("n_boxes", C.c_int32),
                ("box", WgradBox * CB_WGRAD_MAX_BOXES)]


class WgradDesc(C.Structure):
    _fields_ = [("dz_ptr", C.c_void_p), ("dz_lo_ptr", C.c_void_p), ("dz_pitch", C.c_int32), ("k_splits", C.c_int32),
                ("rows_total", C.c_int64), ("x_ptr", C.c_void_p * 2), ("x_rows", C.c_int64 * 2), ("x_pitch", C.c_int32 * 2),
                ("x_lo_rows", C.c_int32 * 2), ("dw", C.c_void_p), ("n_units", C.c_int32), ("pad_", C.c_int32),
                ("units", WgradUnit * CB_WGRAD_MAX_UNITS)]


class PackJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("s_r1", C.c_int64), ("s_r0", C.c_int64), ("s_k1", C.c_int64),
                ("s_k0", C.c_int64), ("first", C.c_int64), ("R0", C.c_int32), ("K0", C.c_int32), ("rows", C.c_int32),
                ("K", C.c_int32), ("dst_ld", C.c_int32), ("k_off", C.c_int32), ("lo_col_off", C.c_int32), ("pad_", C.c_int32)]


class Map(C.Structure):
    _fields_ = [("n_img", C.c_int32), ("Hp", C.c_int32), ("Wp", C.c_int32), ("c_total", C.c_int32), ("c_mod", C.c_int32),
                ("y_mode", C.c_int32), ("y_pitch", C.c_int32), ("y_ch_off", C.c_int32), ("up_k", C.c_int32),
                ("y_Hp", C.c_int32), ("y_Wp", C.c_int32), ("z_at_y", C.c_int32), ("z_pitch", C.c_int32),
                ("y_plane_rows", C.c_int64)]


_P, _I, _L, _F, _D = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

EXPORTS = {
    "cb_wgrad": (C.c_int, [C.POINTER(WgradDesc), _I, _P]),
    "cb_wgrad_simt": (C.c_int, [C.POINTER(WgradDesc), _P]),
    "cb_bn_stats": (C.c_int, [_P, _L, C.POINTER(Map), _P, _P]),
    "cb_bn_finalize": (C.c_int, [_P, _I, _D, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cb_bn_stats_finalize": (C.c_int, [_P, _L, C.POINTER(Map), _P, _P, _D, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cb_bn_apply": (C.c_int, [_P, _L, _P, _P, _P, _L, _P, _P, _P, C.c_int32, _L, _I, C.POINTER(Map), _P, _L, _P]),
    "cb_bn_bwd_reduce": (C.c_int, [_P, _L, _P, _L, _I, _P, _L, _P, _P, _P, _P, C.POINTER(Map), _P, _P]),
    "cb_bn_bwd_apply": (C.c_int, [_P, _L, _P, _L, _I, _P, _L, _P, _P, _P, _P, _P, _P, _D, C.POINTER(Map), _P, _L, _P, _L, _P,
                                  _P, _P]),
    "cb_heads_grad_pack": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _L, _P, _P]),
    "cb_warp_att_fuse_bwd": (C.c_int, [_P, _I, _L, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _L, _P, _P]),
    "cb_grad_combine": (C.c_int, [_P, _P, _L, _I, _I, _I, _I, _I, _I, _P, _L, _P]),
    "cb_pfn_train_stats": (C.c_int, [_P, _P, _P, _I, _P, _I, _P, _P, _P, _P]),
    "cb_pfn_train_finalize": (C.c_int, [_P, _P, _I, _I, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    "cb_pfn_bwd": (C.c_int, [_P, _P, _P, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P, _P]),
    "cb_pfn_bwd_finalize": (C.c_int, [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "cb_pack_weight": (C.c_int, [_P, _I, _I, _I, _I, _L, _L, _L, _L, _P, _I, _I, _I, _P]),
    "cb_pack_weights_batch": (C.c_int, [_P, _I, _L, _P]),
    "cb_permute_f32": (C.c_int, [_P, _I, _I, _I, _I, _L, _L, _L, _L, _F, _P, _P]),
    "cb_adam_step": (C.c_int, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _F, _P, _I, _P]),
    "cb_version": (C.c_int, []),
    "cb_set_option": (C.c_int, [C.c_int, C.c_int]),
    "cb_get_option": (C.c_int, [C.c_int]),
    "cb_device_check": (C.c_int, []),
    "cb_voxelize_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "cb_voxelize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "cb_pfn_scatter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "cb_canvas_clear": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "cb_points_to_canvas": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    "cb_points_to_canvas_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                          C.c_void_p]),
    "cb_upload_i32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "cb_conv_gemm": (C.c_int, [C.POINTER(ConvDesc), C.c_int, C.c_void_p]),
    "cb_conv_gemm_pair": (C.c_int, [C.POINTER(ConvDesc), C.c_int, C.c_void_p]),
    "cb_conv_gemm_t": (C.c_int, [C.POINTER(ConvDesc), C.c_int, C.c_void_p]),
    "cb_conv_gemm_halo": (C.c_int, [C.POINTER(ConvDesc), C.c_int, C.c_void_p]),
    "cb_conv_gemm_t_halo": (C.c_int, [C.POINTER(ConvDesc), C.c_int, C.c_void_p]),
    "cb_conv_gemm_simt": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "cb_normalize_affine": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                      C.c_void_p]),
    "cb_warp_att_fuse": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "cb_postprocess_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "cb_postprocess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "cb_postprocess_stage1": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_void_p]),
    "cb_pointpillar_loss_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "cb_pointpillar_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    "cb_ps_to_pf": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                              C.c_void_p]),
    "cb_nchw_to_layout": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                    C.c_void_p]),
    "cb_layout_to_nchw": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p]),
    "cb_nchw_to_ps_pad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                    C.c_void_p]),
    "cb_nhwc_to_ps_pad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                    C.c_void_p]),
    "cb_lift_splat": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 6 + [C.c_void_p] * 5),
    "cb_upsample_concat": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_int64, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


def load(check_device: bool = False):
    """Load the CUDA library (raises if it was not built).  With check_device, also require sm_100."""
    global _lib
    if _lib is None:
        path = os.environ.get("COALIGN_B200_LIB", LIB_PATH)            # override: A/B timing of two builds (development)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the .so is git-ignored; coalign_b200 has no CPU fallback)")
        lib = C.CDLL(path)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    if check_device:
        rc = _lib.cb_device_check()
        if rc != 0:
            raise RuntimeError(f"coalign_b200 needs a B200 (sm_100) CUDA device, cb_device_check() = {rc}")
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        kind = "cudaError" if rc > 0 else "argument/driver error"
        raise RuntimeError(f"libcoalign_b200 {what} failed: {kind} {rc}")
