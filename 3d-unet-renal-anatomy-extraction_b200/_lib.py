"""ctypes binding of libunet3d_b200.so (the C ABI declared in include/unet3d_b200.h).

There is no CPU fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libunet3d_b200.so")

MAX_SRC = 8


class Src(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int), ("W", C.c_int), ("H", C.c_int), ("D", C.c_int), ("N", C.c_int),
                ("sW", C.c_longlong), ("sH", C.c_longlong), ("sD", C.c_longlong), ("sN", C.c_longlong)]


class ConvArgs(C.Structure):
    _fields_ = [("n_src", C.c_int), ("src", Src * MAX_SRC), ("tab", C.c_void_p), ("w", C.c_void_p), ("out", C.c_void_p),
                ("out2", C.c_void_p), ("bias", C.c_void_p), ("addend", C.c_void_p), ("addend2", C.c_void_p),
                ("stats", C.c_void_p), ("err", C.c_void_p),
                ("N", C.c_int), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("Dt", C.c_int), ("n_nblk", C.c_int), ("nblk", C.c_int), ("G", C.c_int), ("n_cg", C.c_int),
                ("n_taps", C.c_int), ("fuse", C.c_int), ("nbuf", C.c_int), ("wT", C.c_int), ("w_stages", C.c_int), ("in_f16", C.c_int), ("out_f16", C.c_int),
                ("out_sN", C.c_longlong), ("out_sD", C.c_longlong), ("out_sH", C.c_longlong), ("out_sW", C.c_longlong),
                ("out_C", C.c_int), ("stats_C", C.c_int), ("omul", C.c_int), ("zD", C.c_int), ("zH", C.c_int),
                ("zW", C.c_int), ("act", C.c_int), ("a_stages", C.c_int), ("dense", C.c_int), ("dbg_out", C.c_void_p)]


class GatherJob(C.Structure):
    _fields_ = [("src0", C.c_void_p), ("src1", C.c_void_p), ("idx", C.c_void_p), ("out", C.c_void_p), ("n", C.c_longlong),
                ("n0", C.c_int), ("mode", C.c_int)]


class UnpackJob(C.Structure):
    _fields_ = [("dw", C.c_void_p), ("rowmap", C.c_void_p), ("out", C.c_void_p), ("k3", C.c_int), ("Kp", C.c_int),
                ("Np", C.c_int), ("Ncols", C.c_int), ("col_stride", C.c_longlong)]


class WgradArgs(C.Structure):
    _fields_ = [("n_src", C.c_int), ("src", Src * 16), ("box_w", C.c_int * 16), ("box_h", C.c_int * 16), ("box_c", C.c_int * 16),
                ("tab", C.c_void_p), ("dw", C.c_void_p), ("err", C.c_void_p),
                ("N", C.c_int), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("n_jobs", C.c_int), ("job_stride", C.c_int), ("split", C.c_int), ("x_f16", C.c_int)]


_SIGS = {
    "unet3d_version": (C.c_char_p, []),
    "unet3d_last_error_string": (C.c_char_p, []),
    "unet3d_num_sms": (C.c_int, []),
    "unet3d_set_sm_limit": (C.c_int, [C.c_int]),
    "unet3d_conv_gemm": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "unet3d_conv_gemm_smem_bytes": (C.c_size_t, [C.c_int] * 7),
    "unet3d_weight_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "unet3d_gather_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "unet3d_dw_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "unet3d_wgrad_gemm": (C.c_int, [C.POINTER(WgradArgs), C.c_void_p]),
    "unet3d_in_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_void_p]),
    "unet3d_in_apply": (C.c_int, [C.c_void_p] * 5 + [C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "unet3d_in_bwd_reduce": (C.c_int, [C.c_void_p] * 8 + [C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "unet3d_in_bwd_apply": (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 7 + [C.c_void_p]),
    "unet3d_in_bwd_small": (C.c_int, [C.c_void_p] * 8 + [C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "unet3d_channel_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "unet3d_stem_fwd": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 7 + [C.c_void_p]),
    "unet3d_stem_wgrad": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "unet3d_head_fwd": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "unet3d_head_bwd": (C.c_int, [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "unet3d_loss_fwd": (C.c_int, [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_longlong, C.c_float, C.c_void_p]),
    "unet3d_loss_bwd": (C.c_int, [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_longlong, C.c_float, C.c_int, C.c_void_p]),
    "unet3d_sw_accumulate": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 11 + [C.c_void_p]),
    "unet3d_sw_finalize": (C.c_int, [C.c_void_p] * 3 + [C.c_int, C.c_longlong, C.c_void_p]),
    "unet3d_aug_flip": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]),
    "unet3d_aug_stats": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "unet3d_aug_affine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "unet3d_aug_gamma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_float, C.c_float, C.c_void_p]),
    "unet3d_att_gate_fwd": (C.c_int, [C.c_void_p] * 3 + [C.c_longlong, C.c_int, C.c_void_p]),
    "unet3d_att_gate_bwd": (C.c_int, [C.c_void_p] * 6 + [C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "unet3d_att_mid_bwd": (C.c_int, [C.c_void_p] * 6 + [C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "unet3d_maxpool3d_fwd": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 6 + [C.c_void_p]),
    "unet3d_maxpool3d_bwd": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p]),
    "unet3d_ccl_label": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p]),
    "unet3d_ccl_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "unet3d_region_accumulate": (C.c_int, [C.c_void_p] * 3 + [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_longlong),
                                           C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p]),
    "unet3d_overlap_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "unet3d_merge_finalize": (C.c_int, [C.c_void_p] * 3 + [C.c_int, C.c_longlong, C.c_void_p]),
    "unet3d_zoom_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "unet3d_zoom_linear": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int),
                                     C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_longlong),
                                     C.POINTER(C.c_float), C.c_void_p, C.c_size_t, C.c_void_p]),
    "unet3d_zoom_label": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_longlong),
                                    C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None


class Unet3dError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built: there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Unet3dError(f"{LIB_PATH} not found: build it with __graft_entry__.build() "
                              f"(or csrc/build.sh); this package has no CPU / PyTorch fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def exported_symbols():
    return list(_SIGS)


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().unet3d_last_error_string().decode()
        raise Unet3dError(f"{what} failed with code {rc}: {msg}")
