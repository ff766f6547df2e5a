"""ctypes binding of libdpf_sm100.so -- the C ABI declared in include/dpf_sm100.h.

There is no fallback of any kind: if the shared library is missing, or the device is not an sm_100a part,
using the ops raises.  (`python __graft_entry__.py` or `make -C dualpixelface_b200/csrc` builds it.)
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("DPF_SM100_LIB", _HERE / "libdpf_sm100.so"))

c_void_p, c_int, c_float, c_ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
c_int_p = C.POINTER(C.c_int)
c_float_p = C.POINTER(C.c_float)


class ConvArgs(C.Structure):
    """struct dpf_conv3d_args (include/dpf_sm100.h)."""
    _fields_ = [
        ("kind", c_int), ("B", c_int), ("D", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int),
        ("x", c_void_p), ("w", c_void_p), ("y", c_void_p),
        ("y_f32", c_int), ("y_cstride", c_int), ("y_coff", c_int),
        ("scale", c_void_p), ("shift", c_void_p), ("residual", c_void_p),
        ("relu", c_int), ("stats", c_void_p),
        ("x_cstride", c_int), ("x_coff", c_int), ("res_pre", c_int), ("slope", c_float),
    ]


# name -> (restype, argtypes); every symbol of include/dpf_sm100.h
SIGNATURES = {
    "dpf_abi_version": (c_int, []),
    "dpf_last_error": (C.c_char_p, []),
    "dpf_device_check": (c_int, []),
    "dpf_launch_count": (c_ll, []),
    "dpf_costvol_fwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int_p, c_void_p]),
    "dpf_costvol_bwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int_p, c_void_p]),
    "dpf_asm_sample_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dpf_asm_blend_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dpf_channel_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p]),
    "dpf_channel_stats_ws_floats": (c_ll, [c_int, c_ll, c_int]),
    "dpf_conv3d_fwd": (c_int, [C.POINTER(ConvArgs), c_void_p]),
    "dpf_conv3d_weight_elems": (c_ll, [c_int, c_int, c_int]),
    "dpf_regress_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "dpf_regress_fwd_halfpixel": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "dpf_regress_fwd_tile": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 8 + [c_float, c_float, c_void_p]),
    "dpf_regress_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "dpf_anm_select": (c_int, [c_void_p, c_void_p, c_void_p, c_float_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dpf_anm_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dpf_bias_act": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_float, c_void_p]),
    "dpf_anm_tail": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dpf_affine_act": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_float, c_void_p]),
    "dpf_bn_bwd_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_float, c_void_p]),
    "dpf_bn_bwd_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_float, c_void_p]),
    "dpf_asm_blend_bwd": (c_int, [c_void_p] * 7 + [c_int] * 10 + [c_void_p]),
    "dpf_asm_sample_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dpf_conv3d_wgrad": (c_int, [c_int, c_void_p, c_void_p, c_void_p] + [c_int] * 10 + [c_void_p]),
    "dpf_dcn3d_fwd": (c_int, [c_void_p] * 6 + [c_int] * 9 + [c_void_p]),
    "dpf_dcn3d_bwd_data": (c_int, [c_void_p] * 6 + [c_int] * 7 + [c_void_p]),
    "dpf_dcn3d_bwd_weight": (c_int, [c_void_p] * 4 + [c_int] * 6 + [c_void_p]),
    "dpf_conv2d_fwd": (c_int, [c_void_p] * 6 + [c_int] * 11 + [c_float, c_void_p]),
    "dpf_conv3d_s2_fwd": (c_int, [c_void_p] * 5 + [c_int] * 11 + [c_void_p]),
    "dpf_conv2d_tc_npad": (c_int, [c_int]),
    "dpf_conv2d_tc_weight_elems": (c_ll, [c_int, c_int]),
    "dpf_conv2d_tc_fwd": (c_int, [c_void_p] * 6 + [c_int] * 11 + [c_float, c_void_p]),
    "dpf_anm_tail_tile": (c_int, [c_void_p, c_void_p] + [c_int] * 9 + [c_void_p]),
    "dpf_fused_losses_ws_floats": (c_ll, [c_ll]),
    "dpf_fused_losses": (c_int, [c_void_p, c_int] + [c_void_p] * 8 + [c_int] * 3 + [c_void_p]),
    "dpf_channel_max": (c_int, [c_void_p, c_void_p, C.c_longlong, c_int, c_void_p]),
    "dpf_fpn_merge": (c_int, [c_void_p] * 4 + [c_int] * 6 + [c_void_p]),
    "dpf_pyramid_cat": (c_int, [c_void_p] * 4 + [c_int] * 8 + [c_void_p]),
    "dpf_pyramid_cat_tile": (c_int, [c_void_p] * 4 + [c_int] * 10 + [c_void_p]),
    "dpf_anm_tail_bwd": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p]),
    "dpf_anm_gather_bwd": (c_int, [c_void_p] * 4 + [c_int] * 7 + [c_void_p]),
    "dpf_conv3d_head_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float] + [c_int] * 5 + [c_void_p]),
    "dpf_stem_conv_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dpf_bn_fwd_coefs": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dpf_bn_bwd_coefs": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dpf_softargmin_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_float, c_float, c_void_p]),
}

_lib = None


class DpfError(RuntimeError):
    pass


def load(check_device: bool = False) -> C.CDLL:
    """Load the shared library and bind every declared symbol (raises if any is missing)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.is_file():
            raise DpfError(f"{LIB_PATH} not found: build it with `make -C {_HERE / 'csrc'}` "
                           f"(there is no CPU or PyTorch fallback for the hot path)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.dpf_abi_version() != 1:
            raise DpfError(f"ABI version mismatch: library {lib.dpf_abi_version()} != binding 1")
        _lib = lib
    if check_device and _lib.dpf_device_check() != 0:
        raise DpfError(_lib.dpf_last_error().decode())
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        raise DpfError(f"{what}: {_lib.dpf_last_error().decode()}")
