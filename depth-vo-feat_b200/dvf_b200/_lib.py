"""ctypes binding of libdvf_b200.so (C ABI declared in include/dvf_b200.h).

There is no fallback: if the shared library is missing it is built with nvcc (dvf_b200.build);
if that is impossible the import fails loudly.  Nothing here touches the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

DVF_MAX_VIEWS = 4
DVF_MAX_LEVELS = 6

F32, BF16 = 0, 1
NCHW, NHWC = 0, 1
PADDING = {"zeros": 0, "border": 1}
ROTATION = {"euler": 0, "quat": 1}
FLAG_ALIGN_CORNERS, FLAG_ZERO_GSRC, FLAG_NAN_CHECK, FLAG_NO_TMA, FLAG_PDL, FLAG_DISPARITY, FLAG_PDL_CHAINED = 1, 2, 4, 8, 16, 32, 64
FLAG_REF_CUDA = 128   # dvf_inverse_warp_fwd: torch-CUDA's per-pixel rounding
ABI_VERSION = 2


class DvfError(RuntimeError):
    pass


class dvf_desc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("dtype", C.c_int32), ("layout", C.c_int32), ("padding", C.c_int32), ("flags", C.c_int32)]


class dvf_level(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32),
                ("depth", C.c_void_p), ("tgt", C.c_void_p), ("src", C.c_void_p * DVF_MAX_VIEWS),
                ("expl", C.c_void_p), ("expl_bstride", C.c_int64),
                ("P", C.c_void_p), ("Kinv", C.c_void_p),
                ("gdepth", C.c_void_p), ("gexpl", C.c_void_p), ("gsrc", C.c_void_p * DVF_MAX_VIEWS),
                ("gtgt", C.c_void_p), ("gP", C.c_void_p)]


class dvf_pose_args(C.Structure):
    _fields_ = [("vec", C.c_void_p), ("K", C.c_void_p), ("Kinv", C.c_void_p), ("downscale", C.POINTER(C.c_float)),
                ("rotation", C.c_int32), ("reserved", C.c_int32), ("gvec", C.c_void_p)]


class dvf_reg_level(C.Structure):
    _fields_ = [("x", C.c_void_p), ("g", C.c_void_p), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("weight", C.c_float)]


class dvf_loss_desc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("V", C.c_int32), ("n_levels", C.c_int32),
                ("dtype", C.c_int32), ("layout", C.c_int32), ("padding", C.c_int32), ("flags", C.c_int32),
                ("mean_batch", C.c_int32), ("grad_dtype", C.c_int32), ("piece_overhead", C.c_int32), ("ctas_per_sm", C.c_int32),
                ("upstream", C.c_void_p), ("nan_flags", C.c_void_p),
                ("n_peers", C.c_int32), ("peer_rank", C.c_int32), ("peer_terms", C.POINTER(C.c_void_p)),
                ("disp_eps", C.c_float), ("img_scale", C.c_float)]


_vp, _i32, _sz, _fp = C.c_void_p, C.c_int32, C.c_size_t, C.POINTER(C.c_float)

# name -> (restype, argtypes); one entry per symbol of include/dvf_b200.h
SIGNATURES = {
    "dvf_version": (C.c_int, []),
    "dvf_strerror": (C.c_char_p, [C.c_int]),
    "dvf_pose_proj_fwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _fp, _i32, _vp, _vp, _vp, _vp]),
    "dvf_pose_proj_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _fp, _i32, _vp, _vp]),
    "dvf_pixel2cam": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "dvf_cam2pixel": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "dvf_pixel2cam_bwd": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "dvf_cam2pixel_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "dvf_inverse_warp_fwd": (C.c_int, [C.POINTER(dvf_desc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dvf_inverse_warp_bwd_workspace_bytes": (_sz, [C.POINTER(dvf_desc)]),
    "dvf_inverse_warp_bwd": (C.c_int, [C.POINTER(dvf_desc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dvf_photo_loss_workspace_bytes": (_sz, [C.POINTER(dvf_loss_desc), C.POINTER(dvf_level)]),
    "dvf_photo_loss_fused": (C.c_int, [C.POINTER(dvf_loss_desc), C.POINTER(dvf_level), _vp, _vp, _sz, _vp]),
    "dvf_photo_loss_fused_pose": (C.c_int, [C.POINTER(dvf_loss_desc), C.POINTER(dvf_level), C.POINTER(dvf_pose_args), _vp, _vp,
                                            _sz, _vp]),
    "dvf_area_pyramid": (C.c_int, [_vp, _i32, _i32, _i32, _i32, C.POINTER(_vp), _vp]),
    "dvf_area_downsample": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "dvf_transpose_planes": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "dvf_transpose_planes_scaled": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "dvf_reg_workspace_bytes": (_sz, [C.POINTER(dvf_reg_level), _i32]),
    "dvf_smooth_loss": (C.c_int, [C.POINTER(dvf_reg_level), _i32, _vp, _vp, _sz, _vp]),
    "dvf_explainability_loss": (C.c_int, [C.POINTER(dvf_reg_level), _i32, _vp, _vp, _sz, _vp]),
    "dvf_se3_exp_fwd": (C.c_int, [_vp, _i32, _vp, _vp]),
    "dvf_se3_exp_bwd": (C.c_int, [_vp, _vp, _i32, _vp, _vp]),
    "dvf_caffe_geo_fwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "dvf_caffe_geo_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "dvf_caffe_pinhole_fwd": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "dvf_caffe_pinhole_bwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "dvf_caffe_warp_fwd": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "dvf_caffe_warp_bwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "dvf_caffe_abs_loss": (C.c_int, [_vp, _vp, C.c_uint64, _i32, C.c_float, _vp, _vp, _vp, _vp, _vp]),
    "dvf_caffe_edge_smooth_loss": (C.c_int, [_vp, _vp, _i32, _i32, _i32, C.c_float, _vp, _vp, _vp, _vp]),
    "dvf_ssim_loss": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "dvf_torch_sincos": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp]),
    "dvf_selftest_fast_div": (C.c_int, [C.c_uint64, C.c_uint64, _i32, _vp, _vp]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if needed) libdvf_b200.so and declare every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path) or (os.environ.get("DVF_REBUILD") == "1"):
        path = _build.build(force=os.environ.get("DVF_REBUILD") == "1")
    try:
        lib = C.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise DvfError(f"cannot load {path}: {e}.  The CUDA library is the only implementation of this "
                       f"package; build it with `python -m dvf_b200.build` (needs nvcc).") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so is stale: loud by design
        fn.restype = res
        fn.argtypes = args
    if lib.dvf_version() != ABI_VERSION:
        raise DvfError(f"{path}: ABI version {lib.dvf_version()} != {ABI_VERSION}; rebuild with DVF_REBUILD=1")
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().dvf_strerror(status).decode()
        raise DvfError(f"{what} failed: {msg} (status {status})")
