"""ctypes/numpy front-end of the CPU ORACLE (oracle/dvf_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / --impl reference legs.  The product package
(depth-vo-feat_b200/) must never import this module.

All arrays are C-contiguous float32 numpy arrays (images NCHW).  Function
names mirror the reference's (pytorch_version/inverse_warp.py,
loss_functions*.py); see the C file for the line-by-line citations.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdvf_oracle.so")

PAD = {"zeros": 0, "border": 1, "zeros_align": 2, "border_align": 3,   # *_align: grid_sample(align_corners=True)
       "zeros_cuda": 4, "border_cuda": 5}   # *_cuda: torch-CUDA's per-pixel rounding (forward only, DVFO_REF_CUDA)
ROT = {"euler": 0, "quat": 1}


def build(force: bool = False) -> str:
    """Compile oracle/dvf_oracle.c with gcc (Makefile recipe)."""
    src = os.path.join(_HERE, "dvf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libdvf_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.dvfo_smooth_loss.restype = C.c_double
        _lib.dvfo_ssim_loss.restype = C.c_double
        _lib.dvfo_explainability_loss.restype = C.c_double
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def torch_trig(x):
    """torch.sin / torch.cos (CPU fp32) of x as restated in oracle/torch_trig.h -> (sin, cos)"""
    x = _f32(x).reshape(-1)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().dvfo_torch_trig(_p(x), C.c_long(x.size), _p(s), _p(c))
    return s, c


def pose_vec2mat(vec, rotation_mode="euler"):
    vec = _f32(vec)
    n = vec.shape[0]
    out = np.empty((n, 3, 4), np.float32)
    lib().dvfo_pose_vec2mat(_p(vec), n, ROT[rotation_mode], _p(out))
    return out


def project(K, posemat):
    K, posemat = _f32(K), _f32(posemat)
    n = K.shape[0]
    out = np.empty((n, 3, 4), np.float32)
    lib().dvfo_project(_p(K), _p(posemat), n, _p(out))
    return out


def scale_intrinsics(K, Kinv, downscale):
    K, Kinv = _f32(K), _f32(Kinv)
    Ks, Kinvs = np.empty_like(K), np.empty_like(Kinv)
    lib().dvfo_scale_intrinsics(_p(K), _p(Kinv), K.shape[0], C.c_float(downscale), _p(Ks), _p(Kinvs))
    return Ks, Kinvs


def pixel2cam(depth, Kinv):
    depth, Kinv = _f32(depth), _f32(Kinv)
    B, H, W = depth.shape
    out = np.empty((B, 3, H, W), np.float32)
    lib().dvfo_pixel2cam(_p(depth), _p(Kinv), B, H, W, _p(out))
    return out


def grid(depth, P, Kinv, padding_mode="zeros"):
    depth, P, Kinv = _f32(depth), _f32(P), _f32(Kinv)
    B, H, W = depth.shape
    out = np.empty((B, H, W, 2), np.float32)
    lib().dvfo_grid(_p(depth), _p(P), _p(Kinv), B, H, W, PAD[padding_mode], _p(out))
    return out


def grid_sample(img, grid_, padding_mode="zeros"):
    img, grid_ = _f32(img), _f32(grid_)
    B, Cc, H, W = img.shape
    out = np.empty_like(img)
    lib().dvfo_grid_sample(_p(img), _p(grid_), B, Cc, H, W, PAD[padding_mode], _p(out))
    return out


def inverse_warp_P(img, depth, P, Kinv, padding_mode="zeros"):
    """inverse_warp given P = K @ pose_vec2mat(pose).  Returns (warped, valid)."""
    img, depth, P, Kinv = _f32(img), _f32(depth), _f32(P), _f32(Kinv)
    B, Cc, H, W = img.shape
    out = np.empty_like(img)
    valid = np.empty((B, H, W), np.uint8)
    lib().dvfo_inverse_warp_fwd(_p(img), _p(depth), _p(P), _p(Kinv), B, Cc, H, W,
                                PAD[padding_mode], _p(out), _p(valid))
    return out, valid


def inverse_warp(img, depth, pose, K, Kinv, rotation_mode="euler", padding_mode="zeros"):
    """Reference signature (inverse_warp.py:160).  Returns warped only."""
    P = project(K, pose_vec2mat(pose, rotation_mode))
    return inverse_warp_P(img, depth, P, Kinv, padding_mode)[0]


def inverse_warp_bwd_P(gout, img, depth, P, Kinv, padding_mode="zeros", need_gimg=True):
    gout, img, depth, P, Kinv = map(_f32, (gout, img, depth, P, Kinv))
    B, Cc, H, W = img.shape
    gimg = np.empty_like(img) if need_gimg else None
    gdepth = np.empty((B, H, W), np.float32)
    gP = np.empty((B, 3, 4), np.float32)
    lib().dvfo_inverse_warp_bwd(_p(gout), _p(img), _p(depth), _p(P), _p(Kinv), B, Cc, H, W,
                                PAD[padding_mode], _p(gimg), _p(gdepth), _p(gP))
    return gimg, gdepth, gP


def pose_bwd(gP, K, vec, rotation_mode="euler", need_gK=False):
    gP, K, vec = _f32(gP), _f32(K), _f32(vec)
    n = vec.shape[0]
    gvec = np.empty((n, 6), np.float32)
    gK = np.empty((n, 3, 3), np.float32) if need_gK else None
    lib().dvfo_pose_bwd(_p(gP), _p(K), _p(vec), n, ROT[rotation_mode], _p(gvec), _p(gK))
    return (gvec, gK) if need_gK else gvec


def photo_loss_P(tgt, srcs, depth, P, Kinv, expl=None, padding_mode="zeros",
                 need_gsrc=False, need_gtgt=False):
    """One-scale masked L1 reconstruction loss over V source views given
    P[B,V,3,4]; gradients are for upstream dL/dterm = 1.
    Returns dict(terms[V] float64, gdepth, gP[B,V,3,4], gexpl, gsrc[list], gtgt, valid[V,B,H,W])."""
    tgt, depth, P, Kinv = map(_f32, (tgt, depth, P, Kinv))
    srcs = [_f32(s) for s in srcs]
    B, Cc, H, W = tgt.shape
    V = len(srcs)
    assert P.shape == (B, V, 3, 4)
    expl = None if expl is None else _f32(expl)
    terms = np.zeros(V, np.float64)
    gdepth = np.empty((B, H, W), np.float32)
    gP = np.empty((B, V, 3, 4), np.float32)
    gexpl = None if expl is None else np.empty((B, V, H, W), np.float32)
    gsrc = [np.empty_like(tgt) for _ in range(V)] if need_gsrc else None
    gtgt = np.empty_like(tgt) if need_gtgt else None
    valid = np.empty((V, B, H, W), np.uint8)
    src_arr = (C.c_void_p * V)(*[s.ctypes.data for s in srcs])
    gsrc_arr = (C.c_void_p * V)(*[g.ctypes.data for g in gsrc]) if need_gsrc else None
    lib().dvfo_photo_loss(_p(tgt), src_arr, _p(depth), _p(P), _p(Kinv), _p(expl), B, Cc, H, W, V,
                          PAD[padding_mode], _p(terms), _p(gdepth), _p(gP), _p(gexpl),
                          gsrc_arr, _p(gtgt), _p(valid))
    return dict(terms=terms, gdepth=gdepth, gP=gP, gexpl=gexpl, gsrc=gsrc, gtgt=gtgt, valid=valid)


def area_downsample(img, h, w):
    img = _f32(img)
    B, Cc, H, W = img.shape
    out = np.empty((B, Cc, h, w), np.float32)
    lib().dvfo_area_downsample(_p(img), B * Cc, H, W, h, w, _p(out))
    return out


def smooth_loss_one(d, need_grad=False):
    d = _f32(d)
    B, H, W = d.shape[0], d.shape[-2], d.shape[-1]
    g = np.empty_like(d) if need_grad else None
    val = lib().dvfo_smooth_loss(_p(d), B, H, W, _p(g))
    return (val, g) if need_grad else val


def explainability_loss_one(mask, need_grad=False):
    mask = _f32(mask)
    g = np.empty_like(mask) if need_grad else None
    val = lib().dvfo_explainability_loss(_p(mask), C.c_size_t(mask.size), _p(g))
    return (val, g) if need_grad else val


def se3_exp(vec6):
    """se3_generate.py forward: [B,6] (w,u) -> [B,4,4] float64."""
    v = _f32(vec6).reshape(-1, 6)
    out = np.empty((v.shape[0], 4, 4), np.float64)
    lib().dvfo_se3_exp_fwd(_p(v), v.shape[0], _p(out))
    return out


def se3_exp_bwd(vec6, gout):
    v = _f32(vec6).reshape(-1, 6)
    g = np.ascontiguousarray(gout, dtype=np.float64).reshape(-1, 4, 4)
    gin = np.empty((v.shape[0], 6), np.float32)
    lib().dvfo_se3_exp_bwd(_p(v), _p(g), v.shape[0], _p(gin))
    return gin


# ---- Caffe-convention layers (parity unpinned: see dvf_oracle.c) ------------------------------
def caffe_geo_fwd(depth, T, K):
    depth, T, K = _f32(depth), _f32(T).reshape(-1, 16), _f32(K).reshape(-1, 4)
    N, H, W = depth.shape
    pts = np.empty((N, 3, H, W), np.float32)
    lib().dvfo_caffe_geo_fwd(_p(depth), _p(T), _p(K), N, H, W, _p(pts))
    return pts


def caffe_geo_bwd(top, depth, T, K):
    top, depth, T, K = _f32(top), _f32(depth), _f32(T).reshape(-1, 16), _f32(K).reshape(-1, 4)
    N, H, W = depth.shape
    dd, dT, dK = np.empty_like(depth), np.empty((N, 16), np.float32), np.empty((N, 4), np.float32)
    lib().dvfo_caffe_geo_bwd(_p(top), _p(depth), _p(T), _p(K), N, H, W, _p(dd), _p(dT), _p(dK))
    return dd, dT, dK


def caffe_pinhole_fwd(pts, K):
    pts, K = _f32(pts), _f32(K).reshape(-1, 4)
    N, _, H, W = pts.shape
    out = np.empty((N, 2, H, W), np.float32)
    lib().dvfo_caffe_pinhole_fwd(_p(pts), _p(K), N, H, W, _p(out))
    return out


def caffe_pinhole_bwd(cdiff, pts, K):
    cdiff, pts, K = _f32(cdiff), _f32(pts), _f32(K).reshape(-1, 4)
    N, _, H, W = pts.shape
    dp, dK = np.empty_like(pts), np.empty((N, 4), np.float32)
    lib().dvfo_caffe_pinhole_bwd(_p(cdiff), _p(pts), _p(K), N, H, W, _p(dp), _p(dK))
    return dp, dK


def caffe_warp_fwd(img, xy):
    img, xy = _f32(img), _f32(xy)
    N, Cc, H, W = img.shape
    out = np.empty_like(img)
    lib().dvfo_caffe_warp_fwd(_p(img), _p(xy), N, Cc, H, W, _p(out))
    return out


def caffe_warp_bwd(top, img, xy, need_gimg=True):
    top, img, xy = _f32(top), _f32(img), _f32(xy)
    N, Cc, H, W = img.shape
    gi = np.empty_like(img) if need_gimg else None
    gxy = np.empty_like(xy)
    lib().dvfo_caffe_warp_bwd(_p(top), _p(img), _p(xy), N, Cc, H, W, _p(gi), _p(gxy))
    return gi, gxy


def caffe_abs_loss(a, b, weight=1.0):
    a, b = _f32(a), _f32(b)
    ga, gb = np.empty_like(a), np.empty_like(b)
    f = lib().dvfo_caffe_abs_loss
    f.restype = C.c_double
    val = f(_p(a), _p(b), C.c_size_t(a.size), a.shape[0], C.c_float(weight), _p(ga), _p(gb))
    return val, ga, gb


def caffe_edge_smooth(img, inv_depth, weight=10.0, need_grad=True):
    """(loss[2] float64, ginv) of the Caffe graphs' edge-aware smoothness (PARITY UNPINNED, see dvf_oracle.c)."""
    img, inv_depth = _f32(img), _f32(inv_depth)
    N, _, H, W = img.shape
    loss = np.zeros(2, np.float64)
    g = np.empty((N, 1, H, W), np.float32) if need_grad else None
    lib().dvfo_caffe_edge_smooth(_p(img), _p(inv_depth), N, H, W, C.c_float(weight), loss.ctypes.data_as(C.c_void_p), _p(g))
    return loss, g


def ssim_loss(x, y, valid=None, need_grad=True):
    """(loss, gy) of the SSIM term defined in csrc/dvf_ssim.cu (new functionality, PARITY UNPINNED)."""
    x, y = _f32(x), _f32(y)
    B, Cc, H, W = x.shape
    v = None if valid is None else np.ascontiguousarray(valid, np.uint8)
    g = np.empty_like(y) if need_grad else None
    loss = lib().dvfo_ssim_loss(_p(x), _p(y), None if v is None else v.ctypes.data_as(C.c_void_p), B, Cc, H, W, _p(g))
    return float(loss), g
