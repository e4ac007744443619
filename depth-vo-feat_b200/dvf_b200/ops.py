"""torch.autograd bindings of the CUDA operators in libdvf_b200.so.

PyTorch is used for device memory, streams and autograd plumbing only: every tensor handed to
the library is a raw data_ptr(), every launch goes to torch's current CUDA stream.  There is no
CPU path -- tensors that are not on a CUDA device raise DvfError.

Reference functions replaced (pytorch_version/): inverse_warp.py:26-193 and the
photometric_reconstruction_loss variants of loss_functions.py:7-20, loss_functions_sfm.py:9-46,
loss_function_sfm_old.py:7-46.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import DvfError, PADDING, ROTATION, dvf_desc, dvf_level, dvf_loss_desc, dvf_pose_args, dvf_reg_level

_WS = {}


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device.  The raw getter costs ~0.3 us; going through
    torch.cuda.current_stream() costs 5-15 us and this is called for every launch (it was 120 us of a 390 us host-bound
    loss call, profiles/host_profile.py)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req(t, name, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise DvfError(f"{name} is on {t.device}: dvf_b200 is CUDA-only (sm_100a); there is no CPU fallback")
    if t.dtype != torch.float32:
        raise DvfError(f"{name}: dtype {t.dtype} not supported by this entry (float32 expected)")
    if ndim is not None and t.dim() != ndim:
        raise AssertionError(f"wrong size for {name}, expected {ndim} dims, got {list(t.size())}")
    return t.contiguous()


def _nhwc_ok(t: torch.Tensor) -> bool:
    """True if `t` ([B,C,H,W] view) is physically channels-last and its channel count suits the NHWC kernel
    (16 bytes of channels per lane, a power-of-two number of lanes <= 32 per pixel)."""
    if t.dim() != 4 or t.dtype not in (torch.float32, torch.bfloat16):
        return False
    if t.is_contiguous() or not t.is_contiguous(memory_format=torch.channels_last):
        return False
    vec = 8 if t.dtype == torch.bfloat16 else 4
    C_ = t.shape[1]
    lpp = C_ // vec
    return C_ % vec == 0 and 1 <= lpp <= 32 and (lpp & (lpp - 1)) == 0 and t.data_ptr() % 16 == 0


# Feature maps in the reference layout (dense NCHW, what the reference's FeatExtractor hands over, unsupervise.py:104-109)
# are re-laid out channels-last on the way in when their channel count suits the NHWC kernel: one transpose pass per map
# buys a 4-7x faster loss kernel (C4 shape: 1.84 ms generic NCHW vs 0.26 ms channels-last).  Their gradients come back
# as channels-last tensors of the same logical shape.  False: keep NCHW maps on the generic kernel.
NCHW_FEATURES_VIA_NHWC = True


def to_channels_last(t: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] fp32 / bf16 -> the same tensor in channels-last memory (tiled transpose kernel for dense NCHW input;
    not differentiable: callers take gradients w.r.t. the tensor they passed in)."""
    if t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous():
        return t
    src = t.detach()
    if not src.is_contiguous():
        return src.contiguous(memory_format=torch.channels_last)
    B, Cc, H, W = src.shape
    out = torch.empty_like(src, memory_format=torch.channels_last)
    _lib.check(_lib.load().dvf_transpose_planes(src.data_ptr(), out.data_ptr(), B, Cc, H * W, src.element_size(), _stream()),
               "dvf_transpose_planes")
    return out


def _nhwc_candidate(t: torch.Tensor) -> bool:
    if t.dim() != 4 or t.dtype not in (torch.float32, torch.bfloat16):
        return False
    vec = 8 if t.dtype == torch.bfloat16 else 4
    C_ = t.shape[1]
    lpp = C_ // vec
    return C_ >= 8 and C_ % vec == 0 and 1 <= lpp <= 32 and (lpp & (lpp - 1)) == 0


def _req_maps(maps, name):
    """Image / feature tensors of one loss call -> (tensors, layout, dtype).  Channels-last fp32 / bf16 maps go to
    the NHWC kernel as they are, dense NCHW feature maps after a re-layout (see NCHW_FEATURES_VIA_NHWC); anything
    else is brought to dense NCHW fp32 (the reference layout)."""
    for t in maps:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
        if not t.is_cuda:
            raise DvfError(f"{name} is on {t.device}: dvf_b200 is CUDA-only (sm_100a); there is no CPU fallback")
        if t.dim() != 4:
            raise AssertionError(f"wrong size for {name}, expected 4 dims, got {list(t.size())}")
    if all(_nhwc_ok(t) for t in maps) and len({t.dtype for t in maps}) == 1:
        return list(maps), _lib.NHWC, (_lib.BF16 if maps[0].dtype == torch.bfloat16 else _lib.F32)
    if NCHW_FEATURES_VIA_NHWC and len({t.dtype for t in maps}) == 1 and all(_nhwc_candidate(t) for t in maps):
        conv = [t if _nhwc_ok(t) else to_channels_last(t) for t in maps]
        if all(_nhwc_ok(t) for t in conv):
            return conv, _lib.NHWC, (_lib.BF16 if conv[0].dtype == torch.bfloat16 else _lib.F32)
    out = []
    for t in maps:
        if t.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            raise DvfError(f"{name}: dtype {t.dtype} not supported")
        out.append(t.float().contiguous())
    return out, _lib.NCHW, _lib.F32


def workspace(nbytes: int, device: torch.device, signature) -> torch.Tensor:
    """Zero-initialised scratch, one per (device, stream, size, call signature).  The kernels restore the
    ticket counters they use, so a workspace is reusable by later calls WITH THE SAME PLAN on the same stream
    (launches on one stream are serialised); anything that changes the plan -- shapes, layout, element type,
    view count, tuning fields -- must be part of `signature`, because another plan lays the counters out
    elsewhere.  `drop_workspace` forgets a buffer whose launch failed (its counters may be dirty)."""
    key = (device.index, _stream(), int(nbytes), signature)
    ws = _WS.get(key)
    if ws is None:
        if len(_WS) > 64:
            _WS.clear()     # stream-ordered allocator: in-flight kernels keep their memory until they finish
        ws = torch.zeros(int(nbytes), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def drop_workspace(ws: torch.Tensor):
    for k in [k for k, v in _WS.items() if v is ws]:
        del _WS[k]


def _checked(status: int, what: str, ws: Optional[torch.Tensor] = None):
    if status != 0 and ws is not None:
        drop_workspace(ws)
    _lib.check(status, what)


def _same_device(*tensors):
    """All tensors of a call live on one CUDA device; returns a context that makes it current (the library
    launches on the current device's stream)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise DvfError(f"tensors of one call are on different devices ({dev} and {t.device})")
    return torch.cuda.device(dev)


# ------------------------------------------------------------------------------------------------
# pose -> projection
# ------------------------------------------------------------------------------------------------
# Arithmetic profile of the pose chain (include/dvf_b200.h: DVF_ROT_REF_CUDA): "ref_cpu" reproduces the reference run on
# torch-CPU bit for bit (the default, and what the golden vectors pin), "ref_cuda" the reference run with torch-CUDA eager
# (libdevice sin / cos, FMA-chain tiny matmuls).  With "ref_cuda" the loss runs in its three-launch form (projection
# matrices from dvf_pose_proj_fwd, loss kernel given P, dvf_pose_proj_bwd), because the single-launch entry evaluates the
# torch-CPU profile in its prologue.
ARITHMETIC = "ref_cpu"
ROT_REF_CUDA = 0x100


def _rot(rotation_mode):
    if ARITHMETIC not in ("ref_cpu", "ref_cuda"):
        raise DvfError(f"ops.ARITHMETIC = {ARITHMETIC!r}: expected 'ref_cpu' or 'ref_cuda'")
    return ROTATION[rotation_mode] | (ROT_REF_CUDA if ARITHMETIC == "ref_cuda" else 0)


def pose_proj_fwd(vec, K, Kinv, V, rotation_mode, downscales: Sequence[float], want_posemat=False):
    """vec [B*V,6] (b-major) -> (posemat [B*V,3,4] | None, P [L,B*V,3,4] | None, Kinv_s [L,B,3,3] | None)."""
    lib = _lib.load()
    n = vec.shape[0]
    B = n // V
    L = len(downscales)
    dev = vec.device
    posemat = torch.empty(n, 3, 4, device=dev, dtype=torch.float32) if (want_posemat or K is None) else None
    P = torch.empty(L, n, 3, 4, device=dev, dtype=torch.float32) if (K is not None and L) else None
    Ks = torch.empty(L, B, 3, 3, device=dev, dtype=torch.float32) if (Kinv is not None and L) else None
    ds = (C.c_float * max(L, 1))(*[float(d) for d in downscales])
    with _same_device(vec, K, Kinv):
        _lib.check(lib.dvf_pose_proj_fwd(_ptr(vec), _ptr(K), _ptr(Kinv), B, V, _rot(rotation_mode), ds, L,
                                         _ptr(posemat), _ptr(P), _ptr(Ks), _stream()), "dvf_pose_proj_fwd")
    return posemat, P, Ks


def pose_proj_bwd(gP, gposemat, vec, K, V, rotation_mode, downscales: Sequence[float]):
    lib = _lib.load()
    n = vec.shape[0]
    L = len(downscales)
    gvec = torch.empty(n, 6, device=vec.device, dtype=torch.float32)
    ds = (C.c_float * max(L, 1))(*[float(d) for d in downscales])
    with _same_device(gP, gposemat, vec, K):
        _lib.check(lib.dvf_pose_proj_bwd(_ptr(gP), _ptr(gposemat), _ptr(vec), _ptr(K), n // V, V, ROTATION[rotation_mode],
                                         ds, L, _ptr(gvec), _stream()), "dvf_pose_proj_bwd")
    return gvec


class PoseVec2Mat(torch.autograd.Function):
    """pose_vec2mat (inverse_warp.py:141-157) with analytic backward."""

    @staticmethod
    def forward(ctx, vec, rotation_mode):
        vec = _req(vec, "vec", 2)
        posemat, _, _ = pose_proj_fwd(vec, None, None, 1, rotation_mode, [], want_posemat=True)
        ctx.save_for_backward(vec)
        ctx.rotation_mode = rotation_mode
        return posemat

    @staticmethod
    def backward(ctx, g):
        (vec,) = ctx.saved_tensors
        return pose_proj_bwd(None, _req(g, "grad"), vec, None, 1, ctx.rotation_mode, []), None


# ------------------------------------------------------------------------------------------------
# inverse_warp
# ------------------------------------------------------------------------------------------------
# F.grid_sample convention of every sampling entry.  False = torch >= 1.3's default, which is what the reference runs
# with today (inverse_warp.py:191 passes nothing); True = the torch <= 1.2 behaviour the reference was written for.
ALIGN_CORNERS = False


def _flags(align_corners=None):
    ac = ALIGN_CORNERS if align_corners is None else align_corners
    return _lib.FLAG_ALIGN_CORNERS if ac else 0


def _desc(img, padding_mode, align_corners=None):
    B, Cc, H, W = img.shape
    return dvf_desc(B, Cc, H, W, _lib.F32, _lib.NCHW, PADDING[padding_mode], _flags(align_corners))


def inverse_warp_fwd_P(img, depth, P, Kinv, padding_mode="zeros", want_valid=False):
    """Non-autograd forward given P = K @ pose_vec2mat(pose) [B,3,4]."""
    lib = _lib.load()
    img, depth, P, Kinv = _req(img, "img", 4), _req(depth, "depth", 3), _req(P, "P", 3), _req(Kinv, "intrinsics_inv", 3)
    warped = torch.empty_like(img)
    valid = torch.empty(depth.shape, dtype=torch.uint8, device=img.device) if want_valid else None
    d = _desc(img, padding_mode)
    if ARITHMETIC == "ref_cuda":
        d.flags |= _lib.FLAG_REF_CUDA   # forward only: the backward entry follows torch-CPU
    with _same_device(img, depth, P, Kinv):
        _lib.check(lib.dvf_inverse_warp_fwd(C.byref(d), _ptr(img), _ptr(depth), _ptr(P), _ptr(Kinv), _ptr(warped),
                                            _ptr(valid), _stream()), "dvf_inverse_warp_fwd")
    return (warped, valid) if want_valid else warped


def inverse_warp_bwd_P(gout, img, depth, P, Kinv, padding_mode="zeros", need_gimg=True):
    """Non-autograd backward: returns (gimg | None, gdepth, gP[B,3,4])."""
    lib = _lib.load()
    gout, img, depth, P, Kinv = (_req(gout, "grad_output", 4), _req(img, "img", 4), _req(depth, "depth", 3),
                                 _req(P, "P", 3), _req(Kinv, "intrinsics_inv", 3))
    d = _desc(img, padding_mode)
    gdepth = torch.empty_like(depth)
    gP = torch.empty(img.shape[0], 3, 4, device=img.device, dtype=torch.float32)
    gimg = torch.zeros_like(img) if need_gimg else None
    nbytes = lib.dvf_inverse_warp_bwd_workspace_bytes(C.byref(d))
    with _same_device(gout, img, depth, P, Kinv):
        ws = workspace(nbytes, img.device, ('warp_bwd',) + tuple(img.shape))
        _checked(lib.dvf_inverse_warp_bwd(C.byref(d), _ptr(gout), _ptr(img), _ptr(depth), _ptr(P), _ptr(Kinv),
                                          _ptr(gdepth), _ptr(gP), _ptr(gimg), _ptr(ws), ws.numel(), _stream()),
                 "dvf_inverse_warp_bwd", ws)
    return gimg, gdepth, gP


class InverseWarp(torch.autograd.Function):
    """inverse_warp (inverse_warp.py:160-193): differentiable w.r.t. img, depth and pose."""

    @staticmethod
    def forward(ctx, img, depth, pose, intrinsics, intrinsics_inv, rotation_mode, padding_mode):
        img, depth, pose = _req(img, "img", 4), _req(depth, "depth", 3), _req(pose, "pose", 2)
        K, Kinv = _req(intrinsics, "intrinsics", 3), _req(intrinsics_inv, "intrinsics_inv", 3)
        if ctx.needs_input_grad[3] or ctx.needs_input_grad[4]:
            raise DvfError("gradients w.r.t. the camera intrinsics are not implemented (unused by the reference)")
        _, P, _ = pose_proj_fwd(pose, K, None, 1, rotation_mode, [1.0])
        P = P[0]
        warped = inverse_warp_fwd_P(img, depth, P, Kinv, padding_mode)
        ctx.save_for_backward(img, depth, pose, K, Kinv, P)
        ctx.cfg = (rotation_mode, padding_mode)
        return warped

    @staticmethod
    def backward(ctx, gout):
        img, depth, pose, K, Kinv, P = ctx.saved_tensors
        rotation_mode, padding_mode = ctx.cfg
        gimg, gdepth, gP = inverse_warp_bwd_P(gout, img, depth, P, Kinv, padding_mode, need_gimg=ctx.needs_input_grad[0])
        gpose = None
        if ctx.needs_input_grad[2]:
            gpose = pose_proj_bwd(gP.unsqueeze(0), None, pose, K, 1, rotation_mode, [1.0])
        return gimg, (gdepth if ctx.needs_input_grad[1] else None), gpose, None, None, None, None


class Pixel2Cam(torch.autograd.Function):
    """pixel2cam (inverse_warp.py:26-40): depth [B,H,W], K^-1 [B,3,3] -> camera-frame points [B,3,H,W]; differentiable
    w.r.t. depth (the intrinsics are data in the reference)."""

    @staticmethod
    def forward(ctx, depth, intrinsics_inv):
        lib = _lib.load()
        depth, Kinv = _req(depth, "depth", 3), _req(intrinsics_inv, "intrinsics_inv", 3)
        if ctx.needs_input_grad[1]:
            raise DvfError("gradients w.r.t. the camera intrinsics are not implemented (unused by the reference)")
        B, H, W = depth.shape
        cam = torch.empty(B, 3, H, W, device=depth.device, dtype=torch.float32)
        with _same_device(depth, Kinv):
            _lib.check(lib.dvf_pixel2cam(_ptr(depth), _ptr(Kinv), B, H, W, _ptr(cam), _stream()), "dvf_pixel2cam")
        ctx.save_for_backward(Kinv)
        return cam

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gcam):
        lib = _lib.load()
        (Kinv,) = ctx.saved_tensors
        g = _req(gcam, "grad_cam", 4)
        B, _, H, W = g.shape
        gdepth = torch.empty(B, H, W, device=g.device, dtype=torch.float32)
        with _same_device(g, Kinv):
            _lib.check(lib.dvf_pixel2cam_bwd(_ptr(g), _ptr(Kinv), B, H, W, _ptr(gdepth), _stream()), "dvf_pixel2cam_bwd")
        return gdepth, None


def pixel2cam(depth, intrinsics_inv):
    """pixel2cam (inverse_warp.py:26-40)."""
    return Pixel2Cam.apply(depth, intrinsics_inv)


class Cam2Pixel(torch.autograd.Function):
    """cam2pixel (inverse_warp.py:43-74): [B,3,H,W] points, optional rotation [B,3,3] and translation [B,3,1] -> sampling
    grid [B,H,W,2]; differentiable w.r.t. all three."""

    @staticmethod
    def forward(ctx, cam_coords, proj_c2p_rot, proj_c2p_tr, padding_mode):
        lib = _lib.load()
        cam = _req(cam_coords, "cam_coords", 4)
        rot = None if proj_c2p_rot is None else _req(proj_c2p_rot, "proj_c2p_rot", 3)
        tr = None if proj_c2p_tr is None else _req(proj_c2p_tr, "proj_c2p_tr").reshape(cam.shape[0], 3).contiguous()
        B, _, H, W = cam.shape
        grid = torch.empty(B, H, W, 2, device=cam.device, dtype=torch.float32)
        with _same_device(cam, rot, tr):
            _lib.check(lib.dvf_cam2pixel(_ptr(cam), _ptr(rot), _ptr(tr), B, H, W, PADDING[padding_mode], _ptr(grid), _stream()),
                       "dvf_cam2pixel")
        ctx.save_for_backward(cam, *[t for t in (rot, tr) if t is not None])
        ctx.meta = (rot is not None, tr is not None, padding_mode, None if proj_c2p_tr is None else tuple(proj_c2p_tr.shape))
        return grid

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, ggrid):
        lib = _lib.load()
        has_rot, has_tr, padding_mode, tr_shape = ctx.meta
        saved = list(ctx.saved_tensors)
        cam = saved.pop(0)
        rot = saved.pop(0) if has_rot else None
        tr = saved.pop(0) if has_tr else None
        g = _req(ggrid, "grad_grid", 4)
        B, _, H, W = cam.shape
        need_cam, need_rot, need_tr = ctx.needs_input_grad[0], has_rot and ctx.needs_input_grad[1], has_tr and ctx.needs_input_grad[2]
        gcam = torch.empty_like(cam) if need_cam else None
        grot = torch.empty(B, 3, 3, device=cam.device, dtype=torch.float32) if need_rot else None
        gtr = torch.empty(B, 3, device=cam.device, dtype=torch.float32) if need_tr else None
        with _same_device(g, cam, rot, tr):
            _lib.check(lib.dvf_cam2pixel_bwd(_ptr(g), _ptr(cam), _ptr(rot), _ptr(tr), B, H, W, PADDING[padding_mode], _ptr(gcam),
                                             _ptr(grot), _ptr(gtr), _stream()), "dvf_cam2pixel_bwd")
        return gcam, grot, (None if gtr is None else gtr.reshape(tr_shape)), None


def cam2pixel(cam_coords, proj_c2p_rot, proj_c2p_tr, padding_mode):
    """cam2pixel (inverse_warp.py:43-74)."""
    return Cam2Pixel.apply(cam_coords, proj_c2p_rot, proj_c2p_tr, padding_mode)


# ------------------------------------------------------------------------------------------------
# area pyramid
# ------------------------------------------------------------------------------------------------
def area_downsample(img, size):
    """F.interpolate(img, size, mode='area') (loss_functions_sfm.py:18-19), no autograd."""
    lib = _lib.load()
    img = _req(img, "img", 4)
    B, Cc, H, W = img.shape
    h, w = int(size[0]), int(size[1])
    if (h, w) == (H, W):
        return img
    out = torch.empty(B, Cc, h, w, device=img.device, dtype=torch.float32)
    _lib.check(lib.dvf_area_downsample(_ptr(img), B * Cc, H, W, h, w, _ptr(out), _stream()), "dvf_area_downsample")
    return out


def area_pyramid(img, sizes: Sequence[Sequence[int]]):
    """All requested levels of the 'area' pyramid of img; a single pass when they are /2,/4,/8."""
    lib = _lib.load()
    img = _req(img, "img", 4)
    if img.requires_grad and torch.is_grad_enabled() and any(tuple(s) != tuple(img.shape[2:]) for s in sizes):
        raise DvfError("area down-sampling of a tensor that requires grad is not implemented "
                       "(the reference only down-samples input images)")
    B, Cc, H, W = img.shape
    sizes = [(int(s[0]), int(s[1])) for s in sizes]
    small = [s for s in sizes if s != (H, W)]
    uniq = sorted(set(small), reverse=True)
    fast = (len(uniq) > 0 and uniq == [(H >> (k + 1), W >> (k + 1)) for k in range(len(uniq))]
            and H % (1 << len(uniq)) == 0 and W % (1 << len(uniq)) == 0 and len(uniq) <= 3)
    made = {(H, W): img}
    if fast:
        outs = [torch.empty(B, Cc, s[0], s[1], device=img.device, dtype=torch.float32) for s in uniq]
        arr = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        _lib.check(lib.dvf_area_pyramid(_ptr(img), B * Cc, H, W, len(outs), arr, _stream()), "dvf_area_pyramid")
        made.update(dict(zip(uniq, outs)))
    else:
        for s in uniq:
            made[s] = area_downsample(img, s)
    return [made[s] for s in sizes]


# ------------------------------------------------------------------------------------------------
# fused reconstruction loss
# ------------------------------------------------------------------------------------------------
class _LossCfg:
    __slots__ = ("V", "L", "rotation_mode", "padding_mode", "downscales", "has_expl", "expl_channels", "align_corners",
                 "global_batch", "nan_check", "disparity_eps", "img_scale")


def _scaled(tensors: List[Optional[torch.Tensor]], g: torch.Tensor):
    """[t * g] out of place (the unit gradients stay valid for another backward call, retain_graph=True)."""
    idx = [i for i, t in enumerate(tensors) if t is not None]
    out = [None] * len(tensors)
    if idx:
        for i, r in zip(idx, torch._foreach_mul([tensors[i] for i in idx], g)):
            out[i] = r
    return out


_NAN_WORDS = {}


def nan_flags(device=None) -> torch.Tensor:
    """The device word the loss kernels OR NaN bits into when nan_check is on (bit l*V+v <-> terms[l*V+v]); one per
    device.  Reading it (`.item()`) is the only synchronisation, and the caller chooses when."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    w = _NAN_WORDS.get(dev.index)
    if w is None:
        w = torch.zeros(1, dtype=torch.int32, device=dev)
        _NAN_WORDS[dev.index] = w
    return w


class _LossCall:
    """One validated loss call: prepared tensors + which gradients are wanted; can launch the fused kernel
    forward-only or forward+backward (any number of times)."""

    def __init__(self, cfg, pose, K, Kinv, tensors, needs):
        V, L = cfg.V, cfg.L
        self.cfg = cfg
        self.pose = _req(pose, "pose", 3)
        self.K, self.Kinv = _req(K, "intrinsics", 3), _req(Kinv, "intrinsics_inv", 3)
        maps, self.layout, self.dtype = _req_maps(list(tensors[0:L + L * V]), "tgt/src")
        # dense NCHW feature maps that were re-laid out channels-last on the way in (the reference's FeatExtractor output)
        self.relaid = [m is not t and self.layout == _lib.NHWC and t.is_contiguous() for m, t in zip(maps, tensors[0:L + L * V])]
        self.tgts, self.srcs = maps[0:L], maps[L:L + L * V]
        self.in_dtypes = [t.dtype for t in tensors[0:L + L * V]]
        self.depths = [_req(t, "depth", 3) for t in tensors[L + L * V:2 * L + L * V]]
        self.expls = [_req(t, "explainability_mask", 4) for t in tensors[2 * L + L * V:]] if cfg.has_expl else []
        off = 4  # index of the first *tensors entry in needs_input_grad
        self.need_pose = needs[1]
        self.need_tgt = [needs[off + i] for i in range(L)]
        self.need_src = [needs[off + L + i] for i in range(L * V)]
        self.need_depth = [needs[off + L + L * V + i] for i in range(L)]
        self.need_expl = [needs[off + 2 * L + L * V + i] for i in range(L)] if cfg.has_expl else [False] * L
        self.B, self.C = self.tgts[0].shape[0], self.tgts[0].shape[1]
        for l in range(L):
            h, w = self.depths[l].shape[1], self.depths[l].shape[2]
            if self.tgts[l].shape != (self.B, self.C, h, w):
                raise AssertionError(f"level {l}: target {list(self.tgts[l].shape)} does not match depth {list(self.depths[l].shape)}")
            for v in range(V):
                if self.srcs[l * V + v].shape != self.tgts[l].shape:
                    raise AssertionError(f"level {l} view {v}: source {list(self.srcs[l * V + v].shape)} != target "
                                         f"{list(self.tgts[l].shape)}")
            if cfg.has_expl:
                e = self.expls[l]
                if e.shape[0] != self.B or e.shape[1] < V or e.shape[2:] != (h, w):
                    raise AssertionError(f"level {l}: explainability mask {list(e.shape)} does not match depth")
        self.any_map_grad = any(self.need_tgt) or any(self.need_src)
        self.any_grad = self.any_map_grad or self.need_pose or any(self.need_depth) or any(self.need_expl)

    def run(self, want_grads: bool, upstream: Optional[torch.Tensor] = None):
        """-> (terms [L*V], grads | None) with grads = [pose] + tgt(L) + src(L*V) + depth(L) (+ expl(L))."""
        lib = _lib.load()
        cfg, V, L, B, Cc = self.cfg, self.cfg.V, self.cfg.L, self.B, self.C
        dev = self.pose.device
        with _same_device(self.pose, self.K, self.Kinv, *self.tgts, *self.srcs, *self.depths, *self.expls):
            vec = self.pose.reshape(B * V, 6)
            # the kernel derives P and K^-1_s from the pose itself and finishes with the pose backward (one launch);
            # dP is only materialised for the pose gradient, inside the kernel's workspace
            g_pose = torch.empty(B, V, 6, device=dev, dtype=torch.float32) if (want_grads and self.need_pose) else None
            ds_arr = (C.c_float * L)(*[float(x) for x in cfg.downscales])
            pargs = dvf_pose_args(vec.data_ptr(), self.K.data_ptr(), self.Kinv.data_ptr(), ds_arr, ROTATION[cfg.rotation_mode],
                                  0, _ptr(g_pose))
            terms = torch.empty(L * V, device=dev, dtype=torch.float32)
            levels = (dvf_level * L)()
            g_tgt, g_src, g_depth, g_expl = [None] * L, [None] * (L * V), [None] * L, [None] * L
            # channels-last bf16 maps get their target-map gradient in bf16 straight from the kernel (source-map gradients are
            # accumulated in fp32 and converted below)
            bf16_grads = self.layout == _lib.NHWC and self.dtype == _lib.BF16
            gdt = torch.bfloat16 if bf16_grads else torch.float32
            for l in range(L):
                lv = levels[l]
                h, w = self.depths[l].shape[1], self.depths[l].shape[2]
                lv.H, lv.W = h, w
                lv.depth, lv.tgt = self.depths[l].data_ptr(), self.tgts[l].data_ptr()
                for v in range(V):
                    sv = self.srcs[l * V + v]
                    lv.src[v] = sv.data_ptr()
                    if want_grads and self.need_src[l * V + v]:
                        g_src[l * V + v] = torch.empty_like(sv, dtype=torch.float32)   # accumulated in fp32; zero-filled by the entry
                        lv.gsrc[v] = g_src[l * V + v].data_ptr()
                if cfg.has_expl:
                    e = self.expls[l]
                    lv.expl, lv.expl_bstride = e.data_ptr(), e.shape[1] * h * w
                    if want_grads and self.need_expl[l]:
                        # dense [B,V,h,w]; channels >= V of a wider mask get zero gradient
                        g_expl[l] = torch.empty(B, V, h, w, device=dev, dtype=torch.float32)
                        lv.gexpl = g_expl[l].data_ptr()
                if want_grads and self.need_depth[l]:
                    g_depth[l] = torch.empty_like(self.depths[l])
                    lv.gdepth = g_depth[l].data_ptr()
                if want_grads and self.need_tgt[l]:
                    g_tgt[l] = torch.empty_like(self.tgts[l], dtype=gdt)
                    lv.gtgt = g_tgt[l].data_ptr()
            flags = _flags(cfg.align_corners) | _lib.FLAG_ZERO_GSRC | (_lib.FLAG_NAN_CHECK if cfg.nan_check else 0) | \
                (_lib.FLAG_DISPARITY if cfg.disparity_eps is not None else 0)
            up = None
            if upstream is not None:
                up = upstream.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
            d = dvf_loss_desc(B, Cc, V, L, self.dtype, self.layout, PADDING[cfg.padding_mode], flags,
                              int(cfg.global_batch or 0), _lib.BF16 if bf16_grads else _lib.F32, 0, 0, _ptr(up),
                              nan_flags(dev).data_ptr() if cfg.nan_check else None, 0, 0, None,
                              float(cfg.disparity_eps or 0.0), float(cfg.img_scale))
            nbytes = lib.dvf_photo_loss_workspace_bytes(C.byref(d), levels)
            if nbytes == 0:
                raise DvfError("dvf_photo_loss_workspace_bytes rejected the shapes")
            ws = workspace(nbytes, dev, ("loss", B, Cc, V, cfg.has_expl, self.layout, self.dtype) +
                           tuple(tuple(x.shape[1:]) for x in self.depths))
            if ARITHMETIC == "ref_cuda":
                # torch-CUDA's rounding of the pose chain: P / K^-1_s from dvf_pose_proj_fwd, loss kernel given P, pose
                # backward from the dL/dP it leaves behind (three launches)
                _, P, Kinv_s = pose_proj_fwd(vec, self.K, self.Kinv, V, cfg.rotation_mode, cfg.downscales)
                gP = torch.empty(L, B * V, 3, 4, device=dev, dtype=torch.float32) if g_pose is not None else None
                for l in range(L):
                    levels[l].P, levels[l].Kinv = P[l].data_ptr(), Kinv_s[l].data_ptr()
                    if gP is not None:
                        levels[l].gP = gP[l].data_ptr()
                _checked(lib.dvf_photo_loss_fused(C.byref(d), levels, _ptr(terms), _ptr(ws), ws.numel(), _stream()),
                         "dvf_photo_loss_fused", ws)
                if gP is not None:
                    g_pose = pose_proj_bwd(gP, None, vec, self.K, V, cfg.rotation_mode, cfg.downscales).view(B, V, 6)
            else:
                _checked(lib.dvf_photo_loss_fused_pose(C.byref(d), levels, C.byref(pargs), _ptr(terms), _ptr(ws), ws.numel(),
                                                       _stream()), "dvf_photo_loss_fused_pose", ws)
            if not want_grads:
                return terms, None
            for l in range(L):
                if cfg.has_expl and self.need_expl[l] and self.expls[l].shape[1] > V:
                    full = torch.zeros_like(self.expls[l])
                    full[:, :V] = g_expl[l]
                    g_expl[l] = full
            # gradients of the maps go back in the dtype the caller handed in
            g_tgt = [g if g is None or g.dtype == self.in_dtypes[i] else g.to(self.in_dtypes[i]) for i, g in enumerate(g_tgt)]
            g_src = [g if g is None or g.dtype == self.in_dtypes[L + i] else g.to(self.in_dtypes[L + i])
                     for i, g in enumerate(g_src)]
            return terms, [g_pose] + g_tgt + g_src + g_depth + (g_expl if cfg.has_expl else [])


def _nhwc_to_nchw_scaled(g: Optional[torch.Tensor], scale: torch.Tensor):
    """channels-last fp32 gradient map -> dense NCHW, multiplied by the device scalar `scale` in the same pass"""
    if g is None:
        return None
    B, Cc, H, W = g.shape
    out = torch.empty(B, Cc, H, W, device=g.device, dtype=torch.float32)
    sc = scale.detach().to(device=g.device, dtype=torch.float32).reshape(1).contiguous()
    with _same_device(g):
        _lib.check(_lib.load().dvf_transpose_planes_scaled(g.data_ptr(), out.data_ptr(), B, H * W, Cc, sc.data_ptr(), _stream()),
                   "dvf_transpose_planes_scaled")
    return out


class FusedPhotoLoss(torch.autograd.Function):
    """sum over levels and views of mean|(tgt - warp(src_v)) * valid_v [* expl_v]| (dvf_photo_loss_fused_pose).

    The kernel produces the loss and every gradient in ONE pass.  Two schedules, chosen per call:
      * no gradients to the maps (image losses): the single pass runs in forward() for upstream gradient 1; backward()
        multiplies the (small) depth / pose / mask gradients by the upstream scalar, out of place, so it can be
        called again (retain_graph=True);
      * gradients to the maps (feature losses): forward() runs the kernel forward-only, backward() runs the fused pass
        with the upstream scalar handed over as a device pointer -- scaling three full feature-map gradients after
        the fact costs more HBM traffic than re-reading the inputs (C4 shape: 110 us of scaling vs a ~60 us forward);
      * dense NCHW fp32 feature maps (what the reference's FeatExtractor hands over): they are re-laid out channels-last on
        the way in and their gradients transposed back on the way out; that pass multiplies by the upstream scalar, so
        the single fused pass runs in forward().

    inputs: cfg, pose [B,V,6], K, Kinv, then L target levels, L*V source levels (level-major),
            L depth levels [B,h,w], and L explainability levels [B,>=V,h,w] if cfg.has_expl.
    outputs: (loss scalar, terms [L*V] -- not differentiable, for logging)
    """

    @staticmethod
    def forward(ctx, cfg, pose, K, Kinv, *tensors):
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            raise DvfError("gradients w.r.t. the camera intrinsics are not implemented (unused by the reference)")
        call = _LossCall(cfg, pose, K, Kinv, tensors, ctx.needs_input_grad)
        ctx.call, ctx.unit_grads, ctx.relaid = None, None, None
        if call.any_grad and call.any_map_grad and all(call.relaid) and call.dtype == _lib.F32:
            # dense NCHW fp32 feature maps: their gradients have to be transposed back anyway, and that pass applies the
            # upstream scalar for free -- one fused launch here, no forward-only pass
            terms, ctx.unit_grads = call.run(True)
            ctx.relaid = call.relaid
        elif call.any_grad and not call.any_map_grad:
            terms, ctx.unit_grads = call.run(True)
        else:
            terms, _ = call.run(False)
            if call.any_grad:
                ctx.call = call
        loss = terms.sum()
        ctx.mark_non_differentiable(terms)
        return loss, terms

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        if ctx.call is not None:
            _, grads = ctx.call.run(True, upstream=g_loss)
        elif ctx.relaid is not None:
            n_maps = len(ctx.relaid)
            maps = [_nhwc_to_nchw_scaled(g, g_loss) for g in ctx.unit_grads[1:1 + n_maps]]
            rest = _scaled([ctx.unit_grads[0]] + list(ctx.unit_grads[1 + n_maps:]), g_loss)
            grads = [rest[0]] + maps + rest[1:]
        else:
            grads = _scaled(ctx.unit_grads, g_loss)
        return (None, grads[0], None, None) + tuple(grads[1:])


def fused_photo_loss(tgt_levels, src_levels, depth_levels, pose, K, Kinv, expl_levels=None, downscales=None,
                     rotation_mode="euler", padding_mode="zeros", align_corners=None, global_batch=None, nan_check=False,
                     disparity_eps=None, img_scale=1.0):
    """tgt_levels: L tensors [B,C,h,w]; src_levels: L lists of V tensors; depth_levels: L tensors [B,h,w];
    pose [B,V,6]; expl_levels: None or L tensors [B,>=V,h,w].  Returns (loss, terms[L*V]).
    global_batch: size of the whole (sharded) batch when this call holds only B of its images (dvf_b200.dist);
    nan_check: collect NaN terms in ops.nan_flags() (the reference asserts per view and scale, loss_functions_sfm.py:34)."""
    L = len(depth_levels)
    V = pose.shape[1]
    if V > _lib.DVF_MAX_VIEWS or L > _lib.DVF_MAX_LEVELS:
        raise DvfError(f"at most {_lib.DVF_MAX_VIEWS} views and {_lib.DVF_MAX_LEVELS} levels per call")
    cfg = _LossCfg()
    cfg.V, cfg.L = V, L
    cfg.rotation_mode, cfg.padding_mode = rotation_mode, padding_mode
    cfg.downscales = [1.0] * L if downscales is None else [float(x) for x in downscales]
    cfg.has_expl = expl_levels is not None
    cfg.align_corners, cfg.global_batch, cfg.nan_check = align_corners, global_batch, bool(nan_check)
    cfg.disparity_eps, cfg.img_scale = disparity_eps, float(img_scale)
    flat = list(tgt_levels) + [s for lvl in src_levels for s in lvl] + list(depth_levels)
    if cfg.has_expl:
        flat += list(expl_levels)
    return FusedPhotoLoss.apply(cfg, pose, K, Kinv, *flat)


# ------------------------------------------------------------------------------------------------
# regularisers: smooth_loss / explainability_loss (all scales in one launch, value + gradient)
# ------------------------------------------------------------------------------------------------
class RegLoss(torch.autograd.Function):
    """kind = 'smooth' (loss_functions.py:23-41) or 'explainability' (loss_functions_sfm.py:49-56).
    weights[l] multiplies the l-th map's term; maps are [B,C,h,w] (or [B,h,w]) fp32 CUDA tensors."""

    @staticmethod
    def forward(ctx, kind, weights, *maps):
        lib = _lib.load()
        xs = [_req(m, "map") for m in maps]
        L = len(xs)
        if L == 0 or L > _lib.DVF_MAX_LEVELS:
            raise DvfError(f"1..{_lib.DVF_MAX_LEVELS} maps per call")
        dev = xs[0].device
        levels = (dvf_reg_level * L)()
        grads = []
        for l, x in enumerate(xs):
            if x.dim() < 2:
                raise AssertionError("maps must have at least 2 dimensions")
            h, w = x.shape[-2], x.shape[-1]
            n = x.numel() // (h * w)
            g = torch.empty_like(x) if ctx.needs_input_grad[2 + l] else None
            grads.append(g)
            levels[l] = dvf_reg_level(x.data_ptr(), _ptr(g), n, h, w, float(weights[l]))
        out = torch.empty(1, device=dev, dtype=torch.float32)
        nbytes = lib.dvf_reg_workspace_bytes(levels, L)
        ws = workspace(nbytes, dev, ("reg", kind) + tuple(tuple(x.shape) for x in xs))
        fn = lib.dvf_smooth_loss if kind == "smooth" else lib.dvf_explainability_loss
        _lib.check(fn(levels, L, _ptr(out), _ptr(ws), ws.numel(), _stream()), "dvf_" + kind + "_loss")
        ctx.unit_grads = grads
        return out[0]

    @staticmethod
    def backward(ctx, g_out):
        return (None, None) + tuple(_scaled(ctx.unit_grads, g_out))


def smooth_loss(maps, scale_factor=1):
    maps = list(maps) if type(maps) in (tuple, list) else [maps]
    weights, w = [], 1.0
    for _ in maps:
        weights.append(w)
        w /= scale_factor
    return RegLoss.apply("smooth", weights, *maps)


def explainability_loss(masks):
    masks = list(masks) if type(masks) in (tuple, list) else [masks]
    return RegLoss.apply("explainability", [1.0] * len(masks), *masks)


# ------------------------------------------------------------------------------------------------
# se(3) -> SE(3) exponential map
# ------------------------------------------------------------------------------------------------
class SE3Exp(torch.autograd.Function):
    """SE3_Generator_KITTI (se3_generate.py:7-103): [B,6,1,1] (w,u) -> [B,1,4,4] float64, on the device."""

    @staticmethod
    def forward(ctx, input):
        lib = _lib.load()
        x = _req(input, "input")
        B = x.shape[0]
        flat = x.reshape(B, 6)
        out = torch.empty(B, 1, 4, 4, device=x.device, dtype=torch.float64)
        _lib.check(lib.dvf_se3_exp_fwd(_ptr(flat), B, _ptr(out), _stream()), "dvf_se3_exp_fwd")
        ctx.save_for_backward(flat)
        ctx.in_shape = tuple(input.shape)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        lib = _lib.load()
        (flat,) = ctx.saved_tensors
        B = flat.shape[0]
        g = grad_output.to(torch.float64).contiguous()
        gin = torch.empty(B, 6, device=flat.device, dtype=torch.float32)
        _lib.check(lib.dvf_se3_exp_bwd(_ptr(flat), _ptr(g), B, _ptr(gin), _stream()), "dvf_se3_exp_bwd")
        return gin.reshape(ctx.in_shape)


# ---- Caffe-convention layers (csrc/dvf_caffe.cu) --------------------------------------------------------
def _k4(cam_intrinsic, N):
    k = _req(cam_intrinsic, "cam_intrinsic")
    if k.numel() != N * 4:
        raise AssertionError(f"wrong size for cam_intrinsic, expected [{N},4,1,1] (fx,fy,cx,cy), got {list(k.size())}")
    return k.reshape(N, 4)


class GeoTransform(torch.autograd.Function):
    """GeoTransformLayer (caffe/src/caffe/layers/geometry_transformation.cu:10-176; geo_transform.py:6-38):
    depthmap [N,1,H,W], pose [N,1,4,4] (fp32 or the fp64 SE3Exp output), cam_intrinsic [N,4,1,1] -> points [N,3,H,W]."""

    @staticmethod
    def forward(ctx, depthmap, pose, cam_intrinsic):
        lib = _lib.load()
        d = _req(depthmap, "depthmap", 4)
        N, one, H, W = d.shape
        if one != 1:
            raise AssertionError(f"wrong size for depthmap, expected [N,1,H,W], got {list(d.size())}")
        if not isinstance(pose, torch.Tensor) or pose.numel() != N * 16:
            raise AssertionError(f"wrong size for pose, expected [{N},1,4,4]")
        T = _req(pose.to(torch.float32), "pose").reshape(N, 16)
        k = _k4(cam_intrinsic, N)
        pts = torch.empty(N, 3, H, W, device=d.device, dtype=torch.float32)
        _lib.check(lib.dvf_caffe_geo_fwd(_ptr(d), _ptr(T), _ptr(k), N, H, W, _ptr(pts), _stream()), "dvf_caffe_geo_fwd")
        ctx.save_for_backward(d, T, k)
        ctx.meta = (tuple(pose.shape), pose.dtype, tuple(cam_intrinsic.shape))
        return pts

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gpts):
        lib = _lib.load()
        d, T, k = ctx.saved_tensors
        N, _, H, W = d.shape
        pose_shape, pose_dtype, k_shape = ctx.meta
        g = _req(gpts, "grad_points", 4)
        nd, nT, nK = ctx.needs_input_grad
        gd = torch.empty_like(d) if nd else None
        gT = torch.empty(N, 16, device=d.device, dtype=torch.float32) if nT else None
        gK = torch.empty(N, 4, device=d.device, dtype=torch.float32) if nK else None
        _lib.check(lib.dvf_caffe_geo_bwd(_ptr(g), _ptr(d), _ptr(T), _ptr(k), N, H, W, _ptr(gd), _ptr(gT), _ptr(gK), _stream()),
                   "dvf_caffe_geo_bwd")
        return gd, None if gT is None else gT.reshape(pose_shape).to(pose_dtype), None if gK is None else gK.reshape(k_shape)


class PinHoleProject(torch.autograd.Function):
    """PinHoleLayer (pin_hole_layer.cu:10-146; geo_transform.py:40-59): points [N,3,H,W], cam_intrinsic [N,4,1,1] ->
    proj_coords [N,2,H,W] in pixels (the layer's `flows` top is silenced in every reference prototxt)."""

    @staticmethod
    def forward(ctx, transformed_points, cam_intrinsic):
        lib = _lib.load()
        p = _req(transformed_points, "transformed_points", 4)
        N, three, H, W = p.shape
        if three != 3:
            raise AssertionError(f"wrong size for transformed_points, expected [N,3,H,W], got {list(p.size())}")
        k = _k4(cam_intrinsic, N)
        out = torch.empty(N, 2, H, W, device=p.device, dtype=torch.float32)
        _lib.check(lib.dvf_caffe_pinhole_fwd(_ptr(p), _ptr(k), N, H, W, _ptr(out), _stream()), "dvf_caffe_pinhole_fwd")
        ctx.save_for_backward(p, k)
        ctx.k_shape = tuple(cam_intrinsic.shape)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gcoords):
        lib = _lib.load()
        p, k = ctx.saved_tensors
        N, _, H, W = p.shape
        g = _req(gcoords, "grad_coords", 4)
        gp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        gK = torch.empty(N, 4, device=p.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        _lib.check(lib.dvf_caffe_pinhole_bwd(_ptr(g), _ptr(p), _ptr(k), N, H, W, _ptr(gp), _ptr(gK), _stream()),
                   "dvf_caffe_pinhole_bwd")
        return gp, None if gK is None else gK.reshape(ctx.k_shape)


class PixelWarp(torch.autograd.Function):
    """InverseWarpingLayer (inverse_warping_layer.cu:10-169; geo_transform.py:76-125): img [N,C,H,W] sampled
    bilinearly at proj_coords [N,2,H,W] given in PIXELS; taps outside the image contribute zero."""

    @staticmethod
    def forward(ctx, img, proj_coords):
        lib = _lib.load()
        u = _req(img, "img", 4)
        N, Cc, H, W = u.shape
        xy = _req(proj_coords, "proj_coords", 4)
        if tuple(xy.shape) != (N, 2, H, W):
            raise AssertionError(f"wrong size for proj_coords, expected {[N, 2, H, W]}, got {list(xy.size())}")
        out = torch.empty_like(u)
        _lib.check(lib.dvf_caffe_warp_fwd(_ptr(u), _ptr(xy), N, Cc, H, W, _ptr(out), _stream()), "dvf_caffe_warp_fwd")
        ctx.save_for_backward(u, xy)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        lib = _lib.load()
        u, xy = ctx.saved_tensors
        N, Cc, H, W = u.shape
        g = _req(gout, "grad_output", 4)
        gu = torch.empty_like(u) if ctx.needs_input_grad[0] else None
        gxy = torch.empty_like(xy) if ctx.needs_input_grad[1] else None
        _lib.check(lib.dvf_caffe_warp_bwd(_ptr(g), _ptr(u), _ptr(xy), N, Cc, H, W, _ptr(gu), _ptr(gxy), _stream()),
                   "dvf_caffe_warp_bwd")
        return gu, gxy


class AbsLoss(torch.autograd.Function):
    """AbsLossLayer (abs_loss_layer.cu:10-50): sum|a-b| / a.size(0); d/da = +-1/num with sign(0) := -1."""

    @staticmethod
    def forward(ctx, a, b):
        lib = _lib.load()
        x, y = _req(a, "a"), _req(b, "b")
        if x.shape != y.shape:
            raise AssertionError(f"wrong size for b, expected {list(x.size())}, got {list(y.size())}")
        loss = torch.empty(1, device=x.device, dtype=torch.float32)
        ga = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        ws = workspace(8, x.device, ("abs_loss",))
        _lib.check(lib.dvf_caffe_abs_loss(_ptr(x), _ptr(y), x.numel(), x.shape[0], 1.0, _ptr(loss), _ptr(ga), _ptr(gb), _ptr(ws),
                                          _stream()), "dvf_caffe_abs_loss")
        ctx.save_for_backward(*[t for t in (ga, gb) if t is not None])
        ctx.has = (ga is not None, gb is not None)
        return loss[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        saved = list(ctx.saved_tensors)
        ga = saved.pop(0) * g if ctx.has[0] else None
        gb = saved.pop(0) * g if ctx.has[1] else None
        return ga, gb


class EdgeAwareSmoothness(torch.autograd.Function):
    """Edge-aware smoothness of the Caffe training graphs (experiments/depth/train.prototxt:4022-4234: EdgeX / EdgeY
    convolutions, AbsVal, -0.33 1x1 convolution, Exp, Eltwise PROD, AbsLoss with loss_weight): returns
    weight * (sum|gx * dx(inv_depth)| + sum|gy * dy(inv_depth)|) / N with gx = exp(-0.33 * sum_c |dx(img_c)|); one launch for
    the value and d/d inv_depth.  img [N,3,H,W] is data, inv_depth [N,1,H,W]."""

    @staticmethod
    def forward(ctx, img, inv_depth, weight):
        lib = _lib.load()
        I, D = _req(img, "img", 4), _req(inv_depth, "inv_depth", 4)
        N, three, H, W = I.shape
        if three != 3 or tuple(D.shape) != (N, 1, H, W):
            raise AssertionError(f"wrong sizes, expected img [N,3,H,W] and inv_depth [N,1,H,W], got {list(I.shape)} and {list(D.shape)}")
        if ctx.needs_input_grad[0]:
            raise DvfError("the image is data in the reference graph (lr_mult 0 edge convolutions): no gradient w.r.t. img")
        terms = torch.empty(2, device=I.device, dtype=torch.float32)
        g = torch.empty_like(D) if ctx.needs_input_grad[1] else None
        ws = workspace(16, I.device, ("edge_smooth",))
        with _same_device(I, D):
            _lib.check(lib.dvf_caffe_edge_smooth_loss(_ptr(I), _ptr(D), N, H, W, float(weight), _ptr(terms), _ptr(g), _ptr(ws),
                                                      _stream()), "dvf_caffe_edge_smooth_loss")
        ctx.unit_grad = g
        return terms.sum() * float(weight)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        return None, (None if ctx.unit_grad is None else ctx.unit_grad * gout), None


def edge_aware_smoothness(img, inv_depth, weight=10.0):
    return EdgeAwareSmoothness.apply(img, inv_depth, weight)


class SSIMLoss(torch.autograd.Function):
    """SSIM reconstruction term (csrc/dvf_ssim.cu; NOT in the reference -- new functionality, parity unpinned): 3x3
    average-pool SSIM, loss = mean over windows of clamp((1 - SSIM(x, y)) / 2, 0, 1), windows that touch an invalid pixel
    contribute zero.  x = target (data), y = warped image (differentiable), valid = uint8 [B,H,W] or None."""

    @staticmethod
    def forward(ctx, x, y, valid):
        lib = _lib.load()
        xs, ys = _req(x, "x", 4), _req(y, "y", 4)
        if xs.shape != ys.shape:
            raise AssertionError(f"wrong size for y, expected {list(xs.size())}, got {list(ys.size())}")
        if ctx.needs_input_grad[0]:
            raise DvfError("ssim_loss differentiates w.r.t. its second argument (the warped image) only")
        B, Cc, H, W = xs.shape
        v = None
        if valid is not None:
            if valid.dtype != torch.uint8 or tuple(valid.shape) != (B, H, W) or not valid.is_cuda:
                raise AssertionError(f"wrong valid mask, expected uint8 CUDA [B,H,W], got {valid.dtype} {list(valid.size())}")
            v = valid.contiguous()
        loss = torch.empty(1, device=xs.device, dtype=torch.float32)
        gy = torch.empty_like(ys) if ctx.needs_input_grad[1] else None
        ws = workspace(8, xs.device, ("ssim",))
        with _same_device(xs, ys, v):
            _lib.check(lib.dvf_ssim_loss(_ptr(xs), _ptr(ys), _ptr(v), B, Cc, H, W, _ptr(loss), _ptr(gy), _ptr(ws), _stream()),
                       "dvf_ssim_loss")
        ctx.unit_grad = gy
        return loss[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        return None, (None if ctx.unit_grad is None else ctx.unit_grad * gout), None


def ssim_loss(x, y, valid=None):
    return SSIMLoss.apply(x, y, valid)
