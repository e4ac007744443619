"""Drop-in for pytorch_version/inverse_warp.py of Depth-VO-Feat, backed by libdvf_b200.so.

Same public names, argument order, defaults and assertion behaviour as the reference module
(reference lines cited per function); the arithmetic runs in hand-written CUDA kernels
(csrc/dvf_pose.cu, csrc/dvf_warp.cu).  CUDA tensors only -- there is no CPU fallback.
"""
from __future__ import division

import torch

from dvf_b200 import ops as _ops

pixel_coords = None  # kept for API compatibility (inverse_warp.py:5); the kernels derive (j, i) from thread indices


def set_id_grid(depth):
    """inverse_warp.py:8-15.  No-op: there is no cached pixel grid to (re)build."""
    return None


def check_sizes(input, input_name, expected):
    """inverse_warp.py:18-23: same assertion text."""
    ok = input.ndimension() == len(expected)
    if ok:
        for k, size in enumerate(expected):
            if size.isdigit() and input.size(k) != int(size):
                ok = False
    assert ok, "wrong size for {}, expected {}, got  {}".format(input_name, 'x'.join(expected), list(input.size()))


def pixel2cam(depth, intrinsics_inv):
    """inverse_warp.py:26-40: depth [B,H,W], intrinsics_inv [B,3,3] -> camera-frame points [B,3,H,W]."""
    return _ops.pixel2cam(depth, intrinsics_inv)


def cam2pixel(cam_coords, proj_c2p_rot, proj_c2p_tr, padding_mode):
    """inverse_warp.py:43-74: camera-frame points -> normalised sampling grid [B,H,W,2]."""
    return _ops.cam2pixel(cam_coords, proj_c2p_rot, proj_c2p_tr, padding_mode)


def _rotation(vec3, mode):
    zeros = torch.zeros_like(vec3)
    return _ops.PoseVec2Mat.apply(torch.cat([zeros, vec3], dim=1), mode)[:, :, :3]


def euler2mat(angle):
    """inverse_warp.py:77-114: [B,3] euler angles -> [B,3,3] (R = Rx @ Ry @ Rz)."""
    return _rotation(angle, 'euler')


def quat2mat(quat):
    """inverse_warp.py:117-138: last three quaternion coefficients [B,3] -> [B,3,3]."""
    return _rotation(quat, 'quat')


def pose_vec2mat(vec, rotation_mode='euler'):
    """inverse_warp.py:141-157: (tx,ty,tz,rx,ry,rz) [B,6] -> [R|t] [B,3,4]."""
    assert rotation_mode in ('euler', 'quat')
    return _ops.PoseVec2Mat.apply(vec, rotation_mode)


def inverse_warp(img, depth, pose, intrinsics, intrinsics_inv, rotation_mode='euler', padding_mode='zeros',
                 check_channels=True):
    """inverse_warp.py:160-193: warp the source image `img` [B,3,H,W] onto the target view given the
    target depth [B,H,W], the target->source pose [B,6] and the intrinsics [B,3,3] (+ inverse).
    check_channels=False reproduces the copies embedded in loss_functions.py (:211), which accept any C."""
    check_sizes(img, 'img', 'B3HW' if check_channels else 'BCHW')
    check_sizes(depth, 'depth', 'BHW')
    check_sizes(pose, 'pose', 'B6')
    check_sizes(intrinsics, 'intrinsics', 'B33')
    check_sizes(intrinsics_inv, 'intrinsics', 'B33')
    assert intrinsics_inv.size() == intrinsics.size()
    return _ops.InverseWarp.apply(img, depth, pose, intrinsics, intrinsics_inv, rotation_mode, padding_mode)
