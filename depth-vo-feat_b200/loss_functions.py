"""Drop-in for pytorch_version/loss_functions.py of Depth-VO-Feat (used by unsupervise.py:28).

photometric_reconstruction_loss runs as ONE fused CUDA launch (csrc/dvf_loss.cu) that warps both
source views, applies the value-based validity masks and produces the loss together with its
gradients; smooth_loss is one fused launch over all scales.  CUDA tensors only.
"""
from __future__ import division

import torch

from dvf_b200 import ops as _ops
from inverse_warp import inverse_warp as _inverse_warp_checked


def photometric_reconstruction_loss(img_R2, img_R1, img_L2, depth, T_2to1, T_R2L, intrinsics, intrinsics_inv,
                                    rotation_mode='euler', padding_mode='zeros'):
    """loss_functions.py:7-20: mean|(R2 - warp(R1; T_2to1)) * valid| + mean|(R2 - warp(L2; T_R2L)) * valid|.
    Works for images (C=3) and for feature maps (any C, gradients flow to all three maps)."""
    pose = torch.stack((T_2to1, T_R2L), dim=1)  # [B,2,6]
    loss, _ = _ops.fused_photo_loss([img_R2], [[img_R1, img_L2]], [depth], pose, intrinsics, intrinsics_inv,
                                    rotation_mode=rotation_mode, padding_mode=padding_mode)
    return loss


def photometric_reconstruction_loss_fused_inputs(img_R2, img_R1, img_L2, inv_depth, T_2to1, T_R2L, intrinsics, intrinsics_inv,
                                                rotation_mode='euler', padding_mode='zeros', img_scale=0.004, eps=1e-4):
    """unsupervise.py:99-101 in one launch: the caller's
        depth = (1/(inv_depth+1e-4)).squeeze(1)
        photometric_reconstruction_loss(0.004*img_R2, 0.004*img_R1, 0.004*img_L2, depth, ...)
    with the reciprocal and the three image scalings folded into the fused kernel (they cost more HBM traffic as separate
    torch passes than the loss itself).  inv_depth [B,1,H,W] or [B,H,W]: the DispNet output; gradients flow to it."""
    disp = inv_depth.squeeze(1) if inv_depth.dim() == 4 else inv_depth
    pose = torch.stack((T_2to1, T_R2L), dim=1)
    loss, _ = _ops.fused_photo_loss([img_R2], [[img_R1, img_L2]], [disp], pose, intrinsics, intrinsics_inv,
                                    rotation_mode=rotation_mode, padding_mode=padding_mode, disparity_eps=eps,
                                    img_scale=img_scale)
    return loss


def smooth_loss(pred_map, scale_factor=1):
    """loss_functions.py:23-41: second-order smoothness, sum over scales with weight /= scale_factor.
    One fused CUDA launch for all scales (value + gradient, csrc/dvf_reg.cu)."""
    return _ops.smooth_loss(pred_map, scale_factor)


def inverse_warp(img, depth, pose, intrinsics, intrinsics_inv, rotation_mode='euler', padding_mode='zeros'):
    """The copy embedded in loss_functions.py:198-231 has the 'B3HW' check disabled (:211)."""
    return _inverse_warp_checked(img, depth, pose, intrinsics, intrinsics_inv, rotation_mode, padding_mode,
                                 check_channels=False)


def photometric_ssim_reconstruction_loss(img_R2, img_R1, img_L2, depth, T_2to1, T_R2L, intrinsics, intrinsics_inv,
                                         rotation_mode='euler', padding_mode='zeros', alpha=0.85):
    """NOT in the reference (it has no SSIM term; BASELINE's north_star names "masked photometric (L1/SSIM)"): the usual
    monocular-depth mix  (1 - alpha) * L1 + alpha * SSIM  over the same two views and the same validity masks as
    photometric_reconstruction_loss above.  The L1 part is the fused launch; the SSIM part warps each source image
    (dvf_inverse_warp_fwd, which also yields the mask), evaluates dvf_ssim_loss (value + d/d warped in one launch) and
    back-propagates through dvf_inverse_warp_bwd.  Parity unpinned (csrc/dvf_ssim.cu states the definition)."""
    l1 = photometric_reconstruction_loss(img_R2, img_R1, img_L2, depth, T_2to1, T_R2L, intrinsics, intrinsics_inv,
                                         rotation_mode, padding_mode)
    ssim = 0
    for src, pose in ((img_R1, T_2to1), (img_L2, T_R2L)):
        warped = inverse_warp(src, depth, pose, intrinsics, intrinsics_inv, rotation_mode, padding_mode)
        valid = (warped.detach() != 0).any(1).to(torch.uint8)
        ssim = ssim + _ops.ssim_loss(img_R2, warped, valid)
    return (1.0 - alpha) * l1 + alpha * ssim
