"""Drop-in for pytorch_version/loss_function_sfm_old.py of Depth-VO-Feat (signature used by
unsupervise_sfm.py:98; that script imports it under the name `loss_function_sfm`, see the alias
module next to this file).  Multi-scale, temporal view only (the stereo term is commented out in
the reference, :28-34), optional mask channel 0 (:25).  CUDA tensors only.
"""
from __future__ import division

from dvf_b200 import ops as _ops
from loss_functions import smooth_loss, inverse_warp  # noqa: F401
from loss_functions_sfm import explainability_loss  # noqa: F401


def photometric_reconstruction_loss(img_R2, img_R1, img_L2, depth, T_2to1, T_R2L, mask, intrinsics, intrinsics_inv,
                                    rotation_mode='euler', padding_mode='zeros'):
    """loss_function_sfm_old.py:7-46."""
    masks = list(mask) if type(mask) in (tuple, list) else [mask]
    depths = list(depth) if type(depth) in (tuple, list) else [depth]
    n = min(len(depths), len(masks))
    depths, masks = depths[:n], masks[:n]
    sizes = [(d.size(2), d.size(3)) for d in depths]
    downscales = [img_R2.size(2) / s[0] for s in sizes]
    tgt_pyr = _ops.area_pyramid(img_R2, sizes)
    r1_pyr = _ops.area_pyramid(img_R1, sizes)
    has_mask = masks[0] is not None
    loss, _ = _ops.fused_photo_loss(tgt_pyr, [[r] for r in r1_pyr], [d[:, 0] for d in depths], T_2to1.unsqueeze(1),
                                    intrinsics, intrinsics_inv, expl_levels=masks if has_mask else None,
                                    downscales=downscales, rotation_mode=rotation_mode, padding_mode=padding_mode)
    return loss
