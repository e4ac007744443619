"""Drop-in for pytorch_version/loss_functions_sfm.py of Depth-VO-Feat (used by train.py:25).

The multi-scale, multi-view, explainability-weighted photometric loss is ONE fused CUDA launch over
all scales and views (csrc/dvf_loss.cu) preceded by one pass that builds the 'area' pyramids
(csrc/dvf_aux.cu).  The NaN assertion of the reference (:34, a device synchronisation per view and scale)
becomes an in-kernel flag: with NAN_CHECK = True the launch ORs bit level*V+view into a device word when
that loss term is NaN, and assert_no_nan() -- called whenever the training loop likes, e.g. once per
logging interval -- performs the reference's assertion with ONE synchronisation.  NaNs propagate into the
returned loss either way.  CUDA tensors only.
"""
from __future__ import division

import torch
from torch import nn

from dvf_b200 import ops as _ops
from inverse_warp import inverse_warp  # noqa: F401  (re-exported like the reference, :6)


NAN_CHECK = False   # True: collect NaN loss terms in dvf_b200.ops.nan_flags() (DVF_FLAG_NAN_CHECK)


def assert_no_nan():
    """The reference's `assert((reconstruction_loss == reconstruction_loss).item() == 1)` (:34) for every loss term
    computed since the flags were last cleared; clears them."""
    flags = _ops.nan_flags()
    bits = int(flags.item())
    flags.zero_()
    assert bits == 0, "NaN reconstruction loss at (scale, view) bits {:#b}".format(bits)


def _as_list(x):
    return list(x) if type(x) in (tuple, list) else [x]


def photometric_reconstruction_loss(tgt_img, ref_imgs, intrinsics, intrinsics_inv, depth, explainability_mask, pose,
                                    rotation_mode='euler', padding_mode='zeros', disparity_eps=None):
    """loss_functions_sfm.py:9-46.  disparity_eps (extension): when not None, `depth` holds the DispNet disparities and
    the kernel evaluates depth = 1 / (disp + disparity_eps) itself -- train.py:188's `depth = [1/disp for disp in
    disparities]` (eps 0) folded into the launch; gradients then flow to the disparities."""
    masks = _as_list(explainability_mask)
    depths = _as_list(depth)
    n = min(len(depths), len(masks))           # zip() semantics of :44
    depths, masks = depths[:n], masks[:n]
    assert pose.size(1) == len(ref_imgs)        # :12
    has_mask = masks[0] is not None
    sizes = []
    for d, m in zip(depths, masks):
        assert m is None or d.size()[2:] == m.size()[2:]   # :11
        sizes.append((d.size(2), d.size(3)))
    H = tgt_img.size(2)
    downscales = [H / s[0] for s in sizes]      # :16
    tgt_pyr = _ops.area_pyramid(tgt_img, sizes)
    ref_pyr = [_ops.area_pyramid(r, sizes) for r in ref_imgs]
    src_levels = [[ref_pyr[v][l] for v in range(len(ref_imgs))] for l in range(n)]
    loss, _ = _ops.fused_photo_loss(tgt_pyr, src_levels, [d[:, 0] for d in depths], pose, intrinsics, intrinsics_inv,
                                    expl_levels=masks if has_mask else None, downscales=downscales,
                                    rotation_mode=rotation_mode, padding_mode=padding_mode, nan_check=NAN_CHECK,
                                    disparity_eps=disparity_eps)
    return loss


def explainability_loss(mask):
    """loss_functions_sfm.py:49-56: sum over scales of BCE(mask, 1); one fused CUDA launch (csrc/dvf_reg.cu)."""
    return _ops.explainability_loss(mask)


def smooth_loss(pred_map, scale_factor):
    """loss_functions_sfm.py:59-77."""
    from loss_functions import smooth_loss as _smooth
    return _smooth(pred_map, scale_factor)


@torch.no_grad()
def compute_errors(gt, pred, crop=True):
    """loss_functions_sfm.py:80-116: Eigen depth metrics (abs_diff, abs_rel, sq_rel, a1, a2, a3), batch-averaged."""
    B = gt.size(0)
    sums = [0.0] * 6
    if crop:  # Garg ECCV16 crop
        y1, y2 = int(0.40810811 * gt.size(1)), int(0.99189189 * gt.size(1))
        x1, x2 = int(0.03594771 * gt.size(2)), int(0.96405229 * gt.size(2))
        crop_mask = torch.zeros_like(gt[0], dtype=torch.bool)
        crop_mask[y1:y2, x1:x2] = True
    for g, p in zip(gt, pred):
        valid = (g > 0) & (g < 80)
        if crop:
            valid = valid & crop_mask
        vg = g[valid]
        vp = p[valid].clamp(1e-3, 80)
        vp = vp * torch.median(vg) / torch.median(vp)
        ratio = torch.max(vg / vp, vp / vg)
        err = (vg - vp).abs()
        vals = [err.mean(), (err / vg).mean(), ((vg - vp) ** 2 / vg).mean(),
                (ratio < 1.25).float().mean(), (ratio < 1.25 ** 2).float().mean(), (ratio < 1.25 ** 3).float().mean()]
        sums = [s + v for s, v in zip(sums, vals)]
    return [float(s) / B for s in sums]
