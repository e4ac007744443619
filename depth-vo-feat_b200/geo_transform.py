"""Drop-in for pytorch_version/geo_transform.py of Depth-VO-Feat (imported by unsupervise_dvo.py:27, called at
:104-108).  The reference file is an unfinished transliteration of the Caffe layers (it calls exit(0) inside
geo_transform, geo_transform.py:31, and spawns a Python thread per pixel in inverse_warp); this module provides
the same three names with the semantics of the Caffe kernels they transliterate
(caffe/src/caffe/layers/geometry_transformation.cu, pin_hole_layer.cu, inverse_warping_layer.cu), as CUDA
kernels with autograd.  abs_loss is AbsLossLayer (abs_loss_layer.cu), the loss those prototxts attach.

    pts3D       = geo_transform(depth[N,1,H,W], SE3[N,1,4,4], K[N,4,1,1])     # K = (fx,fy,cx,cy)
    proj_coords = pin_hole_project(pts3D, K)                                   # [N,2,H,W], pixels
    warp_Itgt   = inverse_warp(Isrc[N,C,H,W], proj_coords)

CUDA tensors only; parity of this family is "unpinned" (Caffe cannot be built here -- see DESIGN.md)."""
from dvf_b200 import ops as _ops

geo_transform = _ops.GeoTransform.apply
pin_hole_project = _ops.PinHoleProject.apply
inverse_warp = _ops.PixelWarp.apply
abs_loss = _ops.AbsLoss.apply

edge_aware_smoothness = _ops.edge_aware_smoothness   # experiments/depth/train.prototxt:4022-4234


def two_view_warp_errors(inv_depth_imR2, T_R2L, T_2to1, K, norm_imL2, norm_imR1, norm_imR2, se3=None):
    """The batch-concatenated two-view form of the Caffe depth-odometry graph
    (experiments/depth_odometry/train.prototxt:4309-4437): the stereo and the temporal view share ONE pass through the
    geometry layers by concatenation along the batch axis --
        T = Concat(T_R2L, T_2to1), SE3 = SE3_Generator(T), inv_depth = Concat(inv_depth_imR2, inv_depth_imR2),
        depth = Power(inv_depth: power -1, scale 1, shift 1e-4), pts3D = GeoTransform(depth, SE3, K), proj = PinHole(pts3D, K),
        warp = InverseWarping(Concat(imL2, imR1), proj), (warp_LR, warp_R12) = Slice(warp),
        Warp_error_LR = AbsLoss(warp_LR, imR2), Warp_error_R12 = AbsLoss(warp_R12, imR2).
    The layers are per-image kernels, so the concatenated batch is just a batch of 2N images: one launch per layer instead
    of two.  T_*: [N,6,1,1] se(3) vectors (w, u) as SE3_Generator_KITTI takes them, or pass ready [N,1,4,4] matrices and
    se3=False.  K: [N,4,1,1] (fx, fy, cx, cy), the same intrinsics for both views.  Returns (Warp_error_LR, Warp_error_R12)."""
    import torch
    from se3_generate import SE3_Generator_KITTI
    T = torch.cat((T_R2L, T_2to1), dim=0)
    if se3 is None:
        se3 = T.dim() == 4 and T.shape[1] == 6
    SE3 = SE3_Generator_KITTI.apply(T) if se3 else T
    inv_depth = torch.cat((inv_depth_imR2, inv_depth_imR2), dim=0)
    depth = (inv_depth + 1e-4).pow(-1)              # Power layer: (shift + scale * x) ^ power = (1e-4 + x) ^ -1
    K2 = torch.cat((K, K), dim=0)
    pts3D = geo_transform(depth, SE3, K2)
    proj_coords = pin_hole_project(pts3D, K2)
    warp = inverse_warp(torch.cat((norm_imL2, norm_imR1), dim=0), proj_coords)
    n = norm_imR2.shape[0]
    return abs_loss(warp[:n].contiguous(), norm_imR2), abs_loss(warp[n:].contiguous(), norm_imR2)
