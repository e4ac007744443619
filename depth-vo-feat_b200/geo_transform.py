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
