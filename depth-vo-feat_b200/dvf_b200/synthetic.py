"""Deterministic KITTI-shaped synthetic inputs for the warp + reconstruction-loss path.

The reference's loaders (pytorch_version/un_dataset.py:43-84) yield
(img_R1, img_L2, img_R2, K[3,3], K^-1, raw_K, T_R2L) from KITTI on disk; there
is no dataset here, so tests, smoke() and bench.py draw tensors of the same
shapes and value ranges instead:

* images   : box-filtered uniform noise scaled by 255*0.004 (unsupervise.py:101
             multiplies uint8-range images by 0.004) or plain U[0,1) noise;
* depth    : 1/disp with disp in [0.02, 0.32] (DispNetS.py:112 range), either a
             17x17 box-filtered field (network-like, primary) or iid (stress);
* K        : first record of data/kitti_eigen/train_K (fx,fy,cx,cy at 160x608 =
             352.187, 305.896, 297.480, 78.815) rescaled to HxW;
* poses    : (tx,ty,tz,rx,ry,rz) order of inverse_warp.py:146; temporal
             'kitti' = forward motion ~0.85 m with small rotation, 'tiny' =
             N(0,0.01^2) (PoseExpNet at init), stereo = (0.53233,0,0,0,0,0).

Everything is generated on the CPU with an explicit torch.Generator so that the
same seed gives the same tensors on every machine; callers move them to a device.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

KITTI_K_160x608 = (352.187, 305.896, 297.480, 78.815)  # fx, fy, cx, cy
STEREO_BASELINE = 0.53233


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def box_filter(x, k):
    """k x k mean filter with reflect padding on [...,H,W]."""
    if k <= 1:
        return x
    shp = x.shape
    x4 = x.reshape(-1, 1, shp[-2], shp[-1])
    pad = k // 2
    ph, pw = min(pad, shp[-2] - 1), min(pad, shp[-1] - 1)
    x4 = F.pad(x4, (pw, pw, ph, ph), mode="reflect")
    x4 = F.avg_pool2d(x4, (2 * ph + 1, 2 * pw + 1), stride=1)
    return x4.reshape(shp)


def intrinsics(B, H, W):
    fx, fy, cx, cy = KITTI_K_160x608
    K = torch.tensor([[fx * W / 608.0, 0.0, cx * W / 608.0],
                      [0.0, fy * H / 160.0, cy * H / 160.0],
                      [0.0, 0.0, 1.0]], dtype=torch.float32)
    K = K.unsqueeze(0).repeat(B, 1, 1).contiguous()
    return K, torch.inverse(K).contiguous()


def images(B, C, H, W, seed, smooth=True, n=1):
    g = _gen(seed)
    out = []
    for _ in range(n):
        x = torch.rand(B, C, H, W, generator=g)
        if smooth:
            x = box_filter(x, 9)
            # stretch back to ~[0,1] so the photometric signal is not flat
            x = ((x - 0.5) * 4.0 + 0.5).clamp_(0.0, 1.0)
        out.append((x * (255.0 * 0.004)).contiguous())
    return out


def features(B, C, H, W, seed, n=1):
    g = _gen(seed)
    return [box_filter(torch.randn(B, C, H, W, generator=g), 5).mul_(3.0).contiguous() for _ in range(n)]


def depth(B, H, W, seed, smooth=True):
    g = _gen(seed)
    disp = torch.rand(B, H, W, generator=g)
    if smooth:
        disp = box_filter(disp, 17)
        disp = ((disp - 0.5) * 6.0 + 0.5).clamp_(0.0, 1.0)
    disp = disp * 0.3 + 0.02
    return (1.0 / disp).contiguous()


def pose(B, kind, seed):
    g = _gen(seed)
    if kind == "stereo":
        p = torch.zeros(B, 6)
        p[:, 0] = STEREO_BASELINE
    elif kind == "kitti":
        p = torch.empty(B, 6)
        p[:, 0:2] = torch.randn(B, 2, generator=g) * 0.02
        p[:, 2] = -0.85 + torch.randn(B, generator=g) * 0.05
        p[:, 3:] = torch.randn(B, 3, generator=g) * 0.005
    elif kind == "tiny":
        p = torch.randn(B, 6, generator=g) * 0.01
    elif kind == "large":
        p = torch.randn(B, 6, generator=g) * torch.tensor([0.5, 0.2, 0.5, 0.1, 0.2, 0.1])
    else:
        raise ValueError(kind)
    return p.contiguous()


def explainability(B, V, H, W, seed):
    g = _gen(seed)
    return torch.sigmoid(torch.randn(B, V, H, W, generator=g)).contiguous()


def stereo_temporal_batch(B, H, W, seed=0, C=3, smooth=True, temporal="kitti", feature=False):
    """The (R2 target, R1 temporal source, L2 stereo source) sample of
    un_dataset.py:78-84 plus depth/poses/intrinsics, as a dict of CPU tensors."""
    if feature:
        tgt, r1, l2 = features(B, C, H, W, seed + 1, n=3)
    else:
        tgt, r1, l2 = images(B, C, H, W, seed + 1, smooth=smooth, n=3)
    K, Kinv = intrinsics(B, H, W)
    return dict(img_R2=tgt, img_R1=r1, img_L2=l2,
                depth=depth(B, H, W, seed + 2, smooth=smooth),
                T_2to1=pose(B, temporal, seed + 3), T_R2L=pose(B, "stereo", seed + 4),
                intrinsics=K, intrinsics_inv=Kinv)
