"""torch-op restatement of the reference's hot path (TEST / BASELINE INFRASTRUCTURE, not product).

oracle/dvf_oracle.c pins the arithmetic; this module restates the same path with the SAME torch
operator sequence the reference issues (pytorch_version/inverse_warp.py:26-193,
loss_functions.py:7-20, loss_functions_sfm.py:9-46), so that timing it on the host cores is a fair
stand-in for "the reference's CPU PyTorch path" on machines where the reference checkout is absent
(the GPU box).  bench.py uses it for the cpu_baseline / --impl reference legs and, on the GPU, as
the "unfused stock-PyTorch" comparison; tests check it against the golden vectors.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this file.
"""
from __future__ import annotations

import warnings

import torch
import torch.nn.functional as F

_grid_cache = {}


def _pixel_grid(depth):
    # inverse_warp.py:8-15 builds (j, i, 1) once and slices it (:38); cache per (h, w, dtype, device)
    b, h, w = depth.shape
    key = (h, w, depth.dtype, depth.device)
    g = _grid_cache.get(key)
    if g is None:
        ii = torch.arange(0, h).view(1, h, 1).expand(1, h, w).type_as(depth)
        jj = torch.arange(0, w).view(1, 1, w).expand(1, h, w).type_as(depth)
        g = torch.stack((jj, ii, torch.ones(1, h, w).type_as(depth)), dim=1)
        _grid_cache[key] = g
    return g


def euler_matrix(angle):
    # inverse_warp.py:77-114
    B = angle.size(0)
    x, y, z = angle[:, 0], angle[:, 1], angle[:, 2]
    cz, sz = torch.cos(z), torch.sin(z)
    zero = z.detach() * 0
    one = zero.detach() + 1
    zm = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], dim=1).reshape(B, 3, 3)
    cy, sy = torch.cos(y), torch.sin(y)
    ym = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], dim=1).reshape(B, 3, 3)
    cx, sx = torch.cos(x), torch.sin(x)
    xm = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], dim=1).reshape(B, 3, 3)
    return xm @ ym @ zm


def quat_matrix(quat):
    # inverse_warp.py:117-138
    q = torch.cat([quat[:, :1].detach() * 0 + 1, quat], dim=1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    B = quat.size(0)
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz = w * x, w * y, w * z
    xy, xz, yz = x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).reshape(B, 3, 3)


def pose_matrix(vec, rotation_mode="euler"):
    # inverse_warp.py:141-157
    t = vec[:, :3].unsqueeze(-1)
    R = euler_matrix(vec[:, 3:]) if rotation_mode == "euler" else quat_matrix(vec[:, 3:])
    return torch.cat([R, t], dim=2)


def warp(img, depth, pose, K, Kinv, rotation_mode="euler", padding_mode="zeros", align_corners=False):
    """inverse_warp.py:160-193 (pixel2cam :26-40 and cam2pixel :43-74 inlined, same op order).  align_corners=True is the
    grid_sample convention of torch <= 1.2, which the reference was written for; today's torch runs it with False."""
    b, h, w = depth.shape
    grid = _pixel_grid(depth)[:, :, :h, :w].expand(b, 3, h, w).reshape(b, 3, -1)
    cam = (Kinv @ grid).reshape(b, 3, h, w) * depth.unsqueeze(1)
    P = K @ pose_matrix(pose, rotation_mode)
    pc = P[:, :, :3] @ cam.reshape(b, 3, -1) + P[:, :, -1:]
    X, Y = pc[:, 0], pc[:, 1]
    Z = pc[:, 2].clamp(min=1e-3)
    Xn = 2 * (X / Z) / (w - 1) - 1
    Yn = 2 * (Y / Z) / (h - 1) - 1
    if padding_mode == "zeros":
        Xm = ((Xn > 1) + (Xn < -1)).detach()
        Xn[Xm] = 2
        Ym = ((Yn > 1) + (Yn < -1)).detach()
        Yn[Ym] = 2
    coords = torch.stack([Xn, Yn], dim=2).reshape(b, h, w, 2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # the reference does not pass align_corners (-> False + UserWarning)
        if align_corners:
            return F.grid_sample(img, coords, padding_mode=padding_mode, align_corners=True)
        return F.grid_sample(img, coords, padding_mode=padding_mode)


def _masked_l1(tgt, warped, expl=None):
    valid = 1 - (warped == 0).prod(1, keepdim=True).type_as(warped)
    diff = (tgt - warped) * valid
    if expl is not None:
        diff = diff * expl.expand_as(diff)
    return diff.abs().mean()


def loss_two_view(img_R2, img_R1, img_L2, depth, T_2to1, T_R2L, K, Kinv, rotation_mode="euler", padding_mode="zeros"):
    """loss_functions.py:7-20."""
    total = 0
    total = total + _masked_l1(img_R2, warp(img_R1, depth, T_2to1, K, Kinv, rotation_mode, padding_mode))
    total = total + _masked_l1(img_R2, warp(img_L2, depth, T_R2L, K, Kinv, rotation_mode, padding_mode))
    return total


def loss_multi_scale(tgt, refs, K, Kinv, depths, masks, pose, rotation_mode="euler", padding_mode="zeros"):
    """loss_functions_sfm.py:9-46 (without the per-view NaN assert, which only adds a sync)."""
    if type(masks) not in (tuple, list):
        masks = [masks]
    if type(depths) not in (tuple, list):
        depths = [depths]
    total = 0
    for d, m in zip(depths, masks):
        b, _, h, w = d.size()
        ds = tgt.size(2) / h
        tgt_s = F.interpolate(tgt, (h, w), mode="area")
        refs_s = [F.interpolate(r, (h, w), mode="area") for r in refs]
        Ks = torch.cat((K[:, 0:2] / ds, K[:, 2:]), dim=1)
        Kis = torch.cat((Kinv[:, :, 0:2] * ds, Kinv[:, :, 2:]), dim=2)
        for i, r in enumerate(refs_s):
            wv = warp(r, d[:, 0], pose[:, i], Ks, Kis, rotation_mode, padding_mode)
            total = total + _masked_l1(tgt_s, wv, None if m is None else m[:, i:i + 1])
    return total
