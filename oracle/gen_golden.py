"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (read-only checkout at
/root/reference/pytorch_version) on torch-CPU fp32.  Runs only in the build container -- the GPU
box has no reference checkout -- so the vectors are committed together with this script.

    python oracle/gen_golden.py            # rewrites tests/golden/

Each fixture holds the inputs, the reference's intermediate projection matrix P = K @ pose_mat
(so the per-pixel path can be checked bit-for-bit independently of sin/cos implementations) and
the reference's outputs and autograd gradients.
"""
from __future__ import annotations

import importlib
import os
import sys
import warnings

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DVF_REFERENCE", "/root/reference/pytorch_version")
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, os.path.join(REPO, "depth-vo-feat_b200"))
from dvf_b200 import synthetic as syn  # noqa: E402  (pure-torch input generator, no CUDA needed)


def _ref_modules():
    # the package dir of this repo has same-named drop-in modules: make sure the reference wins here
    sys.path = [p for p in sys.path if not p.endswith("depth-vo-feat_b200")]
    sys.path.insert(0, REF)
    for name in ("inverse_warp", "loss_functions", "loss_functions_sfm", "loss_function_sfm_old"):
        sys.modules.pop(name, None)
    mods = {n: importlib.import_module(n) for n in ("inverse_warp", "loss_functions", "loss_functions_sfm",
                                                     "loss_function_sfm_old")}
    for m in mods.values():
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(REF)), m.__file__
    return mods


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _reset(mods):
    for m in mods.values():
        if hasattr(m, "pixel_coords"):
            m.pixel_coords = None


class _align_corners_true:
    """Run the (unmodified) reference with F.grid_sample(align_corners=True): the sampling convention of torch <= 1.2,
    which the reference was written for and never states (inverse_warp.py:191 passes no align_corners)."""

    def __enter__(self):
        import torch.nn.functional as F
        self.F, self.orig = F, F.grid_sample
        F.grid_sample = lambda img, grid, *a, **k: self.orig(img, grid, *a, **{**k, "align_corners": True})

    def __exit__(self, *exc):
        self.F.grid_sample = self.orig


def case_inverse_warp(mods, name, B, C, H, W, kind, rot, pad, seed, smooth=True, feature=False):
    iw = mods["loss_functions"] if C != 3 else mods["inverse_warp"]   # the copy in loss_functions accepts any C
    d = syn.stereo_temporal_batch(B, H, W, seed=seed, C=C, smooth=smooth, temporal=kind if kind != "stereo" else "kitti",
                                  feature=feature)
    pose = d["T_R2L"] if kind == "stereo" else d["T_2to1"]
    img = d["img_R1"].clone().requires_grad_(True)
    depth = d["depth"].clone().requires_grad_(True)
    pose_t = pose.clone().requires_grad_(True)
    K, Kinv = d["intrinsics"], d["intrinsics_inv"]
    _reset(mods)
    warped = iw.inverse_warp(img, depth, pose_t, K, Kinv, rot, pad)
    gout = torch.randn(warped.shape, generator=torch.Generator().manual_seed(seed + 100))
    warped.backward(gout)
    posemat = mods["inverse_warp"].pose_vec2mat(pose, rot)
    P = K @ posemat
    valid = (1 - (warped == 0).prod(1)).to(torch.uint8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="inverse_warp", rotation_mode=rot, padding_mode=pad,
                        img=_np(img), depth=_np(depth), pose=_np(pose), K=_np(K), Kinv=_np(Kinv), gout=_np(gout),
                        posemat=_np(posemat), P=_np(P), warped=_np(warped), valid=_np(valid),
                        gimg=_np(img.grad), gdepth=_np(depth.grad), gpose=_np(pose_t.grad))


def case_loss_functions(mods, name, B, C, H, W, kind, seed, feature=False):
    lf = mods["loss_functions"]
    d = syn.stereo_temporal_batch(B, H, W, seed=seed, C=C, temporal=kind, feature=feature)
    t = {k: v.clone() for k, v in d.items()}
    req = ["depth", "T_2to1", "T_R2L"] + (["img_R2", "img_R1", "img_L2"] if feature else [])
    for k in req:
        t[k].requires_grad_(True)
    _reset(mods)
    loss = lf.photometric_reconstruction_loss(t["img_R2"], t["img_R1"], t["img_L2"], t["depth"], t["T_2to1"], t["T_R2L"],
                                              t["intrinsics"], t["intrinsics_inv"])
    loss.backward()
    pv = mods["inverse_warp"].pose_vec2mat
    P = torch.stack([d["intrinsics"] @ pv(d["T_2to1"]), d["intrinsics"] @ pv(d["T_R2L"])], 1)
    out = dict(kind="loss_functions", loss=_np(loss), P=_np(P))
    for k, v in d.items():
        out[k] = _np(v)
    for k in req:
        out["g_" + k] = _np(t[k].grad)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def case_sfm(mods, name, B, H, W, n_scales, seed, with_mask, rot, pad, old=False):
    d = syn.stereo_temporal_batch(B, H, W, seed=seed)
    depths = [syn.depth(B, H >> s, W >> s, seed + 10 + s).unsqueeze(1).requires_grad_(True) for s in range(n_scales)]
    nch = 2
    masks = [syn.explainability(B, nch, H >> s, W >> s, seed + 20 + s).requires_grad_(True) if with_mask else None
             for s in range(n_scales)]
    pose = torch.stack([d["T_2to1"], d["T_R2L"]], 1).clone().requires_grad_(True)
    _reset(mods)
    if old:
        t1 = pose[:, 0]
        loss = mods["loss_function_sfm_old"].photometric_reconstruction_loss(
            d["img_R2"], d["img_R1"], d["img_L2"], depths, t1, pose[:, 1], masks, d["intrinsics"], d["intrinsics_inv"], rot, pad)
    else:
        loss = mods["loss_functions_sfm"].photometric_reconstruction_loss(
            d["img_R2"], [d["img_R1"], d["img_L2"]], d["intrinsics"], d["intrinsics_inv"], depths, masks, pose, rot, pad)
    loss.backward()
    out = dict(kind="sfm_old" if old else "sfm", rotation_mode=rot, padding_mode=pad, n_scales=n_scales,
               with_mask=with_mask, loss=_np(loss), pose=_np(pose), g_pose=_np(pose.grad))
    for k in ("img_R2", "img_R1", "img_L2", "intrinsics", "intrinsics_inv"):
        out[k] = _np(d[k])
    pv = mods["inverse_warp"].pose_vec2mat
    for s in range(n_scales):
        out[f"depth{s}"] = _np(depths[s])
        out[f"g_depth{s}"] = _np(depths[s].grad)
        ds = H / (H >> s)
        Ks = torch.cat((d["intrinsics"][:, 0:2] / ds, d["intrinsics"][:, 2:]), dim=1)
        out[f"P{s}"] = _np(torch.stack([Ks @ pv(pose[:, v].detach(), rot) for v in range(2)], 1))
        if with_mask:
            out[f"mask{s}"] = _np(masks[s])
            out[f"g_mask{s}"] = _np(masks[s].grad)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def case_se3(name, B, seed):
    """se3_generate.SE3_Generator_KITTI hard-codes `.cuda()` (se3_generate.py:52,103); there is no GPU where the
    fixtures are generated, so Tensor.cuda is made the identity for the duration of the call -- the reference
    source itself is untouched."""
    sys.path.insert(0, REF)
    sys.modules.pop("se3_generate", None)
    se3 = importlib.import_module("se3_generate")
    assert os.path.realpath(se3.__file__).startswith(os.path.realpath(REF))
    g = torch.Generator().manual_seed(seed)
    vec = torch.randn(B, 6, 1, 1, generator=g) * torch.tensor([0.2, 0.3, 0.1, 1.0, 0.5, 2.0]).view(1, 6, 1, 1)
    vec[0, :3] = 0.0            # exercises the theta^2 < 1e-12 branch
    vec[1, :3] *= 1e-4
    x = vec.clone().requires_grad_(True)
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        out = se3.generate_se3(x)
        gout = torch.randn(out.shape, generator=g, dtype=torch.float64)
        out.backward(gout)
    finally:
        torch.Tensor.cuda = orig
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="se3", vec=_np(vec), out=_np(out), gout=_np(gout), gvec=_np(x.grad))


def case_regularisers(mods, name, B, H, W, n_scales, seed):
    """smooth_loss (loss_functions.py:23-41, scale_factor 2) and explainability_loss (loss_functions_sfm.py:49-56)
    of the reference, value and autograd gradients; the masks include the ends of [0, 1] and denormals."""
    rng = np.random.default_rng(seed)
    maps = [syn.depth(B, H >> s, W >> s, seed + s).unsqueeze(1) for s in range(n_scales)]
    maps[0][:, :, ::2] = torch.round(maps[0][:, :, ::2])       # exact zeros among the second differences
    maps.append(torch.from_numpy(rng.uniform(1, 9, (1, 1, 3, 5)).astype(np.float32)))     # smaller than a strip
    t = [m.clone().requires_grad_(True) for m in maps]
    val = mods["loss_functions"].smooth_loss(t, 2.0)
    val.backward()
    out = dict(kind="regularisers", n_maps=len(maps), smooth=_np(val))
    for i, (m, x) in enumerate(zip(maps, t)):
        out[f"map{i}"] = _np(m)
        out[f"g_map{i}"] = _np(x.grad)
    masks = [syn.explainability(B, 2, H >> s, W >> s, seed + 10 + s) for s in range(n_scales - 1)]
    edge = np.array([0.0, 1.0, 1e-30, 1e-12, 0.5, 1.0 - 2.0 ** -24, 1e-45, 0.999, 2.0 ** -126, 0.25], np.float32)
    masks.append(torch.from_numpy(np.tile(edge, 6).reshape(1, 2, 5, 6).copy()))
    tm = [m.clone().requires_grad_(True) for m in masks]
    ev = mods["loss_functions_sfm"].explainability_loss(tm)
    ev.backward()
    out["n_masks"] = len(masks)
    out["expl"] = _np(ev)
    for i, (m, x) in enumerate(zip(masks, tm)):
        out[f"mask{i}"] = _np(m)
        out[f"g_mask{i}"] = _np(x.grad)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def case_trig(name, n, seed):
    """torch.sin / torch.cos (CPU fp32) -- what euler2mat (inverse_warp.py:89-106) evaluates -- on n angles: PoseExpNet-scale
    (0.01 N(0,1)), KITTI-like (0.005 N(0,1)), unit-scale, uniform over +-pi, +-100, +-10000 and special values."""
    g = torch.Generator().manual_seed(seed)
    k = n // 8
    parts = [torch.randn(2 * k, generator=g) * 0.01, torch.randn(2 * k, generator=g) * 0.005, torch.randn(k, generator=g),
             (torch.rand(k, generator=g) * 2 - 1) * float(np.pi), (torch.rand(k, generator=g) * 2 - 1) * 100.0,
             (torch.rand(k, generator=g) * 2 - 1) * 10000.0]
    special = torch.tensor([0.0, -0.0, float(np.pi), -float(np.pi), float(np.pi) / 2, -float(np.pi) / 2, 1e-30, -1e-40, 1e-45,
                            10000.0, -10000.0, 0.53233, 1.0, -1.0, 3.0, 6.2831855, 1.5707964, 0.7853982], dtype=torch.float32)
    x = torch.cat(parts + [special]).to(torch.float32)
    xs = torch.stack([x, x], 1)[:, 0]          # the reference takes sin / cos of strided views (angle[:, k])
    s, c = torch.sin(xs), torch.cos(xs)
    assert torch.equal(s, torch.sin(x)) and torch.equal(c, torch.cos(x))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="trig", x=_np(x), sin=_np(s), cos=_np(c))


def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def case_config1(mods, name, B, H, W, seed, stride=97):
    """BASELINE configs[0]: inverse_warp + photometric_reconstruction_loss (loss_functions.py:7-20) on a B x 3 x 128 x 416
    stereo + temporal batch.  Inputs are regenerated from the seed by dvf_b200.synthetic (their SHA-256 is stored so that a
    generator drift is detected); outputs that must match bit for bit are stored as SHA-256 plus packed mask bits, gradients
    as strided samples (every `stride`-th element) plus their fp64 sums."""
    d = syn.stereo_temporal_batch(B, H, W, seed=seed)
    t = {k: v.clone() for k, v in d.items()}
    for k in ("depth", "T_2to1", "T_R2L"):
        t[k].requires_grad_(True)
    _reset(mods)
    iw = mods["inverse_warp"]
    out = dict(kind="config1", B=B, H=H, W=W, seed=seed, stride=stride,
               inputs_sha=np.array([_sha(_np(d[k])) for k in sorted(d)]), input_names=np.array(sorted(d)))
    for tag, src, pose in (("R1", "img_R1", "T_2to1"), ("L2", "img_L2", "T_R2L")):
        w = iw.inverse_warp(d[src], d["depth"], d[pose], d["intrinsics"], d["intrinsics_inv"])
        valid = (1 - (w == 0).prod(1)).to(torch.uint8)
        out["warped_sha_" + tag] = _sha(_np(w))
        out["warped_sample_" + tag] = _np(w).reshape(-1)[::stride].copy()
        out["valid_bits_" + tag] = np.packbits(_np(valid).reshape(-1))
        out["P_" + tag] = _np(d["intrinsics"] @ iw.pose_vec2mat(d[pose]))
    _reset(mods)
    loss = mods["loss_functions"].photometric_reconstruction_loss(t["img_R2"], t["img_R1"], t["img_L2"], t["depth"], t["T_2to1"],
                                                                  t["T_R2L"], t["intrinsics"], t["intrinsics_inv"])
    loss.backward()
    gd = _np(t["depth"].grad)
    out.update(loss=_np(loss), g_T_2to1=_np(t["T_2to1"].grad), g_T_R2L=_np(t["T_R2L"].grad), g_depth_sha=_sha(gd),
               g_depth_sample=gd.reshape(-1)[::stride].copy(), g_depth_sum=np.float64(gd.astype(np.float64).sum()),
               g_depth_abs_sum=np.float64(np.abs(gd.astype(np.float64)).sum()), g_depth_max=np.float32(np.abs(gd).max()))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(1)
    os.makedirs(OUT, exist_ok=True)
    mods = _ref_modules()
    case_inverse_warp(mods, "iw_large_euler_zeros", 2, 3, 16, 52, "large", "euler", "zeros", seed=1)
    case_inverse_warp(mods, "iw_kitti_euler_zeros", 2, 3, 32, 104, "kitti", "euler", "zeros", seed=2)
    case_inverse_warp(mods, "iw_stereo_euler_zeros", 2, 3, 32, 104, "stereo", "euler", "zeros", seed=3)
    case_inverse_warp(mods, "iw_large_quat_border", 2, 3, 16, 52, "large", "quat", "border", seed=4)
    case_inverse_warp(mods, "iw_tiny_noise_euler_zeros", 1, 3, 24, 80, "tiny", "euler", "zeros", seed=5, smooth=False)
    case_inverse_warp(mods, "iw_feat8_kitti", 1, 8, 16, 52, "kitti", "euler", "zeros", seed=6, feature=True)
    case_loss_functions(mods, "lf_images", 2, 3, 32, 104, "kitti", seed=7)
    case_loss_functions(mods, "lf_features8", 2, 8, 16, 52, "kitti", seed=8, feature=True)
    case_sfm(mods, "sfm_3scales_mask", 2, 32, 104, 3, seed=9, with_mask=True, rot="euler", pad="zeros")
    case_sfm(mods, "sfm_2scales_nomask_quat_border", 2, 32, 104, 2, seed=10, with_mask=False, rot="quat", pad="border")
    case_sfm(mods, "sfm_old_2scales_mask", 2, 32, 104, 2, seed=11, with_mask=True, rot="euler", pad="zeros", old=True)
    case_se3("se3_exp", 6, seed=12)
    case_regularisers(mods, "regularisers", 2, 30, 52, 3, seed=13)
    case_trig("trig_f32", 1 << 17, seed=14)
    case_config1(mods, "config1_4x3x128x416", 4, 128, 416, seed=15)
    with _align_corners_true():
        case_inverse_warp(mods, "iw_kitti_euler_zeros_align", 2, 3, 32, 104, "kitti", "euler", "zeros", seed=16)
        case_inverse_warp(mods, "iw_large_quat_border_align", 2, 3, 16, 52, "large", "quat", "border", seed=17)
        case_loss_functions(mods, "lf_images_align", 2, 3, 32, 104, "kitti", seed=18)
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("wrote", sorted(os.listdir(OUT)), f"{tot / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
