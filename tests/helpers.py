"""Shared helpers of the test-suite: golden fixtures, tolerance metrics, synthetic cases."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerances stated by BASELINE.json north_star: fp32 results within 1e-5 relative, masks bit-exact.
RTOL_F32 = 1e-5


def golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: (z[k].item() if z[k].ndim == 0 and z[k].dtype.kind in "US" else z[k]) for k in z.files}


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def rel_err(a, b):
    """max|a-b| / max|b|  (the survey's acceptance metric, Appendix B.5)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / den


def assert_close(a, b, tol=RTOL_F32, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: max-abs error / max|ref| = {e:.3e} > {tol:.1e}"


def ulp_diff(a, b):
    """largest distance in float32 ulps between two arrays (treats +0/-0 as equal)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return int(np.abs(ia - ib).max()) if a.size else 0


class align_corners:
    """`with align_corners(ops, name)`: fixtures named *_align come from the reference run with
    F.grid_sample(align_corners=True) (oracle/gen_golden.py); the binding's module switch selects that convention."""

    def __init__(self, ops, name):
        self.ops, self.on = ops, str(name).endswith("_align")

    def __enter__(self):
        self.prev = self.ops.ALIGN_CORNERS
        self.ops.ALIGN_CORNERS = self.on
        return self.on

    def __exit__(self, *exc):
        self.ops.ALIGN_CORNERS = self.prev


def oracle_pad(name, pad="zeros"):
    return pad + ("_align" if str(name).endswith("_align") else "")


def torch_ssim_loss(x, y, valid=None):
    """The SSIM term of csrc/dvf_ssim.cu written with torch operators (avg_pool2d, autograd): the independent fp32/fp64
    restatement the kernel and the oracle are compared with (there is no SSIM in the reference)."""
    import torch
    import torch.nn.functional as F
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    mx, my = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    vx = F.avg_pool2d(x * x, 3, 1) - mx * mx
    vy = F.avg_pool2d(y * y, 3, 1) - my * my
    cxy = F.avg_pool2d(x * y, 3, 1) - mx * my
    S = ((2 * mx * my + C1) * (2 * cxy + C2)) / ((mx * mx + my * my + C1) * (vx + vy + C2))
    l = ((1 - S) / 2).clamp(0, 1)
    if valid is not None:
        m = (F.avg_pool2d(valid.to(l.dtype).unsqueeze(1), 3, 1) > 1 - 1e-6).to(l.dtype)
        l = l * m
    return l.mean()
