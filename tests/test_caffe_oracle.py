"""CPU checks of the Caffe-convention oracle block (oracle/dvf_oracle.c, "PARITY UNPINNED": BVLC Caffe cannot be
built here, so there are no reference outputs for these layers).  What can be checked without the reference: the
analytic backward of every layer against central finite differences of its own forward (in fp64-accumulated
losses), closed-form cases, and the sign(0) = -1 convention of AbsLoss (abs_loss_layer.cu:28)."""
import numpy as np
import pytest


def _case(N=2, H=10, W=14, C=2, seed=0):
    rng = np.random.default_rng(seed)
    depth = rng.uniform(3.0, 20.0, (N, H, W)).astype(np.float32)
    T = np.tile(np.eye(4, dtype=np.float32), (N, 1, 1))
    T[:, :3, :3] += rng.standard_normal((N, 3, 3)).astype(np.float32) * 0.01
    T[:, :3, 3] = rng.standard_normal((N, 3)).astype(np.float32) * 0.2
    K = np.tile(np.array([[0.6 * W, 1.8 * H, 0.5 * W, 0.5 * H]], np.float32), (N, 1))
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    img = np.stack([np.sin(0.3 * xx + 0.2 * yy + c) for c in range(C)])[None].repeat(N, 0).astype(np.float32)   # smooth
    return depth, T, K, img, rng


def _fd(f, x, idx, eps):
    xp, xm = x.copy(), x.copy()
    xp[idx] += eps
    xm[idx] -= eps
    return (f(xp) - f(xm)) / (float(xp[idx]) - float(xm[idx]))


def test_geo_backward_matches_finite_differences(oracle):
    depth, T, K, _, rng = _case()
    top = rng.standard_normal((2, 3, 10, 14)).astype(np.float32)
    dd, dT, dK = oracle.caffe_geo_bwd(top, depth, T, K)
    f = lambda d, t, k: float((oracle.caffe_geo_fwd(d, t, k).astype(np.float64) * top).sum())
    for idx in [(0, 0, 0), (1, 4, 7), (0, 9, 13)]:
        assert abs(_fd(lambda d: f(d, T, K), depth, idx, 1e-2) - dd[idx]) < 2e-3 * max(1, abs(dd[idx]))
    for idx in [(0, 0, 0), (1, 1, 2), (0, 2, 3), (1, 0, 1)]:
        ref = _fd(lambda t: f(depth, t, K), T, idx, 1e-3)
        assert abs(ref - dT.reshape(2, 4, 4)[idx]) < 2e-3 * max(1, abs(ref))
    assert np.all(dT.reshape(2, 4, 4)[:, 3] == 0)
    for idx in [(0, 0), (1, 1), (0, 2), (1, 3)]:
        ref = _fd(lambda k: f(depth, T, k), K, idx, 1e-2)
        assert abs(ref - dK[idx]) < 5e-3 * max(1, abs(ref))


def test_pinhole_backward_matches_finite_differences(oracle):
    depth, T, K, _, rng = _case(seed=1)
    pts = oracle.caffe_geo_fwd(depth, T, K)
    top = rng.standard_normal((2, 2, 10, 14)).astype(np.float32)
    dp, dK = oracle.caffe_pinhole_bwd(top, pts, K)
    f = lambda p, k: float((oracle.caffe_pinhole_fwd(p, k).astype(np.float64) * top).sum())
    for idx in [(0, 0, 0, 0), (1, 1, 4, 7), (0, 2, 9, 13), (1, 2, 5, 5)]:
        ref = _fd(lambda p: f(p, K), pts, idx, 1e-2)
        assert abs(ref - dp[idx]) < 5e-3 * max(1, abs(ref))
    for idx in [(0, 0), (1, 1), (0, 2), (1, 3)]:
        ref = _fd(lambda k: f(pts, k), K, idx, 1e-2)
        assert abs(ref - dK[idx]) < 5e-3 * max(1, abs(ref))


def test_warp_backward_matches_finite_differences(oracle):
    depth, T, K, img, rng = _case(seed=2)
    N, C, H, W = img.shape
    xy = np.stack([rng.uniform(0.2, W - 1.2, (N, H, W)), rng.uniform(0.2, H - 1.2, (N, H, W))], 1).astype(np.float32)
    xy = np.floor(xy) + 0.25 + 0.5 * rng.uniform(0, 1, xy.shape).astype(np.float32)     # keep eps inside one cell
    top = rng.standard_normal(img.shape).astype(np.float32)
    gi, gxy = oracle.caffe_warp_bwd(top, img, xy)
    f = lambda u, c: float((oracle.caffe_warp_fwd(u, c).astype(np.float64) * top).sum())
    for idx in [(0, 0, 0, 0), (1, 1, 4, 7), (0, 1, 9, 13)]:
        ref = _fd(lambda c: f(img, c), xy, idx, 1e-2)
        assert abs(ref - gxy[idx]) < 2e-3 * max(1, abs(ref))
    for idx in [(0, 0, 3, 3), (1, 1, 4, 7)]:
        ref = _fd(lambda u: f(u, xy), img, idx, 1e-2)
        assert abs(ref - gi[idx]) < 2e-3 * max(1, abs(ref))


def test_closed_forms(oracle):
    depth, T, K, img, _ = _case(seed=3)
    N, C, H, W = img.shape
    T[:] = np.eye(4, dtype=np.float32)
    xy = oracle.caffe_pinhole_fwd(oracle.caffe_geo_fwd(depth, T, K), K)
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    assert np.abs(xy[:, 0] - xs).max() < 1e-4 and np.abs(xy[:, 1] - ys).max() < 1e-4
    exact = np.broadcast_to(np.stack([xs, ys])[None], (N, 2, H, W)).copy()
    assert np.array_equal(oracle.caffe_warp_fwd(img, exact), img)
    shifted = exact.copy()
    shifted[:, 0] += 1.0                      # integer shift: the last column samples outside -> zero
    out = oracle.caffe_warp_fwd(img, shifted)
    assert np.array_equal(out[..., :-1], img[..., 1:]) and np.all(out[..., -1] == 0)
    val, ga, gb = oracle.caffe_abs_loss(img, img, weight=1.0)
    assert val == 0.0 and np.all(ga == np.float32(-1.0 / N)) and np.all(gb == np.float32(1.0 / N))
    val, ga, _ = oracle.caffe_abs_loss(img + 1.0, img)
    assert abs(val - img.size / N) < 1e-3 and np.all(ga == np.float32(1.0 / N))


def test_edge_aware_smoothness_oracle(oracle):
    """Caffe edge-aware smoothness (experiments/depth/train.prototxt:4022-4234), PARITY UNPINNED: closed forms and central
    finite differences of the oracle's own forward."""
    rng = np.random.default_rng(5)
    N, H, W = 2, 9, 11
    img = rng.random((N, 3, H, W), dtype=np.float32)
    inv = (rng.random((N, 1, H, W), dtype=np.float32) * 0.3 + 0.02).astype(np.float32)
    loss, g = oracle.caffe_edge_smooth(img, inv, weight=10.0)
    # flat image -> gx = gy = 1: plain sums of |central differences| over the valid windows
    l0, _ = oracle.caffe_edge_smooth(np.full_like(img, 0.5), inv, need_grad=False)
    D = inv[:, 0].astype(np.float64)
    dx = 0.5 * (D[:, 2:, 1:-1] - D[:, :-2, 1:-1])
    dy = 0.5 * (D[:, 1:-1, 2:] - D[:, 1:-1, :-2])
    assert abs(l0[0] - np.abs(dx).sum() / N) < 1e-6 and abs(l0[1] - np.abs(dy).sum() / N) < 1e-6
    # ramp in x only: EdgeX (vertical difference) vanishes, EdgeY is the slope
    ramp = np.tile(np.arange(W, dtype=np.float32) * 0.01, (N, 1, H, 1))
    lr, _ = oracle.caffe_edge_smooth(np.full_like(img, 0.5), ramp, need_grad=False)
    assert lr[0] == 0.0 and abs(lr[1] - (H - 2) * (W - 2) * 0.01) < 1e-6
    # finite differences (fp64 restatement of the forward)
    def f64(inv64):
        I = img.astype(np.float64)
        sx = np.abs(0.5 * (I[:, :, 2:, 1:-1] - I[:, :, :-2, 1:-1])).sum(1)
        sy = np.abs(0.5 * (I[:, :, 1:-1, 2:] - I[:, :, 1:-1, :-2])).sum(1)
        d = inv64[:, 0]
        ex = np.exp(-0.33 * sx) * 0.5 * (d[:, 2:, 1:-1] - d[:, :-2, 1:-1])
        ey = np.exp(-0.33 * sy) * 0.5 * (d[:, 1:-1, 2:] - d[:, 1:-1, :-2])
        return 10.0 * (np.abs(ex).sum() + np.abs(ey).sum()) / N
    assert abs(f64(inv.astype(np.float64)) - 10.0 * loss.sum()) < 1e-5 * 10.0 * loss.sum()
    eps = 1e-6
    for (n, r, c) in [(0, 0, 1), (0, 4, 5), (1, 8, 9), (1, 3, 0), (0, 2, 10)]:
        a = inv.astype(np.float64); b = a.copy()
        a[n, 0, r, c] += eps; b[n, 0, r, c] -= eps
        fd = (f64(a) - f64(b)) / (2 * eps)
        assert abs(fd - g[n, 0, r, c]) < 1e-4 * max(1.0, abs(fd)), (n, r, c, fd, g[n, 0, r, c])
