"""CPU: pin the oracle (oracle/dvf_oracle.c) against the golden vectors that oracle/gen_golden.py
produced by running the unmodified reference on torch-CPU.  Forward coordinate chain, warped image
and validity masks must be bit-identical given the reference's P; gradients within 1e-5."""
import numpy as np
import pytest

from helpers import assert_close, golden, golden_names, rel_err, ulp_diff


@pytest.mark.parametrize("name", golden_names("iw_"))
def test_inverse_warp_golden(oracle, name):
    g = golden(name)
    rot, pad = g["rotation_mode"], g["padding_mode"]
    # pose -> P: sin/cos of glibc vs torch's SLEEF may differ in the last place
    pm = oracle.pose_vec2mat(g["pose"], rot)
    assert ulp_diff(pm, g["posemat"]) <= 2
    P = oracle.project(g["K"], g["posemat"])
    assert np.array_equal(P, g["P"]), "K @ pose_mat must reproduce torch's tiny-bmm rounding"
    # per-pixel path on the reference's own P: bit-exact
    warped, valid = oracle.inverse_warp_P(g["img"], g["depth"], g["P"], g["Kinv"], pad)
    assert np.array_equal(warped, g["warped"])
    assert np.array_equal(valid, g["valid"])
    gimg, gdepth, gP = oracle.inverse_warp_bwd_P(g["gout"], g["img"], g["depth"], g["P"], g["Kinv"], pad)
    assert_close(gdepth, g["gdepth"], what="gdepth")
    assert_close(gimg, g["gimg"], what="gimg")
    gpose = oracle.pose_bwd(gP, g["K"], g["pose"], rot)
    assert_close(gpose, g["gpose"], what="gpose")


@pytest.mark.parametrize("name", golden_names("lf_"))
def test_loss_functions_golden(oracle, name):
    g = golden(name)
    feat = "g_img_R1" in g
    r = oracle.photo_loss_P(g["img_R2"], [g["img_R1"], g["img_L2"]], g["depth"], g["P"], g["intrinsics_inv"],
                            need_gsrc=feat, need_gtgt=feat)
    assert abs(r["terms"].sum() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert_close(r["gdepth"], g["g_depth"], what="gdepth")
    assert_close(oracle.pose_bwd(r["gP"][:, 0], g["intrinsics"], g["T_2to1"]), g["g_T_2to1"], what="g_T_2to1")
    assert_close(oracle.pose_bwd(r["gP"][:, 1], g["intrinsics"], g["T_R2L"]), g["g_T_R2L"], what="g_T_R2L")
    if feat:
        assert_close(r["gtgt"], g["g_img_R2"], what="g_img_R2")
        assert_close(r["gsrc"][0], g["g_img_R1"], what="g_img_R1")
        assert_close(r["gsrc"][1], g["g_img_L2"], what="g_img_L2")


@pytest.mark.parametrize("name", golden_names("sfm_"))
def test_sfm_golden(oracle, name):
    g = golden(name)
    old = g["kind"] == "sfm_old"
    rot, pad, n = g["rotation_mode"], g["padding_mode"], int(g["n_scales"])
    with_mask = bool(g["with_mask"])
    H = g["img_R2"].shape[2]
    total = 0.0
    V = 1 if old else 2
    gpose = np.zeros((g["pose"].shape[0], 2, 6), np.float64)
    for s in range(n):
        depth = g[f"depth{s}"][:, 0]
        h, w = depth.shape[1:]
        Ks, Kinvs = oracle.scale_intrinsics(g["intrinsics"], g["intrinsics_inv"], H / h)
        tgt = oracle.area_downsample(g["img_R2"], h, w)
        srcs = [oracle.area_downsample(g[k], h, w) for k in (["img_R1"] if old else ["img_R1", "img_L2"])]
        P = g[f"P{s}"][:, :V]
        expl = g[f"mask{s}"][:, :V] if with_mask else None
        r = oracle.photo_loss_P(tgt, srcs, depth, np.ascontiguousarray(P), Kinvs, expl=expl, padding_mode=pad)
        total += r["terms"].sum()
        assert_close(r["gdepth"], g[f"g_depth{s}"][:, 0], what=f"gdepth{s}")
        if with_mask:
            assert_close(r["gexpl"], g[f"g_mask{s}"][:, :V], what=f"gmask{s}")
            if old:
                assert not g[f"g_mask{s}"][:, 1:].any()
        for v in range(V):
            gpose[:, v] += oracle.pose_bwd(r["gP"][:, v], Ks, g["pose"][:, v], rot)
    assert abs(total - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert_close(gpose, g["g_pose"], what="gpose")


def test_smooth_and_explainability_match_closed_form(oracle):
    rng = np.random.default_rng(0)
    d = rng.random((2, 9, 11), dtype=np.float32)
    val, grad = oracle.smooth_loss_one(d, need_grad=True)
    dx = d[:, :, 1:] - d[:, :, :-1]
    dy = d[:, 1:] - d[:, :-1]
    ref = (np.abs(dx[:, :, 1:] - dx[:, :, :-1]).mean() + np.abs(dx[:, 1:] - dx[:, :-1]).mean()
           + np.abs(dy[:, :, 1:] - dy[:, :, :-1]).mean() + np.abs(dy[:, 1:] - dy[:, :-1]).mean())
    assert abs(val - ref) < 1e-6
    eps = 1e-3
    k = (1, 4, 5)
    dp = d.copy(); dp[k] += eps
    dm = d.copy(); dm[k] -= eps
    fd = (oracle.smooth_loss_one(dp) - oracle.smooth_loss_one(dm)) / (2 * eps)
    assert abs(fd - grad[k]) < 5e-3
    m = rng.random((2, 2, 5, 7), dtype=np.float32) * 0.98 + 0.01
    assert abs(oracle.explainability_loss_one(m) - float(-np.log(m.astype(np.float64)).mean())) < 1e-6


def test_se3_exp_golden(oracle):
    """SE(3) exponential map (se3_generate.py) vs the reference run with Tensor.cuda patched to identity."""
    g = golden("se3_exp")
    out = oracle.se3_exp(g["vec"])
    assert np.abs(out - g["out"][:, 0]).max() < 1e-12
    gin = oracle.se3_exp_bwd(g["vec"], g["gout"])
    assert_close(gin, g["gvec"].reshape(-1, 6), tol=1e-6, what="se3 grad")


def test_regularisers_match_reference_golden(oracle):
    """oracle smooth / explainability terms against the reference's own values and autograd gradients
    (tests/golden/regularisers.npz, generated by oracle/gen_golden.py from loss_functions.py:23-41 and
    loss_functions_sfm.py:49-56)"""
    from helpers import assert_close, golden
    g = golden("regularisers")
    ref, w = 0.0, 1.0
    for i in range(int(g["n_maps"])):
        v, gr = oracle.smooth_loss_one(g[f"map{i}"][:, 0], need_grad=True)
        ref += v * w
        assert_close(w * gr, g[f"g_map{i}"][:, 0], what=f"smooth grad {i}")
        w /= 2.0
    assert abs(ref - float(g["smooth"])) <= 1e-5 * float(g["smooth"])
    ref = 0.0
    for i in range(int(g["n_masks"])):
        v, gr = oracle.explainability_loss_one(g[f"mask{i}"], need_grad=True)
        ref += v
        assert_close(gr, g[f"g_mask{i}"], what=f"expl grad {i}")
    assert abs(ref - float(g["expl"])) <= 1e-5 * float(g["expl"])
