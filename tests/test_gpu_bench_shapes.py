"""GPU parity at the BENCHMARKED shapes (BASELINE.json configs): the CUDA path against the CPU oracle on the same
seeded inputs at C2 (B=64, 4 scales, stereo), a C3 shard (B=32, V=2, explainability masks, 4 scales, PoseExpNet-scale
poses) and C4 (B=128, 64-channel maps at 32x104, fp32 and bf16, gradients to every map), and against the golden of the
reference itself at config 1 (4 x 3 x 128 x 416).  Same tolerances as everywhere: 1e-5 relative in fp32, 1e-2 for
bf16 map gradients, masks / warped images / projection matrices bit-exact."""
import hashlib

import numpy as np
import pytest
import torch

from helpers import RTOL_F32, assert_close, golden

pytestmark = pytest.mark.gpu

H, W, L = 128, 416, 4


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def ops():
    from dvf_b200 import ops as _ops, _lib
    _lib.load()
    return _ops


@pytest.fixture(scope="module")
def syn():
    from dvf_b200 import synthetic
    return synthetic


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_config1_golden_through_the_dropin(ops, syn):
    """configs[0]: inverse_warp + photometric_reconstruction_loss on 4 x 3 x 128 x 416 through the drop-in modules against
    the reference's own outputs: warped images (SHA-256) and masks bit for bit, loss and gradients at 1e-5."""
    import inverse_warp as iw
    import loss_functions as lf
    g = golden("config1_4x3x128x416")
    B = int(g["B"])
    d = syn.stereo_temporal_batch(B, int(g["H"]), int(g["W"]), seed=int(g["seed"]))
    assert [_sha(d[k].numpy()) for k in sorted(d)] == list(g["inputs_sha"]), "synthetic generator drifted: regenerate the goldens"
    t = {k: v.cuda() for k, v in d.items()}
    for tag, src, pose in (("R1", "img_R1", "T_2to1"), ("L2", "img_L2", "T_R2L")):
        w = iw.inverse_warp(t[src], t["depth"], t[pose], t["intrinsics"], t["intrinsics_inv"])
        assert _sha(npy(w)) == g["warped_sha_" + tag], f"{tag}: warped image differs from the reference"
        valid = (w != 0).any(1).to(torch.uint8)
        assert np.array_equal(np.packbits(npy(valid).reshape(-1)), g["valid_bits_" + tag]), f"{tag}: validity mask"
        _, P, _ = ops.pose_proj_fwd(t[pose], t["intrinsics"], None, 1, "euler", [1.0])
        assert np.array_equal(npy(P[0]), g["P_" + tag])
    for k in ("depth", "T_2to1", "T_R2L"):
        t[k].requires_grad_(True)
    loss = lf.photometric_reconstruction_loss(t["img_R2"], t["img_R1"], t["img_L2"], t["depth"], t["T_2to1"], t["T_R2L"],
                                              t["intrinsics"], t["intrinsics_inv"])
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= RTOL_F32 * float(g["loss"])
    gd = npy(t["depth"].grad)
    st = int(g["stride"])
    assert_close(gd.reshape(-1)[::st], g["g_depth_sample"], what="g_depth sample")
    assert abs(np.abs(gd.astype(np.float64)).sum() - float(g["g_depth_abs_sum"])) <= 1e-6 * float(g["g_depth_abs_sum"])
    assert abs(float(np.abs(gd).max()) - float(g["g_depth_max"])) <= 1e-5 * float(g["g_depth_max"])
    assert_close(npy(t["T_2to1"].grad), g["g_T_2to1"], what="g_T_2to1")
    assert_close(npy(t["T_R2L"].grad), g["g_T_R2L"], what="g_T_R2L")


def _multi_scale_vs_oracle(ops, oracle, syn, B, V, with_expl, kinds, seed):
    """4-scale image loss through the drop-in (loss_functions_sfm, pyramid included) against the oracle level by level."""
    import loss_functions_sfm as sfm
    sizes = [(H >> s, W >> s) for s in range(L)]
    ds = [float(1 << s) for s in range(L)]
    imgs = syn.images(B, 3, H, W, seed + 1, n=V + 1)
    tgt, srcs = imgs[0], imgs[1:]
    depths = [syn.depth(B, h, w, seed + 10 + i) for i, (h, w) in enumerate(sizes)]
    expl = [syn.explainability(B, V, h, w, seed + 20 + i) for i, (h, w) in enumerate(sizes)] if with_expl else None
    pose = torch.stack([syn.pose(B, kinds[v], seed + 30 + v) for v in range(V)], 1)
    K, Kinv = syn.intrinsics(B, H, W)

    t_depths = [x.unsqueeze(1).cuda().requires_grad_(True) for x in depths]
    t_expl = [e.cuda().requires_grad_(True) for e in expl] if with_expl else [None] * L
    t_pose = pose.cuda().requires_grad_(True)
    loss = sfm.photometric_reconstruction_loss(tgt.cuda(), [s.cuda() for s in srcs], K.cuda(), Kinv.cuda(), t_depths, t_expl, t_pose)
    loss.backward()

    total, gpose = 0.0, np.zeros((B, V, 6))
    for s, (h, w) in enumerate(sizes):
        oK, oKi = oracle.scale_intrinsics(K.numpy(), Kinv.numpy(), ds[s])
        Pn = np.stack([oracle.project(oK, oracle.pose_vec2mat(pose[:, v].numpy())) for v in range(V)], 1)
        o_tgt = oracle.area_downsample(tgt.numpy(), h, w)
        o_srcs = [oracle.area_downsample(x.numpy(), h, w) for x in srcs]
        r = oracle.photo_loss_P(o_tgt, o_srcs, depths[s].numpy(), Pn, oKi, expl=None if expl is None else expl[s].numpy())
        total += float(r["terms"].sum())
        assert_close(npy(t_depths[s].grad)[:, 0], r["gdepth"], what=f"gdepth level {s}")
        if with_expl:
            assert_close(npy(t_expl[s].grad), r["gexpl"], what=f"gexpl level {s}")
        for v in range(V):
            gpose[:, v] += oracle.pose_bwd(r["gP"][:, v], oK, pose[:, v].numpy())
    assert abs(loss.item() - total) <= RTOL_F32 * total
    assert_close(npy(t_pose.grad), gpose, what="gpose")


def test_c2_shape_vs_oracle(ops, oracle, syn):
    """configs[1] (the benchmarked workload): B=64, 4 scales, one stereo view, fp32"""
    _multi_scale_vs_oracle(ops, oracle, syn, 64, 1, False, ["stereo"], seed=500)


def test_c3_shard_vs_oracle(ops, oracle, syn):
    """configs[2], one rank's shard at 8 GPUs: B=32, V=2 (temporal pose at PoseExpNet-at-init scale + stereo),
    explainability masks, 4 scales"""
    _multi_scale_vs_oracle(ops, oracle, syn, 32, 2, True, ["tiny", "stereo"], seed=600)


@pytest.mark.parametrize("dtype,layout", [("f32", "nhwc"), ("bf16", "nhwc"), ("f32", "nchw")])
def test_c4_shape_vs_oracle(ops, oracle, syn, dtype, layout):
    """configs[3]: feature reconstruction loss on 64-channel maps at 32x104, B=128, V=2, gradients to all three maps;
    channels-last fp32 / bf16 and the reference's dense NCHW fp32 (what FeatExtractor hands over)"""
    import loss_functions as lf
    B, C, h, w, V = 128, 64, 32, 104, 2
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    maps = [m.to(tdt) for m in syn.features(B, C, h, w, 700, n=3)]
    ref = [m.float().numpy() for m in maps]
    depth = syn.depth(B, h, w, 701)
    T_2to1, T_R2L = syn.pose(B, "tiny", 702), syn.pose(B, "stereo", 703)
    K, Kinv = syn.intrinsics(B, h, w)
    put = (lambda t: t.cuda().contiguous(memory_format=torch.channels_last)) if layout == "nhwc" else (lambda t: t.cuda())
    t_maps = [put(m).requires_grad_(True) for m in maps]
    t_depth = depth.cuda().requires_grad_(True)
    t1, t2 = T_2to1.cuda().requires_grad_(True), T_R2L.cuda().requires_grad_(True)
    loss = lf.photometric_reconstruction_loss(t_maps[0], t_maps[1], t_maps[2], t_depth, t1, t2, K.cuda(), Kinv.cuda())
    loss.backward()
    Pn = np.stack([oracle.project(K.numpy(), oracle.pose_vec2mat(p.numpy())) for p in (T_2to1, T_R2L)], 1)
    r = oracle.photo_loss_P(ref[0], ref[1:], depth.numpy(), Pn, Kinv.numpy(), need_gsrc=True, need_gtgt=True)
    tol_geo = RTOL_F32 if dtype == "f32" else 1e-4
    tol_map = RTOL_F32 if dtype == "f32" else 1e-2
    assert abs(loss.item() - float(r["terms"].sum())) <= tol_geo * float(r["terms"].sum())
    assert_close(npy(t_depth.grad), r["gdepth"], tol=tol_geo, what="gdepth")
    assert t_maps[0].grad.shape == t_maps[0].shape and t_maps[0].grad.dtype == tdt
    assert_close(npy(t_maps[0].grad.float()), r["gtgt"], tol=tol_map, what="g target map")
    assert_close(npy(t_maps[1].grad.float()), r["gsrc"][0], tol=tol_map, what="g temporal source map")
    assert_close(npy(t_maps[2].grad.float()), r["gsrc"][1], tol=tol_map, what="g stereo source map")
    assert_close(npy(t1.grad), oracle.pose_bwd(r["gP"][:, 0], K.numpy(), T_2to1.numpy()), tol=tol_geo, what="g T_2to1")
    assert_close(npy(t2.grad), oracle.pose_bwd(r["gP"][:, 1], K.numpy(), T_R2L.numpy()), tol=tol_geo, what="g T_R2L")
