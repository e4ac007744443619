"""GPU parity tests: the CUDA path (through the C ABI of libdvf_b200.so) against the CPU oracle on the
same seeded inputs, and against the golden vectors of the real reference.

Tolerances (BASELINE.json north_star): fp32 warped images, loss and gradients within 1e-5 relative
(max|a-b| / max|ref|); validity masks bit-exact.  The pose chain reproduces torch-CPU's sin / cos bit for bit
(csrc/dvf_pose.cuh), so the projection matrix P, the sampling positions, the warped images and the masks are
compared for EXACT equality through the public signatures -- there is no looser branch anywhere in this file.
"""
import numpy as np
import pytest
import torch

from helpers import RTOL_F32, align_corners, assert_close, golden, golden_names, oracle_pad, rel_err, ulp_diff

pytestmark = pytest.mark.gpu


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


# ------------------------------------------------------------------------------------------------
def test_fast_division_matches_ieee(ops):
    """the shared-reciprocal division used by the coordinate chain == __fdiv_rn on 2^30 operand pairs"""
    from dvf_b200 import _lib
    lib = _lib.load()
    for mode in (0, 1):
        bad = torch.zeros(1, dtype=torch.int64, device="cuda")
        _lib.check(lib.dvf_selftest_fast_div(1234 + mode, 1 << 30, mode, bad.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "selftest")
        assert int(bad.item()) == 0, f"mode {mode}: {int(bad.item())} quotients differ from __fdiv_rn"


def _gpu_sincos(x):
    from dvf_b200 import _lib
    lib = _lib.load()
    xt = cu(x)
    s, c = torch.empty_like(xt), torch.empty_like(xt)
    _lib.check(lib.dvf_torch_sincos(xt.data_ptr(), xt.numel(), s.data_ptr(), c.data_ptr(), torch.cuda.current_stream().cuda_stream),
               "dvf_torch_sincos")
    return npy(s), npy(c)


def test_trig_golden_bit_exact(ops, oracle):
    """euler2mat's torch.sin / torch.cos (inverse_warp.py:89-106): the device routine reproduces torch-CPU bit for bit on
    the 131k golden angles (tiny / KITTI-like / unit / +-pi / +-100 / +-10000 / special values)."""
    g = golden("trig_f32")
    s, c = _gpu_sincos(g["x"])
    assert np.array_equal(s.view(np.uint32), g["sin"].view(np.uint32))
    assert np.array_equal(c.view(np.uint32), g["cos"].view(np.uint32))
    os_, oc = oracle.torch_trig(g["x"])
    assert np.array_equal(os_.view(np.uint32), s.view(np.uint32)) and np.array_equal(oc.view(np.uint32), c.view(np.uint32))


def test_trig_matches_live_torch_cpu():
    """the same against torch.sin / torch.cos evaluated on THIS machine's CPU: 2^24 angles of every magnitude class.
    torch-CPU's fp32 sin / cos are MKL VML kernels that MKL picks by CPU model (oracle/torch_trig.h): the FMA kernels
    (AVX2 / AVX-512 on Intel parts, what generated tests/golden/trig_f32.npz) are the pinned target and must match bit
    for bit; MKL's other code paths (pre-FMA SSE / AVX, non-Intel hosts) round ~2e-5 of the angles differently, i.e.
    torch-CPU does not agree with itself across hosts there.  So: every value within 1 ulp and fewer than 1e-4 of them
    different (the count and whether this host runs the pinned kernels are printed; the bit-exact pin is the golden test)."""
    gen = torch.Generator().manual_seed(77)
    n = 1 << 21
    x = torch.cat([torch.randn(n, generator=gen) * sc for sc in (1e-3, 0.01, 0.1, 1.0, 10.0)] +
                  [(torch.rand(n, generator=gen) * 2 - 1) * r for r in (3.1415927, 1000.0, 10000.0)])
    # bit patterns around multiples of pi/2, where the reduction switches n
    k = torch.arange(-2000, 2001, dtype=torch.float64) * (np.pi / 2)
    near = torch.cat([torch.nextafter(k.float(), torch.tensor(float(sgn) * 1e9)) for sgn in (-1, 1)] + [k.float()])
    x = torch.cat([x, near]).contiguous()
    s, c = _gpu_sincos(x.numpy())
    g = golden("trig_f32")
    gx = torch.from_numpy(g["x"])
    host_is_pinned = (np.array_equal(torch.sin(gx).numpy().view(np.uint32), g["sin"].view(np.uint32)) and
                      np.array_equal(torch.cos(gx).numpy().view(np.uint32), g["cos"].view(np.uint32)))
    for name, mine, ref in (("sin", s, torch.sin(x).numpy()), ("cos", c, torch.cos(x).numpy())):
        diff = mine.view(np.uint32) != ref.view(np.uint32)
        nd = int(diff.sum())
        print(f"live torch-CPU {name}: {nd} of {x.numel()} differ; host runs the pinned MKL kernels: {host_is_pinned}; "
              f"cpu capability {torch.backends.cpu.get_cpu_capability()}")
        # bit equality is pinned by the golden vectors (test_trig_golden_bit_exact) and by oracle/check_torch_trig.log; the
        # LIVE comparison gets the tolerance torch-CPU needs against itself (different MKL code paths differ in the last
        # place on ~2e-5 of the angles), so that a host quirk cannot fail the suite.  On hosts that reproduce the golden
        # vectors nd has been 0 in every run so far (printed above).
        ulp = np.abs(mine.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
        assert nd < 1e-4 * x.numel() and int(ulp.max()) <= 1, (name, nd, int(ulp.max()), host_is_pinned)


@pytest.mark.parametrize("rot", ["euler", "quat"])
@pytest.mark.parametrize("kind", ["kitti", "tiny", "large", "stereo"])
def test_pose_proj_vs_oracle(ops, oracle, syn, rot, kind):
    B, V = 16, 2
    pose = syn.pose(B * V, kind, 3).numpy()
    K, Kinv = syn.intrinsics(B, 128, 416)
    ds = [1.0, 2.0, 4.0, 8.0]
    pm, P, Ks = ops.pose_proj_fwd(cu(pose), K.cuda(), Kinv.cuda(), V, rot, ds, want_posemat=True)
    opm = oracle.pose_vec2mat(pose, rot)
    assert np.array_equal(npy(pm), opm), "pose_vec2mat must be bit-identical (torch-CPU sin / cos order)"
    for l, d in enumerate(ds):
        oK, oKi = oracle.scale_intrinsics(K.numpy(), Kinv.numpy(), d)
        assert np.array_equal(npy(Ks[l]), oKi)
        # same posemat in => bit-identical P out
        oP = oracle.project(np.repeat(oK, V, axis=0), npy(pm))
        assert np.array_equal(npy(P[l]), oP)
    # backward (fp64 analytic on both sides)
    g = np.random.default_rng(0).standard_normal((len(ds), B * V, 3, 4)).astype(np.float32)
    gv = ops.pose_proj_bwd(cu(g), None, cu(pose), K.cuda(), V, rot, ds)
    ref = np.zeros((B * V, 6), np.float64)
    for l, d in enumerate(ds):
        oK, _ = oracle.scale_intrinsics(K.numpy(), Kinv.numpy(), d)
        ref += oracle.pose_bwd(g[l], np.repeat(oK, V, axis=0), pose, rot)
    assert_close(npy(gv), ref, what="gvec")


CASES = [  # B, C, H, W, pose kind, smooth, padding
    (4, 3, 128, 416, "kitti", True, "zeros"),
    (4, 3, 128, 416, "stereo", True, "zeros"),
    (2, 3, 128, 416, "tiny", False, "zeros"),
    (3, 3, 16, 52, "large", True, "zeros"),
    (2, 3, 64, 208, "kitti", True, "border"),
    (2, 5, 33, 71, "large", True, "zeros"),     # ragged sizes, odd channel count
    (1, 1, 7, 9, "kitti", True, "border"),
    (1, 32, 32, 104, "kitti", True, "zeros"),   # feature-map shape
]


def _case(syn, B, C, H, W, kind, smooth, seed=21):
    d = syn.stereo_temporal_batch(B, H, W, seed=seed, C=C, smooth=smooth, temporal=kind if kind != "stereo" else "kitti",
                                  feature=C not in (1, 3))
    pose = d["T_R2L"] if kind == "stereo" else d["T_2to1"]
    return d, pose


@pytest.mark.parametrize("case", CASES)
def test_inverse_warp_fwd_bwd_vs_oracle(ops, oracle, syn, case):
    B, C, H, W, kind, smooth, pad = case
    d, pose = _case(syn, B, C, H, W, kind, smooth)
    K, Kinv = d["intrinsics"], d["intrinsics_inv"]
    P = oracle.project(K.numpy(), oracle.pose_vec2mat(pose.numpy()))
    img, depth = d["img_R1"].numpy(), d["depth"].numpy()
    ow, ov = oracle.inverse_warp_P(img, depth, P, Kinv.numpy(), pad)
    w, v = ops.inverse_warp_fwd_P(cu(img), cu(depth), cu(P), Kinv.cuda(), pad, want_valid=True)
    assert np.array_equal(npy(v), ov), "validity mask must be bit-exact"
    assert np.array_equal(npy(w), ow), "same P => bit-identical warped image"
    gout = np.random.default_rng(1).standard_normal(img.shape).astype(np.float32)
    ogi, ogd, ogP = oracle.inverse_warp_bwd_P(gout, img, depth, P, Kinv.numpy(), pad)
    gi, gd, gP = ops.inverse_warp_bwd_P(cu(gout), cu(img), cu(depth), cu(P), Kinv.cuda(), pad)
    assert_close(npy(gd), ogd, what="gdepth")
    assert_close(npy(gi), ogi, what="gimg")
    assert_close(npy(gP), ogP, what="gP")
    # without d img, images take the route through the fused loss kernel (TMA ring, balanced split): the same chain
    _, gd2, gP2 = ops.inverse_warp_bwd_P(cu(gout), cu(img), cu(depth), cu(P), Kinv.cuda(), pad, need_gimg=False)
    assert np.array_equal(npy(gd2), npy(gd)), "d depth must not depend on the route"
    assert_close(npy(gP2), ogP, what="gP (no d img)")


def test_inverse_warp_bwd_route_and_workspace_reuse(ops, oracle, syn):
    """dvf_inverse_warp_bwd picks its kernel per call (d img requested or not, alignment): alternating calls on ONE cached
    workspace, a batch that is cut into many pieces, and an unaligned depth view that falls back to the plain kernel."""
    B, H, W = 16, 128, 416
    d, pose = _case(syn, B, 3, H, W, "kitti", True, seed=33)
    K, Kinv = d["intrinsics"], d["intrinsics_inv"]
    P = oracle.project(K.numpy(), oracle.pose_vec2mat(pose.numpy()))
    img, depth = d["img_R1"].numpy(), d["depth"].numpy()
    gout = np.random.default_rng(2).standard_normal(img.shape).astype(np.float32)
    _, ogd, ogP = oracle.inverse_warp_bwd_P(gout, img, depth, P, Kinv.numpy(), "zeros", need_gimg=False)
    args = (cu(gout), cu(img), cu(depth), cu(P), Kinv.cuda(), "zeros")
    for need in (False, True, False, False, True, False):
        gi, gd, gP = ops.inverse_warp_bwd_P(*args, need_gimg=need)
        assert (gi is not None) == need
        assert np.array_equal(npy(gd), ogd), "d depth is bit-exact on either route"
        assert_close(npy(gP), ogP, what="gP")
    # a depth tensor that starts 4 bytes off a 16-byte boundary: no bulk copies, plain kernel, same numbers
    buf = torch.empty(depth.size + 1, device="cuda")
    dv = buf[1:].view(depth.shape)
    dv.copy_(cu(depth))
    _, gd, gP = ops.inverse_warp_bwd_P(args[0], args[1], dv, args[3], args[4], "zeros", need_gimg=False)
    assert np.array_equal(npy(gd), ogd)
    assert_close(npy(gP), ogP, what="gP (unaligned)")


@pytest.mark.parametrize("V,with_expl,C,pad", [(1, False, 3, "zeros"), (2, False, 3, "zeros"), (2, True, 3, "zeros"),
                                               (3, True, 3, "border"), (4, False, 3, "zeros"), (2, False, 8, "zeros"),
                                               (2, True, 32, "zeros"), (1, True, 1, "zeros")])
def test_fused_loss_single_level_vs_oracle(ops, oracle, syn, V, with_expl, C, pad):
    B, H, W = 3, 48, 136
    feat = C not in (1, 3)
    imgs = syn.features(B, C, H, W, 5, n=V + 1) if feat else syn.images(B, C, H, W, 5, n=V + 1)
    tgt, srcs = imgs[0], imgs[1:]
    depth = syn.depth(B, H, W, 6)
    kinds = ["kitti", "stereo", "large", "tiny"]
    pose = torch.stack([syn.pose(B, kinds[v], 7 + v) for v in range(V)], 1)
    K, Kinv = syn.intrinsics(B, H, W)
    expl = syn.explainability(B, V, H, W, 9) if with_expl else None
    Pn = np.stack([oracle.project(K.numpy(), oracle.pose_vec2mat(pose[:, v].numpy())) for v in range(V)], 1)
    r = oracle.photo_loss_P(tgt.numpy(), [s.numpy() for s in srcs], depth.numpy(), Pn, Kinv.numpy(),
                            expl=None if expl is None else expl.numpy(), padding_mode=pad, need_gsrc=True, need_gtgt=True)

    # dense NCHW feature maps reach the channels-last kernel through a re-layout by default; the generic NCHW kernel is
    # exercised with the switch off
    for via_nhwc in ([True, False] if feat else [True]):
        ops.NCHW_FEATURES_VIA_NHWC = via_nhwc
        try:
            t_tgt = tgt.cuda().requires_grad_(True)
            t_srcs = [s.cuda().requires_grad_(True) for s in srcs]
            t_depth = depth.cuda().requires_grad_(True)
            t_pose = pose.cuda().requires_grad_(True)
            t_expl = None if expl is None else expl.cuda().requires_grad_(True)
            loss, terms = ops.fused_photo_loss([t_tgt], [t_srcs], [t_depth], t_pose, K.cuda(), Kinv.cuda(),
                                               expl_levels=None if expl is None else [t_expl], padding_mode=pad)
            loss.backward()
        finally:
            ops.NCHW_FEATURES_VIA_NHWC = True
        _, P_gpu, _ = ops.pose_proj_fwd(t_pose.detach().reshape(B * V, 6), K.cuda(), None, V, "euler", [1.0])
        assert np.array_equal(npy(P_gpu[0]).reshape(B, V, 3, 4), Pn), "pose -> P must be bit-identical"
        tol = RTOL_F32
        what = f" (via_nhwc={via_nhwc})"
        assert_close(npy(terms), r["terms"], tol=RTOL_F32, what="loss terms" + what)
        assert abs(loss.item() - r["terms"].sum()) <= RTOL_F32 * r["terms"].sum()
        assert_close(npy(t_depth.grad), r["gdepth"], tol=tol, what="gdepth" + what)
        assert t_tgt.grad.shape == t_tgt.shape
        assert_close(npy(t_tgt.grad), r["gtgt"], tol=tol, what="gtgt" + what)
        for v in range(V):
            assert_close(npy(t_srcs[v].grad), r["gsrc"][v], tol=tol, what=f"gsrc{v}" + what)
            gp = oracle.pose_bwd(r["gP"][:, v], K.numpy(), pose[:, v].numpy())
            assert_close(npy(t_pose.grad[:, v]), gp, tol=tol, what=f"gpose{v}" + what)
        if with_expl:
            assert_close(npy(t_expl.grad), r["gexpl"], tol=tol, what="gexpl" + what)


def test_fused_loss_exact_P_path_vs_oracle(ops, oracle, syn):
    """the C ABI takes P directly: with the oracle's P every output is held to 1e-5 and the
    per-image dP reduction to the fp64 reference."""
    import ctypes as C
    from dvf_b200 import _lib
    lib = _lib.load()
    B, Cc, H, W, V = 4, 3, 128, 416, 2
    d = syn.stereo_temporal_batch(B, H, W, seed=31)
    K, Kinv = d["intrinsics"], d["intrinsics_inv"]
    Pn = np.stack([oracle.project(K.numpy(), oracle.pose_vec2mat(d[k].numpy())) for k in ("T_2to1", "T_R2L")], 1)
    r = oracle.photo_loss_P(d["img_R2"].numpy(), [d["img_R1"].numpy(), d["img_L2"].numpy()], d["depth"].numpy(), Pn,
                            Kinv.numpy())
    tgt, s0, s1, depth = d["img_R2"].cuda(), d["img_R1"].cuda(), d["img_L2"].cuda(), d["depth"].cuda()
    P, Ki = cu(Pn), Kinv.cuda()
    gdepth = torch.empty_like(depth)
    gP = torch.empty(B, V, 3, 4, device="cuda")
    terms = torch.empty(V, device="cuda")
    lv = (_lib.dvf_level * 1)()
    lv[0].H, lv[0].W = H, W
    lv[0].depth, lv[0].tgt, lv[0].P, lv[0].Kinv = depth.data_ptr(), tgt.data_ptr(), P.data_ptr(), Ki.data_ptr()
    lv[0].src[0], lv[0].src[1] = s0.data_ptr(), s1.data_ptr()
    lv[0].gdepth, lv[0].gP = gdepth.data_ptr(), gP.data_ptr()
    desc = _lib.dvf_loss_desc(B, Cc, V, 1, 0, 0, 0, 0)
    n = lib.dvf_photo_loss_workspace_bytes(C.byref(desc), lv)
    assert n > 0
    ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for _ in range(3):   # reuse of the workspace + run-to-run determinism
        _lib.check(lib.dvf_photo_loss_fused(C.byref(desc), lv, terms.data_ptr(), ws.data_ptr(), n, st), "loss")
        outs.append((npy(terms).copy(), npy(gdepth).copy(), npy(gP).copy()))
    for o in outs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(o, outs[0])), "results must be bit-reproducible"
    assert_close(outs[0][0], r["terms"], what="terms")
    assert_close(outs[0][1], r["gdepth"], what="gdepth")
    assert_close(outs[0][2], r["gP"], what="gP")
    assert int((outs[0][1] != r["gdepth"]).sum()) == 0, "depth gradient follows the reference's fp32 sequence exactly"


def test_fused_loss_multilevel_matches_per_level(ops, oracle, syn):
    B, H, W, V, L = 2, 64, 208, 2, 4
    d = syn.stereo_temporal_batch(B, H, W, seed=41)
    K, Kinv = d["intrinsics"], d["intrinsics_inv"]
    depths = [syn.depth(B, H >> s, W >> s, 50 + s) for s in range(L)]
    expl = [syn.explainability(B, V, H >> s, W >> s, 60 + s) for s in range(L)]
    pose = torch.stack([d["T_2to1"], d["T_R2L"]], 1)
    sizes = [(H >> s, W >> s) for s in range(L)]
    tg = ops.area_pyramid(d["img_R2"].cuda(), sizes)
    r1 = ops.area_pyramid(d["img_R1"].cuda(), sizes)
    l2 = ops.area_pyramid(d["img_L2"].cuda(), sizes)
    for s in range(L):  # pyramid == torch-CPU 'area' interpolation, bit for bit
        assert np.array_equal(npy(tg[s]), oracle.area_downsample(d["img_R2"].numpy(), *sizes[s]))
    t_depths = [x.cuda().requires_grad_(True) for x in depths]
    t_expl = [x.cuda().requires_grad_(True) for x in expl]
    t_pose = pose.cuda().requires_grad_(True)
    ds = [H / s[0] for s in sizes]
    loss, terms = ops.fused_photo_loss(tg, [[r1[s], l2[s]] for s in range(L)], t_depths, t_pose, K.cuda(), Kinv.cuda(),
                                       expl_levels=t_expl, downscales=ds)
    loss.backward()
    pm, P, Ks = ops.pose_proj_fwd(t_pose.detach().reshape(B * V, 6), K.cuda(), Kinv.cuda(), V, "euler", ds, want_posemat=True)
    gpose = np.zeros((B, V, 6))
    for s in range(L):
        Pn = npy(P[s]).reshape(B, V, 3, 4)   # GPU P: per-pixel path compared exactly
        oK, oKi = oracle.scale_intrinsics(K.numpy(), Kinv.numpy(), ds[s])
        r = oracle.photo_loss_P(npy(tg[s]), [npy(r1[s]), npy(l2[s])], depths[s].numpy(), Pn, oKi, expl=expl[s].numpy())
        assert_close(npy(terms[s * V:(s + 1) * V]), r["terms"], what=f"terms level {s}")
        assert_close(npy(t_depths[s].grad), r["gdepth"], what=f"gdepth level {s}")
        assert_close(npy(t_expl[s].grad), r["gexpl"], what=f"gexpl level {s}")
        for v in range(V):
            gpose[:, v] += oracle.pose_bwd(r["gP"][:, v], oK, pose[:, v].numpy())
    assert_close(npy(t_pose.grad), gpose, what="gpose")


# ------------------------------------------------------------------------------------------------
# golden vectors of the real reference, through the drop-in modules
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names("iw_"))
def test_dropin_inverse_warp_golden(ops, name):
    with align_corners(ops, name):   # *_align fixtures: the torch <= 1.2 sampling convention (DVF_FLAG_ALIGN_CORNERS)
        _dropin_inverse_warp_golden(ops, name)


def _dropin_inverse_warp_golden(ops, name):
    import inverse_warp as iw
    g = golden(name)
    rot, pad = g["rotation_mode"], g["padding_mode"]
    img, depth, pose = cu(g["img"]).requires_grad_(True), cu(g["depth"]).requires_grad_(True), cu(g["pose"]).requires_grad_(True)
    K, Kinv = cu(g["K"]), cu(g["Kinv"])
    fn = iw.inverse_warp if g["img"].shape[1] == 3 else __import__("loss_functions").inverse_warp
    warped = fn(img, depth, pose, K, Kinv, rot, pad)
    pm = iw.pose_vec2mat(pose.detach(), rot)
    assert np.array_equal(npy(pm), g["posemat"]), "pose_vec2mat: bit-identical to the reference on torch-CPU"
    _, P, _ = ops.pose_proj_fwd(pose.detach(), K, None, 1, rot, [1.0])
    assert np.array_equal(npy(P[0]), g["P"]), "K @ pose_mat: bit-identical to the reference"
    warped.backward(cu(g["gout"]))
    assert np.array_equal(npy(warped), g["warped"]), "warped image: bit-identical through the public signature"
    valid = (warped != 0).any(1).to(torch.uint8)
    assert np.array_equal(npy(valid), g["valid"]), "validity mask: bit-exact"
    assert_close(npy(depth.grad), g["gdepth"], what="gdepth")
    assert_close(npy(img.grad), g["gimg"], what="gimg")
    assert_close(npy(pose.grad), g["gpose"], what="gpose")
    # exact-P path: feed the reference's own P through the C ABI -> bit-identical forward, 1e-5 backward
    w2, v2 = ops.inverse_warp_fwd_P(img.detach(), depth.detach(), cu(g["P"]), Kinv, pad, want_valid=True)
    assert np.array_equal(npy(w2), g["warped"]) and np.array_equal(npy(v2), g["valid"])
    gi, gd, gP = ops.inverse_warp_bwd_P(cu(g["gout"]), img.detach(), depth.detach(), cu(g["P"]), Kinv, pad)
    assert_close(npy(gd), g["gdepth"], what="gdepth (reference P)")
    assert_close(npy(gi), g["gimg"], what="gimg (reference P)")
    gv = ops.pose_proj_bwd(gP.unsqueeze(0), None, pose.detach(), K, 1, rot, [1.0])
    assert_close(npy(gv), g["gpose"], what="gpose (reference P)")


@pytest.mark.parametrize("name", golden_names("lf_"))
def test_dropin_loss_functions_golden(ops, name):
    with align_corners(ops, name):
        _dropin_loss_functions_golden(ops, name)


def _dropin_loss_functions_golden(ops, name):
    import loss_functions as lf
    g = golden(name)
    feat = "g_img_R1" in g
    t = {k: cu(g[k]) for k in ("img_R2", "img_R1", "img_L2", "depth", "T_2to1", "T_R2L", "intrinsics", "intrinsics_inv")}
    req = ["depth", "T_2to1", "T_R2L"] + (["img_R2", "img_R1", "img_L2"] if feat else [])
    for k in req:
        t[k].requires_grad_(True)
    loss = lf.photometric_reconstruction_loss(t["img_R2"], t["img_R1"], t["img_L2"], t["depth"], t["T_2to1"], t["T_R2L"],
                                              t["intrinsics"], t["intrinsics_inv"])
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= RTOL_F32 * abs(float(g["loss"]))
    pose = torch.stack([t["T_2to1"], t["T_R2L"]], 1).detach()
    B = pose.shape[0]
    _, P, _ = ops.pose_proj_fwd(pose.reshape(B * 2, 6), t["intrinsics"], None, 2, "euler", [1.0])
    assert np.array_equal(npy(P[0]).reshape(B, 2, 3, 4), g["P"]), "pose -> P: bit-identical to the reference"
    for k in req:
        assert_close(npy(t[k].grad), g["g_" + k], what="g_" + k)


@pytest.mark.parametrize("name", golden_names("sfm_"))
def test_dropin_sfm_golden(ops, name):
    g = golden(name)
    old = g["kind"] == "sfm_old"
    rot, pad, n = g["rotation_mode"], g["padding_mode"], int(g["n_scales"])
    with_mask = bool(g["with_mask"])
    depths = [cu(g[f"depth{s}"]).requires_grad_(True) for s in range(n)]
    masks = [cu(g[f"mask{s}"]).requires_grad_(True) if with_mask else None for s in range(n)]
    pose = cu(g["pose"]).requires_grad_(True)
    K, Kinv = cu(g["intrinsics"]), cu(g["intrinsics_inv"])
    R2, R1, L2 = cu(g["img_R2"]), cu(g["img_R1"]), cu(g["img_L2"])
    if old:
        import loss_function_sfm as m
        loss = m.photometric_reconstruction_loss(R2, R1, L2, depths, pose[:, 0], pose[:, 1], masks, K, Kinv, rot, pad)
    else:
        import loss_functions_sfm as m
        loss = m.photometric_reconstruction_loss(R2, [R1, L2], K, Kinv, depths, masks, pose, rot, pad)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= RTOL_F32 * abs(float(g["loss"]))
    B = pose.shape[0]
    H = R2.shape[2]
    ds = [H / depths[s].shape[2] for s in range(n)]
    _, P, _ = ops.pose_proj_fwd(pose.detach().reshape(B * 2, 6), K, None, 2, rot, ds)
    for s in range(n):
        assert np.array_equal(npy(P[s]).reshape(B, 2, 3, 4), g[f"P{s}"]), f"level {s}: pose -> P bit-identical to the reference"
    for s in range(n):
        assert_close(npy(depths[s].grad), g[f"g_depth{s}"], what=f"g_depth{s}")
        if with_mask:
            assert_close(npy(masks[s].grad), g[f"g_mask{s}"], what=f"g_mask{s}")
    gp = npy(pose.grad) if pose.grad is not None else np.zeros_like(g["g_pose"])
    assert_close(gp, g["g_pose"], what="g_pose")


def test_full_size_properties(ops, syn):
    """BASELINE config sizes (B=64, 128x416, 4 scales): size-independent properties instead of the oracle --
    linearity of the loss in the mask weights, batch-sharding invariance, and gradient/loss consistency."""
    B, H, W, L, V = 64, 128, 416, 4, 1
    d = syn.stereo_temporal_batch(B, H, W, seed=71)
    K, Kinv = d["intrinsics"].cuda(), d["intrinsics_inv"].cuda()
    sizes = [(H >> s, W >> s) for s in range(L)]
    ds = [float(1 << s) for s in range(L)]
    tg = ops.area_pyramid(d["img_R2"].cuda(), sizes)
    sr = ops.area_pyramid(d["img_L2"].cuda(), sizes)
    depths = [syn.depth(B, h, w, 80 + i).cuda() for i, (h, w) in enumerate(sizes)]
    pose = d["T_R2L"].cuda().unsqueeze(1)

    def run(sl, expl=None):
        dl = [x[sl].clone().requires_grad_(True) for x in depths]
        p = pose[sl].clone().requires_grad_(True)
        loss, terms = ops.fused_photo_loss([x[sl] for x in tg], [[x[sl]] for x in sr], dl, p, K[sl], Kinv[sl],
                                           expl_levels=expl, downscales=ds)
        loss.backward()
        return loss.item(), terms.detach().cpu().numpy().astype(np.float64), [x.grad for x in dl], p.grad

    full = slice(0, B)
    loss, terms, gd, gp = run(full)
    assert np.isfinite(loss) and loss > 0
    # (1) mask linearity: a constant explainability weight c scales every term and gradient by c
    c = 0.25
    ex = [torch.full((B, 1, h, w), c, device="cuda") for (h, w) in sizes]
    loss_c, terms_c, gd_c, gp_c = run(full, ex)
    np.testing.assert_allclose(terms_c, c * terms, rtol=1e-6)
    assert_close(gp_c.cpu().numpy(), c * gp.cpu().numpy(), what="pose grad linearity")
    assert_close(gd_c[0].cpu().numpy(), c * gd[0].cpu().numpy(), what="depth grad linearity")
    # (2) sharding the batch in two halves: terms average, gradients concatenate (scaled by the batch ratio)
    la, ta, gda, gpa = run(slice(0, B // 2))
    lb, tb, gdb, gpb = run(slice(B // 2, B))
    np.testing.assert_allclose(0.5 * (ta + tb), terms, rtol=1e-6)
    both = torch.cat([gda[0], gdb[0]]) * 0.5
    assert torch.equal(both, gd[0]), "per-pixel depth gradients are independent of the batch split (up to the exact 1/2)"
    assert_close(torch.cat([gpa, gpb]).cpu().numpy() * 0.5, gp.cpu().numpy(), what="pose grad sharding")
    # (3) determinism
    loss2, terms2, gd2, gp2 = run(full)
    assert loss2 == loss and all(torch.equal(a, b) for a, b in zip(gd, gd2)) and torch.equal(gp, gp2)


def test_errors_are_loud(ops, syn):
    from dvf_b200 import DvfError
    import inverse_warp as iw
    d = syn.stereo_temporal_batch(1, 8, 12, seed=1)
    with pytest.raises(DvfError):   # CPU tensors: no fallback
        iw.inverse_warp(d["img_R1"], d["depth"], d["T_2to1"], d["intrinsics"], d["intrinsics_inv"])
    with pytest.raises(AssertionError):   # reference's size check (inverse_warp.py:173)
        iw.inverse_warp(d["img_R1"].cuda()[:, :2], d["depth"].cuda(), d["T_2to1"].cuda(), d["intrinsics"].cuda(),
                        d["intrinsics_inv"].cuda())


def test_fused_pose_entry_matches_three_launch_form(ops, syn):
    """dvf_photo_loss_fused_pose (pose chain inside the kernel) == pose_proj_fwd + photo_loss_fused + pose_proj_bwd,
    bit for bit, multi-level, V=2, with masks."""
    from dvf_b200.plan import FusedLossPlan
    B, H, W, V, L = 3, 64, 208, 2, 3
    d = syn.stereo_temporal_batch(B, H, W, seed=91)
    sizes = [(H >> s, W >> s) for s in range(L)]
    tg = ops.area_pyramid(d["img_R2"].cuda(), sizes)
    r1 = ops.area_pyramid(d["img_R1"].cuda(), sizes)
    l2 = ops.area_pyramid(d["img_L2"].cuda(), sizes)
    depths = [syn.depth(B, h, w, 100 + i).cuda() for i, (h, w) in enumerate(sizes)]
    expl = [syn.explainability(B, V, h, w, 110 + i).cuda() for i, (h, w) in enumerate(sizes)]
    pose = torch.stack([d["T_2to1"], d["T_R2L"]], 1).cuda().contiguous()
    for rot in ("euler", "quat"):
        plan = FusedLossPlan(tg, [[r1[s], l2[s]] for s in range(L)], depths, pose, d["intrinsics"].cuda(),
                             d["intrinsics_inv"].cuda(), expl_levels=expl, downscales=[float(1 << s) for s in range(L)],
                             rotation_mode=rot)
        outs = []
        for fused in (False, True, True):
            plan.terms.fill_(-1); plan.gpose.fill_(-1)
            for g in plan.gdepth + plan.gexpl:
                g.fill_(-1)
            plan.launch(fused_pose=fused)
            torch.cuda.synchronize()
            outs.append([plan.terms.clone(), plan.gpose.clone()] + [g.clone() for g in plan.gdepth + plan.gexpl])
        for other in outs[1:]:
            for a, b in zip(outs[0], other):
                assert torch.equal(a, b)
        assert float(outs[0][0].min()) > 0


@pytest.mark.parametrize("dtype,C,V,with_expl", [("f32", 64, 2, False), ("f32", 32, 1, True), ("f32", 8, 2, False),
                                                 ("bf16", 64, 2, False), ("bf16", 32, 2, True), ("bf16", 16, 1, False)])
def test_feature_loss_channels_last_vs_oracle(ops, oracle, syn, dtype, C, V, with_expl):
    """channels-last (NHWC) feature maps, fp32 and bf16: the NHWC kernel against the oracle.  bf16: maps are stored
    in bf16, arithmetic and geometry in fp32 -> compared with the oracle on the bf16-rounded maps at 1e-2
    (BASELINE north_star tolerance for bf16); fp32: 1e-5."""
    B, H, W = 2, 32, 104
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    maps = [m.to(tdt) for m in syn.features(B, C, H, W, 17, n=V + 1)]
    ref_maps = [m.float().numpy() for m in maps]                 # what the kernel actually reads
    depth = syn.depth(B, H, W, 18)
    kinds = ["kitti", "stereo"]
    pose = torch.stack([syn.pose(B, kinds[v], 19 + v) for v in range(V)], 1)
    K, Kinv = syn.intrinsics(B, H, W)
    expl = syn.explainability(B, V, H, W, 21) if with_expl else None

    cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)   # noqa: E731
    t_maps = [cl(m).requires_grad_(True) for m in maps]
    assert ops._nhwc_ok(t_maps[0])
    t_depth = depth.cuda().requires_grad_(True)
    t_pose = pose.cuda().requires_grad_(True)
    t_expl = None if expl is None else expl.cuda().requires_grad_(True)
    loss, terms = ops.fused_photo_loss([t_maps[0]], [t_maps[1:]], [t_depth], t_pose, K.cuda(), Kinv.cuda(),
                                       expl_levels=None if expl is None else [t_expl])
    loss.backward()
    _, P_gpu, _ = ops.pose_proj_fwd(t_pose.detach().reshape(B * V, 6), K.cuda(), None, V, "euler", [1.0])
    Pn = npy(P_gpu[0]).reshape(B, V, 3, 4)
    r = oracle.photo_loss_P(ref_maps[0], ref_maps[1:], depth.numpy(), Pn, Kinv.numpy(),
                            expl=None if expl is None else expl.numpy(), need_gsrc=True, need_gtgt=True)
    tol_geo = RTOL_F32 if dtype == "f32" else 1e-4    # fp32 outputs; bf16 only changes what is read
    tol_map = RTOL_F32 if dtype == "f32" else 1e-2    # map gradients come back in the map dtype
    assert_close(npy(terms), r["terms"], tol=tol_geo, what="terms")
    assert_close(npy(t_depth.grad), r["gdepth"], tol=tol_geo, what="gdepth")
    assert t_maps[0].grad.dtype == tdt and t_maps[0].grad.is_contiguous(memory_format=torch.channels_last)
    assert_close(npy(t_maps[0].grad.float()), r["gtgt"], tol=tol_map, what="gtgt")
    for v in range(V):
        assert_close(npy(t_maps[1 + v].grad.float()), r["gsrc"][v], tol=tol_map, what=f"gsrc{v}")
        assert_close(npy(t_pose.grad[:, v]), oracle.pose_bwd(r["gP"][:, v], K.numpy(), pose[:, v].numpy()), tol=tol_geo,
                     what=f"gpose{v}")
    if with_expl:
        assert_close(npy(t_expl.grad), r["gexpl"], tol=tol_geo, what="gexpl")


@pytest.mark.parametrize("dtype,C,V,B,sizes,with_expl", [
    ("f32", 128, 3, 2, [(17, 23)], True),            # 32 lanes per pixel, ragged map, three views
    ("bf16", 8, 4, 3, [(9, 11)], False),             # one lane per pixel, four views
    ("f32", 16, 2, 2, [(2, 5)], False),              # two short rows
    ("f32", 32, 2, 36, [(32, 104), (16, 52)], True),  # two levels, more units than CTAs: multi-piece CTAs
    ("bf16", 64, 1, 20, [(32, 104), (16, 52), (8, 26)], False),
])
def test_feature_loss_channels_last_shapes_vs_oracle(ops, oracle, syn, dtype, C, V, B, sizes, with_expl):
    """channels-last kernel on awkward shapes (lanes per pixel 1..32, ragged maps, several levels, CTAs that own
    several pieces) against the oracle, every gradient."""
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    L = len(sizes)
    H0, W0 = sizes[0]
    K, Kinv = syn.intrinsics(B, H0, W0)
    ds = [H0 / h for (h, w) in sizes]
    kinds = ["kitti", "stereo", "tiny", "large"]
    pose = torch.stack([syn.pose(B, kinds[v], 31 + v) for v in range(V)], 1)
    maps = [[m.to(tdt) for m in syn.features(B, C, h, w, 40 + 7 * i, n=V + 1)] for i, (h, w) in enumerate(sizes)]
    depths = [syn.depth(B, h, w, 50 + i) for i, (h, w) in enumerate(sizes)]
    expl = [syn.explainability(B, V, h, w, 60 + i) for i, (h, w) in enumerate(sizes)] if with_expl else None
    cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)   # noqa: E731
    t_maps = [[cl(m).requires_grad_(True) for m in lv] for lv in maps]
    t_depths = [d.cuda().requires_grad_(True) for d in depths]
    t_expl = [e.cuda().requires_grad_(True) for e in expl] if with_expl else None
    t_pose = pose.cuda().requires_grad_(True)
    loss, terms = ops.fused_photo_loss([lv[0] for lv in t_maps], [lv[1:] for lv in t_maps], t_depths, t_pose, K.cuda(),
                                       Kinv.cuda(), expl_levels=t_expl, downscales=ds)
    loss.backward()
    _, P, _ = ops.pose_proj_fwd(t_pose.detach().reshape(B * V, 6), K.cuda(), Kinv.cuda(), V, "euler", ds)
    tol_geo = RTOL_F32 if dtype == "f32" else 1e-4
    tol_map = RTOL_F32 if dtype == "f32" else 1e-2
    gpose = np.zeros((B, V, 6))
    for s in range(L):
        Pn = npy(P[s]).reshape(B, V, 3, 4)
        oK, oKi = oracle.scale_intrinsics(K.numpy(), Kinv.numpy(), ds[s])
        ref = [m.float().numpy() for m in maps[s]]
        r = oracle.photo_loss_P(ref[0], ref[1:], depths[s].numpy(), Pn, oKi, expl=None if expl is None else expl[s].numpy(),
                                need_gsrc=True, need_gtgt=True)
        assert_close(npy(terms[s * V:(s + 1) * V]), r["terms"], tol=tol_geo, what=f"terms level {s}")
        assert_close(npy(t_depths[s].grad), r["gdepth"], tol=tol_geo, what=f"gdepth level {s}")
        assert_close(npy(t_maps[s][0].grad.float()), r["gtgt"], tol=tol_map, what=f"gtgt level {s}")
        for v in range(V):
            assert_close(npy(t_maps[s][1 + v].grad.float()), r["gsrc"][v], tol=tol_map, what=f"gsrc{v} level {s}")
            gpose[:, v] += oracle.pose_bwd(r["gP"][:, v], oK, pose[:, v].numpy())
        if with_expl:
            assert_close(npy(t_expl[s].grad), r["gexpl"], tol=tol_geo, what=f"gexpl level {s}")
    assert_close(npy(t_pose.grad), gpose, tol=tol_geo, what="gpose")


@pytest.mark.parametrize("shape,dtype", [((3, 64, 32, 104), torch.float32), ((2, 8, 5, 7), torch.float32),
                                         ((1, 33, 9, 31), torch.float32), ((2, 64, 32, 104), torch.bfloat16),
                                         ((1, 16, 3, 5), torch.bfloat16)])
def test_to_channels_last_matches_torch(ops, shape, dtype):
    """dvf_transpose_planes behind ops.to_channels_last: a pure re-layout, so the bits must equal torch's."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(*shape, generator=g).to(dtype).cuda()
    y = ops.to_channels_last(x)
    assert y.shape == x.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(y, x)
    assert torch.equal(y.permute(0, 2, 3, 1).contiguous(), x.permute(0, 2, 3, 1).contiguous())


def test_regularisers_vs_oracle(ops, oracle, syn):
    """smooth_loss / explainability_loss, all scales in one launch, against the oracle (value and gradient)."""
    import loss_functions as lf
    import loss_functions_sfm as sfm
    B, H, W = 3, 64, 208
    maps = [syn.depth(B, H >> s, W >> s, 200 + s).unsqueeze(1) for s in range(4)]
    t = [m.cuda().requires_grad_(True) for m in maps]
    val = lf.smooth_loss(t, 2.0)
    (3.0 * val).backward()
    ref, w = 0.0, 1.0
    for m, tm in zip(maps, t):
        v, g = oracle.smooth_loss_one(m[:, 0].numpy(), need_grad=True)
        ref += v * w
        assert_close(npy(tm.grad[:, 0]), 3.0 * w * g, what="smooth grad")
        w /= 2.0
    assert abs(val.item() - ref) <= RTOL_F32 * ref
    masks = [syn.explainability(B, 2, H >> s, W >> s, 210 + s) for s in range(3)]
    tm = [m.cuda().requires_grad_(True) for m in masks]
    val = sfm.explainability_loss(tm)
    val.backward()
    ref = 0.0
    for m, x in zip(masks, tm):
        v, g = oracle.explainability_loss_one(m.numpy(), need_grad=True)
        ref += v
        assert_close(npy(x.grad), g, what="explainability grad")
    assert abs(val.item() - ref) <= RTOL_F32 * ref
    # ragged / degenerate sizes
    small = torch.rand(2, 1, 2, 5).cuda().requires_grad_(True)
    v = lf.smooth_loss(small)
    ov, og = oracle.smooth_loss_one(small.detach().cpu().numpy()[:, 0], need_grad=True)
    v.backward()
    assert abs(v.item() - ov) <= 1e-5 * max(ov, 1e-6)
    assert_close(npy(small.grad[:, 0]), og, what="small smooth grad")


@pytest.mark.parametrize("B,H,W,V,with_expl", [(2, 17, 23, 2, True), (1, 2, 3, 1, False), (3, 40, 52, 1, False), (1, 128, 416, 4, True)])
def test_fused_loss_c3_ragged_shapes_vs_oracle(ops, oracle, syn, B, H, W, V, with_expl):
    """image kernel (C=3) on sizes that exercise the non-bulk-copy path (H*W % 4 != 0), partial last chunks,
    several views and masks; the oracle is fed the GPU's own P so the per-pixel path is held to 1e-5."""
    imgs = syn.images(B, 3, H, W, 31, n=V + 1, smooth=min(H, W) > 9)
    depth = syn.depth(B, H, W, 32, smooth=min(H, W) > 17)
    kinds = ["kitti", "stereo", "tiny", "large"]
    pose = torch.stack([syn.pose(B, kinds[v], 33 + v) for v in range(V)], 1)
    K, Kinv = syn.intrinsics(B, H, W)
    expl = syn.explainability(B, V, H, W, 39) if with_expl else None
    t_depth = depth.cuda().requires_grad_(True)
    t_pose = pose.cuda().requires_grad_(True)
    t_expl = None if expl is None else expl.cuda().requires_grad_(True)
    loss, terms = ops.fused_photo_loss([imgs[0].cuda()], [[m.cuda() for m in imgs[1:]]], [t_depth], t_pose, K.cuda(), Kinv.cuda(),
                                       expl_levels=None if expl is None else [t_expl])
    loss.backward()
    _, P_gpu, _ = ops.pose_proj_fwd(t_pose.detach().reshape(B * V, 6), K.cuda(), None, V, "euler", [1.0])
    Pn = npy(P_gpu[0]).reshape(B, V, 3, 4)
    r = oracle.photo_loss_P(imgs[0].numpy(), [m.numpy() for m in imgs[1:]], depth.numpy(), Pn, Kinv.numpy(),
                            expl=None if expl is None else expl.numpy())
    assert_close(npy(terms), r["terms"], what="terms")
    assert np.array_equal(npy(t_depth.grad), r["gdepth"]), "depth gradient is bit-identical to the reference sequence"
    for v in range(V):
        assert_close(npy(t_pose.grad[:, v]), oracle.pose_bwd(r["gP"][:, v], K.numpy(), pose[:, v].numpy()), what=f"gpose{v}")
    if with_expl:
        assert np.array_equal(npy(t_expl.grad), r["gexpl"])


@pytest.mark.parametrize("V,with_expl", [(1, False), (2, True)])
def test_fused_loss_multi_piece_ctas_vs_oracle(ops, oracle, syn, V, with_expl):
    """Enough work that every CTA of the balanced split owns several units and crosses image / level boundaries
    (more units than resident CTAs: pieces, partial-sum slots, ring parity across pieces), against the oracle."""
    B, H, W, L = 40, 64, 208, 3
    d = syn.stereo_temporal_batch(B, H, W, seed=91)
    K, Kinv = d["intrinsics"], d["intrinsics_inv"]
    sizes = [(H >> s, W >> s) for s in range(L)]
    ds = [float(1 << s) for s in range(L)]
    depths = [syn.depth(B, h, w, 92 + i) for i, (h, w) in enumerate(sizes)]
    expl = [syn.explainability(B, V, h, w, 95 + i) for i, (h, w) in enumerate(sizes)] if with_expl else None
    pose = torch.stack([d["T_R2L"], d["T_2to1"]][:V], 1)
    tg = ops.area_pyramid(d["img_R2"].cuda(), sizes)
    srcs = [ops.area_pyramid(d[n].cuda(), sizes) for n in ["img_L2", "img_R1"][:V]]
    t_depths = [x.cuda().requires_grad_(True) for x in depths]
    t_expl = [x.cuda().requires_grad_(True) for x in expl] if with_expl else None
    t_pose = pose.cuda().requires_grad_(True)
    loss, terms = ops.fused_photo_loss(tg, [[srcs[v][s] for v in range(V)] for s in range(L)], t_depths, t_pose,
                                       K.cuda(), Kinv.cuda(), expl_levels=t_expl, downscales=ds)
    loss.backward()
    _, P, _ = ops.pose_proj_fwd(t_pose.detach().reshape(B * V, 6), K.cuda(), Kinv.cuda(), V, "euler", ds)
    gpose = np.zeros((B, V, 6))
    for s in range(L):
        Pn = npy(P[s]).reshape(B, V, 3, 4)
        oK, oKi = oracle.scale_intrinsics(K.numpy(), Kinv.numpy(), ds[s])
        r = oracle.photo_loss_P(npy(tg[s]), [npy(srcs[v][s]) for v in range(V)], depths[s].numpy(), Pn, oKi,
                                expl=None if expl is None else expl[s].numpy())
        assert_close(npy(terms[s * V:(s + 1) * V]), r["terms"], what=f"terms level {s}")
        assert int((npy(t_depths[s].grad) != r["gdepth"]).sum()) == 0, f"depth gradient level {s} bit-exact"
        if with_expl:
            assert_close(npy(t_expl[s].grad), r["gexpl"], what=f"gexpl level {s}")
        for v in range(V):
            gpose[:, v] += oracle.pose_bwd(r["gP"][:, v], oK, pose[:, v].numpy())
    assert_close(npy(t_pose.grad), gpose, what="gpose")


def test_fused_loss_forward_only_and_nan_inputs(ops, oracle, syn):
    """no_grad mode runs the loss-only kernel; NaN / huge depths take the exact cold path and propagate like the reference."""
    B, H, W = 2, 24, 80
    d = syn.stereo_temporal_batch(B, H, W, seed=55)
    pose = torch.stack([d["T_2to1"], d["T_R2L"]], 1).cuda()
    depth = d["depth"].clone()
    with torch.no_grad():
        loss, terms = ops.fused_photo_loss([d["img_R2"].cuda()], [[d["img_R1"].cuda(), d["img_L2"].cuda()]], [depth.cuda()], pose,
                                           d["intrinsics"].cuda(), d["intrinsics_inv"].cuda())
    _, P_gpu, _ = ops.pose_proj_fwd(pose.reshape(B * 2, 6), d["intrinsics"].cuda(), None, 2, "euler", [1.0])
    Pn = npy(P_gpu[0]).reshape(B, 2, 3, 4)
    r = oracle.photo_loss_P(d["img_R2"].numpy(), [d["img_R1"].numpy(), d["img_L2"].numpy()], depth.numpy(), Pn,
                            d["intrinsics_inv"].numpy())
    assert_close(npy(terms), r["terms"], what="terms (forward only)")
    # huge (but finite) depths force the __fdiv_rn path on those pixels: results must still match the oracle exactly
    depth2 = depth.clone()
    depth2[0, 3, 5] = 3.0e12
    depth2[1, 7, 9] = 1.0e20
    t_depth = depth2.cuda().requires_grad_(True)
    loss, terms = ops.fused_photo_loss([d["img_R2"].cuda()], [[d["img_R1"].cuda(), d["img_L2"].cuda()]], [t_depth], pose,
                                       d["intrinsics"].cuda(), d["intrinsics_inv"].cuda())
    loss.backward()
    r2 = oracle.photo_loss_P(d["img_R2"].numpy(), [d["img_R1"].numpy(), d["img_L2"].numpy()], depth2.numpy(), Pn,
                             d["intrinsics_inv"].numpy())
    assert_close(npy(terms), r2["terms"], what="terms (huge depths)")
    assert np.array_equal(npy(t_depth.grad), r2["gdepth"])
    # NaN depth: the loss becomes NaN, as in the reference
    depth3 = depth.clone()
    depth3[0, 0, 0] = float("nan")
    with torch.no_grad():
        loss3, _ = ops.fused_photo_loss([d["img_R2"].cuda()], [[d["img_R1"].cuda(), d["img_L2"].cuda()]], [depth3.cuda()], pose,
                                        d["intrinsics"].cuda(), d["intrinsics_inv"].cuda())
    assert torch.isnan(loss3)


def test_se3_exp_dropin_golden(ops, oracle):
    import se3_generate
    g = golden("se3_exp")
    x = cu(g["vec"]).requires_grad_(True)
    out = se3_generate.generate_se3(x)
    assert out.dtype == torch.float64 and tuple(out.shape) == tuple(g["out"].shape)
    # theta, c1, c2 are fp32 in the reference; CUDA sinf vs numpy's sin may differ in the last place
    assert np.abs(npy(out) - g["out"]).max() < 5e-7
    out.backward(cu(g["gout"]))
    assert_close(npy(x.grad), g["gvec"], tol=1e-5, what="se3 grad")
    big = np.random.default_rng(3).standard_normal((257, 6, 1, 1)).astype(np.float32)
    o2 = se3_generate.generate_se3(cu(big))
    assert np.abs(npy(o2)[:, 0] - oracle.se3_exp(big)).max() < 5e-6


# ---- Caffe-convention layers -- oracle is "parity unpinned" for this family --------------------------
def _caffe_case(N, H, W, C, seed):
    rng = np.random.default_rng(seed)
    depth = rng.uniform(2.0, 40.0, (N, H, W)).astype(np.float32)
    T = np.tile(np.eye(4, dtype=np.float32), (N, 1, 1))
    w = rng.standard_normal((N, 3)) * 0.02
    for n in range(N):
        wx, wy, wz = w[n]
        T[n, :3, :3] += np.array([[0, -wz, wy], [wz, 0, -wx], [-wy, wx, 0]], np.float32)
    T[:, :3, 3] = rng.standard_normal((N, 3)).astype(np.float32) * np.array([0.54, 0.05, 0.3], np.float32)
    K = np.stack([np.full(N, 0.58 * W), np.full(N, 1.92 * H), np.full(N, 0.49 * W), np.full(N, 0.5 * H)], 1).astype(np.float32)
    K += rng.standard_normal(K.shape).astype(np.float32)
    img = rng.uniform(0, 1, (N, C, H, W)).astype(np.float32)
    tgt = rng.uniform(0, 1, (N, C, H, W)).astype(np.float32)
    return depth, T, K, img, tgt


@pytest.mark.parametrize("shape", [(2, 16, 52, 3), (3, 37, 61, 1), (4, 128, 416, 3), (1, 20, 33, 8)])
def test_caffe_layers_vs_oracle(ops, oracle, shape):
    import geo_transform as gt
    N, H, W, C = shape
    depth, T, K, img, tgt = _caffe_case(N, H, W, C, seed=sum(shape))
    d = cu(depth[:, None]).requires_grad_(True)
    Tt = cu(T[:, None]).requires_grad_(True)
    Kt = cu(K[:, :, None, None]).requires_grad_(True)
    im = cu(img).requires_grad_(True)
    pts = gt.geo_transform(d, Tt, Kt)
    xy = gt.pin_hole_project(pts, Kt)
    wr = gt.inverse_warp(im, xy)
    loss = gt.abs_loss(wr, cu(tgt))
    # forward: same expressions, same roundings -> bit-exact
    o_pts = oracle.caffe_geo_fwd(depth, T, K)
    o_xy = oracle.caffe_pinhole_fwd(o_pts, K)
    o_wr = oracle.caffe_warp_fwd(img, o_xy)
    o_loss, o_ga, _ = oracle.caffe_abs_loss(o_wr, tgt)
    assert np.array_equal(npy(pts), o_pts)
    assert np.array_equal(npy(xy), o_xy)
    assert np.array_equal(npy(wr), o_wr)
    assert abs(float(loss.detach()) - o_loss) <= 1e-6 * abs(o_loss)
    loss.backward()
    o_gi, o_gxy = oracle.caffe_warp_bwd(o_ga, img, o_xy)
    o_gp, o_gK2 = oracle.caffe_pinhole_bwd(o_gxy, o_pts, K)
    o_gd, o_gT, o_gK1 = oracle.caffe_geo_bwd(o_gp, depth, T, K)
    assert_close(npy(im.grad), o_gi, what="img_diff")                 # atomics: order differs
    assert np.array_equal(npy(d.grad)[:, 0], o_gd)                    # per-pixel chain: bit-exact
    assert_close(npy(Tt.grad).reshape(N, 16), o_gT, tol=2e-5, what="T_diff")   # fp32 tree sum vs fp64 ordered sum
    assert_close(npy(Kt.grad).reshape(N, 4), o_gK1 + o_gK2, tol=2e-5, what="K_diff")
    assert np.all(npy(Tt.grad).reshape(N, 16)[:, 12:] == 0)


def test_caffe_layers_edge_cases(ops, oracle):
    import geo_transform as gt
    N, H, W, C = 2, 12, 20, 3
    depth, T, K, img, tgt = _caffe_case(N, H, W, C, seed=5)
    # coordinates far outside, exactly on the border, negative, and a zero-depth pixel (Z + 1e-12 path)
    xy = np.zeros((N, 2, H, W), np.float32)
    rng = np.random.default_rng(1)
    xy[:, 0] = rng.uniform(-3, W + 2, (N, H, W))
    xy[:, 1] = rng.uniform(-3, H + 2, (N, H, W))
    xy[0, :, 0, 0] = (W - 1, H - 1)
    xy[0, :, 0, 1] = (-1.0, -1.0)
    xy[0, :, 0, 2] = (1e9, -1e9)
    xy[0, :, 0, 3] = (0.0, 0.0)
    out = gt.inverse_warp(cu(img), cu(xy))
    assert np.array_equal(npy(out), oracle.caffe_warp_fwd(img, xy))
    g = rng.standard_normal(img.shape).astype(np.float32)
    x = cu(xy).requires_grad_(True)
    gt.inverse_warp(cu(img), x).backward(cu(g))
    assert np.array_equal(npy(x.grad), oracle.caffe_warp_bwd(g, img, xy, need_gimg=False)[1])
    depth[0, 0, 0] = 0.0
    pts = oracle.caffe_geo_fwd(depth, T, K)
    pts[1, 2, 3, 3] = 0.0
    got = gt.pin_hole_project(cu(pts), cu(K[:, :, None, None]))
    assert np.array_equal(npy(got), oracle.caffe_pinhole_fwd(pts, K), equal_nan=True)
    # sign(0) = -1 in AbsLoss
    a = cu(img).requires_grad_(True)
    gt.abs_loss(a, cu(img)).backward()
    assert np.all(npy(a.grad) == np.float32(-1.0 / N))
    with pytest.raises(Exception):
        gt.geo_transform(torch.zeros(1, 1, 4, 4), torch.eye(4).view(1, 1, 4, 4), torch.ones(1, 4, 1, 1))   # CPU tensors
    with pytest.raises(AssertionError):
        gt.pin_hole_project(cu(pts[:, :2]), cu(K[:, :, None, None]))


def test_caffe_chain_matches_pytorch_convention(ops, oracle):
    """The Caffe-era and PyTorch-era formulations describe the same warp: with K = (fx,fy,cx,cy) and the same
    rigid motion, pixel-space sampling (Caffe) equals grid_sample(align_corners=False) up to the reference's known
    (W-1) normalisation shift -- so compare against the closed form instead: identity motion returns the image."""
    import geo_transform as gt
    N, H, W, C = 2, 24, 40, 3
    depth, T, K, img, _ = _caffe_case(N, H, W, C, seed=9)
    T[:] = np.eye(4, dtype=np.float32)
    xy = gt.pin_hole_project(gt.geo_transform(cu(depth[:, None]), cu(T[:, None]), cu(K[:, :, None, None])), cu(K[:, :, None, None]))
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    assert np.abs(npy(xy)[:, 0] - xs).max() < 1e-3 and np.abs(npy(xy)[:, 1] - ys).max() < 1e-3
    out = gt.inverse_warp(cu(img), cu(np.broadcast_to(np.stack([xs, ys])[None], (N, 2, H, W)).copy()))
    assert np.array_equal(npy(out), img)


@pytest.mark.parametrize("shape", [(1, 1, 5), (2, 2, 2), (3, 7, 9), (2, 6, 3), (1, 5, 1), (2, 3, 300), (1, 130, 2), (2, 9, 33)])
def test_smooth_loss_small_and_ragged_maps(ops, oracle, shape):
    """strips of 4 rows per thread: heights that are not multiples of 4, maps narrower / shorter than a stencil
    (terms whose element count is 0 contribute NaN in the reference -- mean of an empty tensor -- so only shapes
    with H, W >= 3 are compared in value; smaller ones must simply not crash or touch memory out of bounds)"""
    B, H, W = shape
    rng = np.random.default_rng(H * 100 + W)
    m = rng.uniform(1.0, 30.0, (B, H, W)).astype(np.float32)
    m[:, :: 2] = np.round(m[:, :: 2])          # exact zeros among the second differences: sign(0) = 0
    t = cu(m[:, None]).requires_grad_(True)
    val = ops.smooth_loss([t])
    val.backward()
    torch.cuda.synchronize()
    if H >= 3 and W >= 3:
        v, g = oracle.smooth_loss_one(m, need_grad=True)
        assert abs(val.item() - v) <= RTOL_F32 * max(abs(v), 1e-30)
        assert_close(npy(t.grad[:, 0]), g, what="smooth grad")
    else:
        assert t.grad.shape == t.shape


def test_explainability_loss_edge_values(ops, oracle):
    """mask values at and beyond the ends of (0, 1): log clamp at -100, gradient denominator clamp at 1e-12, NaN"""
    # (the reference's binary_cross_entropy rejects values outside [0, 1]; NaN is let through)
    m = np.array([0.0, 1.0, 1e-30, 1e-12, 0.5, 1.0 - 2.0 ** -24, 1e-45, 0.999, np.nan, 2.0 ** -126, 0.25, 1e-7], np.float32)
    m = np.tile(m, 11)[: 125].reshape(1, 1, 5, 25).copy()
    t = cu(m).requires_grad_(True)
    val = ops.explainability_loss([t])
    val.backward()
    v, g = oracle.explainability_loss_one(m, need_grad=True)
    fin = np.isfinite(g)
    assert np.array_equal(np.isfinite(npy(t.grad)), fin)
    assert_close(npy(t.grad)[fin], g[fin], what="expl grad")
    assert np.array_equal(npy(t.grad)[~fin], g[~fin], equal_nan=True)
    assert (np.isnan(v) and np.isnan(val.item())) or abs(val.item() - v) <= RTOL_F32 * abs(v)


def test_regularisers_dropins_vs_reference_golden(ops):
    """the drop-in smooth_loss / explainability_loss against the reference's own values and gradients"""
    import loss_functions as lf
    import loss_functions_sfm as sfm
    g = golden("regularisers")
    t = [cu(g[f"map{i}"]).requires_grad_(True) for i in range(int(g["n_maps"]))]
    val = lf.smooth_loss(t, 2.0)
    val.backward()
    assert abs(val.item() - float(g["smooth"])) <= RTOL_F32 * float(g["smooth"])
    for i, x in enumerate(t):
        assert_close(npy(x.grad), g[f"g_map{i}"], what=f"smooth grad {i}")
    tm = [cu(g[f"mask{i}"]).requires_grad_(True) for i in range(int(g["n_masks"]))]
    ev = sfm.explainability_loss(tm)
    ev.backward()
    assert abs(ev.item() - float(g["expl"])) <= RTOL_F32 * float(g["expl"])
    for i, x in enumerate(tm):
        assert_close(npy(x.grad), g[f"g_mask{i}"], what=f"expl grad {i}")


def test_caffe_edge_aware_smoothness_vs_oracle(ops, oracle):
    """Caffe graphs' edge-aware smoothness (experiments/depth/train.prototxt:4022-4234; parity unpinned): the fused gather
    kernel against the oracle's layer-by-layer scatter form, value and gradient."""
    rng = np.random.default_rng(9)
    for (N, H, W) in [(2, 9, 11), (3, 40, 64), (1, 3, 3), (2, 160, 608)]:
        img = rng.random((N, 3, H, W), dtype=np.float32)
        inv = (rng.random((N, 1, H, W), dtype=np.float32) * 0.3 + 0.02).astype(np.float32)
        inv[0, 0, H // 2, :] = inv[0, 0, H // 2, 0]          # exact zeros in dy: AbsLoss's sign(0) = +1
        loss, g = oracle.caffe_edge_smooth(img, inv, weight=10.0)
        d = cu(inv).requires_grad_(True)
        out = ops.edge_aware_smoothness(cu(img), d, 10.0)
        out.backward()
        assert abs(out.item() - 10.0 * loss.sum()) <= 2e-6 * 10.0 * loss.sum()
        assert_close(npy(d.grad), g, tol=2e-6, what=f"d inv_depth {N}x{H}x{W}")


def test_caffe_two_view_batch_concatenated_form(ops, syn):
    """experiments/depth_odometry/train.prototxt:4309-4437: the stereo and the temporal view concatenated along the batch
    axis go through the geometry layers in one pass; the result equals two separate single-view passes (values and
    gradients; parity of the layer family itself is unpinned)."""
    import geo_transform as gt
    import se3_generate
    N, H, W = 3, 40, 128
    g = torch.Generator().manual_seed(4)
    inv = (torch.rand(N, 1, H, W, generator=g) * 0.3 + 0.02).cuda()
    T_s = torch.tensor([0.0, 0.0, 0.0, 0.5, 0.0, 0.0]).view(1, 6, 1, 1).repeat(N, 1, 1, 1).cuda()
    T_t = (torch.randn(N, 6, 1, 1, generator=g) * torch.tensor([0.01, 0.01, 0.01, 0.05, 0.02, 0.3]).view(1, 6, 1, 1)).cuda()
    K = torch.tensor([0.58 * W, 1.92 * H, 0.49 * W, 0.49 * H]).view(1, 4, 1, 1).repeat(N, 1, 1, 1).cuda()
    L2, R1, R2 = [torch.rand(N, 3, H, W, generator=g).cuda() for _ in range(3)]

    def single(T, src, d):
        SE3 = se3_generate.generate_se3(T)
        depth = (d + 1e-4).pow(-1)
        proj = gt.pin_hole_project(gt.geo_transform(depth, SE3, K), K)
        return gt.abs_loss(gt.inverse_warp(src, proj), R2)

    a = inv.clone().requires_grad_(True)
    ta, tb = T_s.clone().requires_grad_(True), T_t.clone().requires_grad_(True)
    e_lr, e_r12 = gt.two_view_warp_errors(a, ta, tb, K, L2, R1, R2)
    (e_lr + e_r12).backward()
    b = inv.clone().requires_grad_(True)
    sa, sb = T_s.clone().requires_grad_(True), T_t.clone().requires_grad_(True)
    r_lr, r_r12 = single(sa, L2, b), single(sb, R1, b)
    (r_lr + r_r12).backward()
    assert abs(e_lr.item() - r_lr.item()) <= 1e-6 * abs(r_lr.item()) and abs(e_r12.item() - r_r12.item()) <= 1e-6 * abs(r_r12.item())
    assert_close(npy(a.grad), npy(b.grad), tol=1e-6, what="d inv_depth")
    assert_close(npy(ta.grad), npy(sa.grad), tol=1e-5, what="d T_R2L")
    assert_close(npy(tb.grad), npy(sb.grad), tol=1e-5, what="d T_2to1")
