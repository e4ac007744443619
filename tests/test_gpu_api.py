"""GPU tests of the API surface around the fused kernels: differentiable stand-alone pixel2cam / cam2pixel
(inverse_warp.py:26-74), descriptor options of ABI v2 (programmatic dependent launch, upstream scalar, sharded means,
NaN flag), autograd behaviour of the drop-in (retain_graph, two schedules)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import RTOL_F32, assert_close

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


# ---- the reference's two helper functions, restated with the same torch operators (inverse_warp.py:26-40, :43-74) ----
def ref_pixel2cam(depth, Kinv):
    b, h, w = depth.size()
    i_range = torch.arange(0, h).view(1, h, 1).expand(1, h, w).type_as(depth)
    j_range = torch.arange(0, w).view(1, 1, w).expand(1, h, w).type_as(depth)
    ones = torch.ones(1, h, w).type_as(depth)
    pix = torch.stack((j_range, i_range, ones), dim=1).expand(b, 3, h, w).reshape(b, 3, -1)
    return (Kinv @ pix).reshape(b, 3, h, w) * depth.unsqueeze(1)


def ref_cam2pixel(cam, rot, tr, padding_mode):
    b, _, h, w = cam.size()
    flat = cam.reshape(b, 3, -1)
    pc = rot @ flat if rot is not None else flat
    if tr is not None:
        pc = pc + tr
    X, Y, Z = pc[:, 0], pc[:, 1], pc[:, 2].clamp(min=1e-3)
    Xn = 2 * (X / Z) / (w - 1) - 1
    Yn = 2 * (Y / Z) / (h - 1) - 1
    if padding_mode == "zeros":
        Xm = ((Xn > 1) + (Xn < -1)).detach()
        Xn[Xm] = 2
        Ym = ((Yn > 1) + (Yn < -1)).detach()
        Yn[Ym] = 2
    return torch.stack([Xn, Yn], dim=2).reshape(b, h, w, 2)


def test_pixel2cam_autograd(ops, syn):
    import inverse_warp as iw
    B, H, W = 3, 30, 52
    depth = syn.depth(B, H, W, 5)
    _, Kinv = syn.intrinsics(B, H, W)
    gout = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(1))
    d_ref = depth.clone().requires_grad_(True)
    c_ref = ref_pixel2cam(d_ref, Kinv)
    c_ref.backward(gout)
    d_gpu = depth.cuda().requires_grad_(True)
    c_gpu = iw.pixel2cam(d_gpu, Kinv.cuda())
    c_gpu.backward(gout.cuda())
    assert np.array_equal(npy(c_gpu), c_ref.detach().numpy()), "pixel2cam forward: bit-identical to torch-CPU"
    assert np.array_equal(npy(d_gpu.grad), d_ref.grad.numpy()), "pixel2cam backward: (g * ray).sum(1) in the same order"


@pytest.mark.parametrize("padding", ["zeros", "border"])
@pytest.mark.parametrize("parts", ["rot+tr", "rot", "tr", "none"])
def test_cam2pixel_autograd(ops, syn, padding, parts):
    import inverse_warp as iw
    B, H, W = 2, 24, 80
    d = syn.stereo_temporal_batch(B, H, W, seed=11, temporal="large")
    K, Kinv = d["intrinsics"], d["intrinsics_inv"]
    cam = ref_pixel2cam(d["depth"], Kinv).detach()
    from oracle import torch_port as tp
    P = (K @ tp.pose_matrix(d["T_2to1"])).detach()
    rot = P[:, :, :3].contiguous() if "rot" in parts else None
    tr = P[:, :, 3:].contiguous() if "tr" in parts else None
    if rot is None:   # without a projection the points must already be pixel-like to land inside the image
        cam = (K @ cam.reshape(B, 3, -1)).reshape(B, 3, H, W).contiguous()
    gout = torch.randn(B, H, W, 2, generator=torch.Generator().manual_seed(2))

    def run(mod_cam2pixel, dev):
        c = cam.detach().clone().to(dev).requires_grad_(True)
        r = None if rot is None else rot.detach().clone().to(dev).requires_grad_(True)
        t = None if tr is None else tr.detach().clone().to(dev).requires_grad_(True)
        g = mod_cam2pixel(c, r, t, padding)
        g.backward(gout.to(dev))
        return g, c.grad, None if r is None else r.grad, None if t is None else t.grad

    g_ref, gc_ref, gr_ref, gt_ref = run(ref_cam2pixel, "cpu")
    g_gpu, gc_gpu, gr_gpu, gt_gpu = run(iw.cam2pixel, "cuda")
    assert np.array_equal(npy(g_gpu), g_ref.detach().numpy()), "cam2pixel forward: bit-identical grid"
    assert_close(npy(gc_gpu), gc_ref.numpy(), what="d cam_coords")
    if rot is not None:
        assert_close(npy(gr_gpu), gr_ref.numpy(), what="d proj_c2p_rot")
    if tr is not None:
        assert gt_gpu.shape == tr.shape
        assert_close(npy(gt_gpu), gt_ref.numpy(), what="d proj_c2p_tr")


# ---- descriptor options ----------------------------------------------------------------------------------------------
def _plans(ops, syn, B, H, W, L, V, n_sets, with_expl=False, **kw):
    from dvf_b200.plan import FusedLossPlan
    sizes = [(H >> s, W >> s) for s in range(L)]
    ds = [float(1 << s) for s in range(L)]
    out = []
    for k in range(n_sets):
        d = syn.stereo_temporal_batch(B, H, W, seed=70 + k)
        tg = ops.area_pyramid(d["img_R2"].cuda(), sizes)
        srcs = [ops.area_pyramid(d[n].cuda(), sizes) for n in ("img_R1", "img_L2")][:V]
        depths = [syn.depth(B, h, w, 80 + 10 * k + i).cuda() for i, (h, w) in enumerate(sizes)]
        expl = [syn.explainability(B, V, h, w, 90 + i).cuda() for i, (h, w) in enumerate(sizes)] if with_expl else None
        pose = torch.stack([d["T_2to1"], d["T_R2L"]][:V], 1).contiguous().cuda()
        out.append(FusedLossPlan(tg, [[sp[l] for sp in srcs] for l in range(L)], depths, pose, d["intrinsics"].cuda(),
                                 d["intrinsics_inv"].cuda(), expl_levels=expl, downscales=ds, **kw))
    return out


def _snapshot(p):
    return [npy(p.terms).copy(), npy(p.gpose).copy()] + [npy(g).copy() for g in p.gdepth] + \
        ([npy(g).copy() for g in p.gexpl] if p.gexpl is not None else [])


@pytest.mark.parametrize("V,with_expl", [(1, False), (2, True)])
def test_pdl_chain_is_bit_identical(ops, syn, V, with_expl):
    """DVF_FLAG_PDL: back-to-back launches on disjoint buffers overlap; every result equals the plain launch bit for bit,
    eagerly and replayed from a CUDA graph."""
    B, H, W, L, S = 6, 64, 208, 3, 3
    plain = _plans(ops, syn, B, H, W, L, V, S, with_expl)
    chained = _plans(ops, syn, B, H, W, L, V, S, with_expl, pdl=True)
    for p in plain:
        p.launch()
    torch.cuda.synchronize()
    want = [_snapshot(p) for p in plain]
    for _ in range(3):
        for p in chained:
            p.launch()
    torch.cuda.synchronize()
    for p, w in zip(chained, want):
        assert all(np.array_equal(a, b) for a, b in zip(_snapshot(p), w)), "PDL launch differs from the plain launch"
    for p in chained:   # poison the outputs, then replay a captured chain
        p.terms.fill_(float("nan"))
        for g in p.gdepth:
            g.fill_(float("nan"))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(2):
                for p in chained:
                    p.launch()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    g.replay()
    torch.cuda.synchronize()
    for p, w in zip(chained, want):
        assert all(np.array_equal(a, b) for a, b in zip(_snapshot(p), w)), "captured PDL chain differs from the plain launch"


def test_pdl_chained_small_grid(ops, syn):
    """DVF_FLAG_PDL_CHAINED: grid sized for overlap (fewer, larger CTAs).  Per-pixel outputs stay bit-identical; sums over CTAs
    are folded in another order (1e-6)."""
    B, H, W, L, V, S = 6, 64, 208, 3, 2, 3
    plain = _plans(ops, syn, B, H, W, L, V, S, True)
    chained = _plans(ops, syn, B, H, W, L, V, S, True, pdl=True, pdl_chained=True)
    for p in plain:
        p.launch()
    for _ in range(3):
        for p in chained:
            p.launch()
    torch.cuda.synchronize()
    for p, q in zip(plain, chained):
        assert_close(npy(q.terms), npy(p.terms), tol=1e-6, what="terms")
        assert_close(npy(q.gpose), npy(p.gpose), tol=1e-6, what="gpose")
        for a, b in zip(p.gdepth + p.gexpl, q.gdepth + q.gexpl):
            assert np.array_equal(npy(a), npy(b)), "per-pixel gradients do not depend on the grid"


def test_upstream_scalar_and_sharded_mean(ops, syn):
    """dvf_loss_desc.upstream scales every gradient (not the terms); mean_batch makes a shard return its share of the
    global mean, so that shards add up to the un-sharded call (dvf_b200.dist)."""
    B, H, W, L, V = 4, 32, 104, 2, 2
    full = _plans(ops, syn, B, H, W, L, V, 1, True)[0]
    full.launch()
    up = torch.tensor([0.25], device="cuda")
    scaled = _plans(ops, syn, B, H, W, L, V, 1, True, upstream=up)[0]
    scaled.launch()
    torch.cuda.synchronize()
    a, b = _snapshot(full), _snapshot(scaled)
    assert np.array_equal(a[0], b[0]), "terms are not scaled"
    for x, y in zip(a[1:], b[1:]):
        assert_close(y, 0.25 * x, tol=2e-6, what="gradient * upstream")
    # shards: images [0,2) and [2,4) with mean_batch = 4
    from dvf_b200.plan import FusedLossPlan
    tg, srcs, depths, pose, K, Kinv, expl, _ = full.inputs
    ds = [1.0, 2.0]
    terms = torch.zeros_like(full.terms)
    for sl in (slice(0, 2), slice(2, 4)):
        c = lambda t: t[sl].contiguous()   # noqa: E731
        p = FusedLossPlan([c(t) for t in tg], [[c(s) for s in lv] for lv in srcs], [c(t) for t in depths], c(pose), c(K), c(Kinv),
                          expl_levels=[c(t) for t in expl], downscales=ds, global_batch=B)
        p.launch()
        torch.cuda.synchronize()
        terms += p.terms
        for l in range(L):
            assert_close(npy(p.gdepth[l]), npy(full.gdepth[l][sl]), tol=1e-6, what="shard gdepth")
        assert_close(npy(p.gpose), npy(full.gpose[sl]), tol=2e-6, what="shard gpose")
    assert_close(npy(terms), npy(full.terms), tol=2e-6, what="sum of shard terms")


def test_nan_flag(ops, syn):
    """nan_check=True: the kernel ORs bit l*V+v into ops.nan_flags() when terms[l*V+v] is NaN -- the reference asserts
    per view and scale with a device sync each (loss_functions_sfm.py:34)."""
    import loss_functions_sfm as sfm
    B, H, W = 2, 32, 104
    d = syn.stereo_temporal_batch(B, H, W, seed=3)
    t = {k: v.cuda() for k, v in d.items()}
    depths = [t["depth"].unsqueeze(1), syn.depth(B, H // 2, W // 2, 4).cuda().unsqueeze(1)]
    pose = torch.stack([t["T_2to1"], t["T_R2L"]], 1)
    flags = ops.nan_flags()
    flags.zero_()
    sfm.NAN_CHECK = True
    try:
        sfm.photometric_reconstruction_loss(t["img_R2"], [t["img_R1"], t["img_L2"]], t["intrinsics"], t["intrinsics_inv"], depths,
                                            [None, None], pose)
        assert int(flags.item()) == 0
        bad = [depths[0], depths[1].clone()]
        bad[1][0, 0, 3, 5] = float("nan")
        loss = sfm.photometric_reconstruction_loss(t["img_R2"], [t["img_R1"], t["img_L2"]], t["intrinsics"],
                                                   t["intrinsics_inv"], bad, [None, None], pose)
        assert torch.isnan(loss).item()
        assert int(flags.item()) == 0b1100, "level 1, both views"
        with pytest.raises(AssertionError):
            sfm.assert_no_nan()
    finally:
        sfm.NAN_CHECK = False
        flags.zero_()


def test_retain_graph_and_two_schedules(ops, syn):
    """backward can run twice (retain_graph=True) for image losses (gradients scaled out of place) and for feature losses
    (forward-only kernel in forward(), fused pass with the upstream scalar in backward())."""
    import loss_functions as lf
    B, H, W = 2, 16, 52
    for feature, Cc in ((False, 3), (True, 16)):
        d = syn.stereo_temporal_batch(B, H, W, seed=21, C=Cc, feature=feature)
        t = {k: v.cuda() for k, v in d.items()}
        req = ["depth", "T_2to1", "T_R2L"] + (["img_R2", "img_R1", "img_L2"] if feature else [])
        for k in req:
            t[k].requires_grad_(True)
        loss = lf.photometric_reconstruction_loss(t["img_R2"], t["img_R1"], t["img_L2"], t["depth"], t["T_2to1"], t["T_R2L"],
                                                  t["intrinsics"], t["intrinsics_inv"])
        (0.1 * loss).backward(retain_graph=True)
        first = {k: t[k].grad.clone() for k in req}
        for k in req:
            t[k].grad = None
        loss.backward()
        for k in req:
            assert_close(npy(first[k]), 0.1 * npy(t[k].grad), tol=RTOL_F32, what=f"{'feature' if feature else 'image'} d {k}")


def test_tensors_on_different_devices_are_rejected(ops, syn):
    from dvf_b200._lib import DvfError
    d = syn.stereo_temporal_batch(1, 16, 52, seed=1)
    with pytest.raises(DvfError):
        import inverse_warp as iw
        iw.inverse_warp(d["img_R1"], d["depth"].cuda(), d["T_2to1"].cuda(), d["intrinsics"].cuda(), d["intrinsics_inv"].cuda())


@pytest.mark.parametrize("shape", [(2, 3, 19, 37), (1, 1, 3, 3), (3, 3, 128, 416), (2, 8, 33, 65)])
def test_ssim_loss_vs_oracle_and_torch(ops, shape):
    """SSIM term (new functionality; parity unpinned, no SSIM in the reference): the CUDA kernel against the C oracle and
    against the torch restatement on the CPU, fp32, 1e-5."""
    from helpers import torch_ssim_loss
    from oracle import cpu_oracle as O
    O.build()
    B, Cc, H, W = shape
    g = torch.Generator().manual_seed(7)
    x = torch.rand(B, Cc, H, W, generator=g)
    y = (x + 0.3 * torch.rand(B, Cc, H, W, generator=g) - 0.1).clamp(0, 1)
    valid = (torch.rand(B, H, W, generator=g) > 0.05).to(torch.uint8)
    for v in (None, valid):
        yc = y.cuda().requires_grad_(True)
        loss = ops.ssim_loss(x.cuda(), yc, None if v is None else v.cuda())
        (2.0 * loss).backward()
        ol, og = O.ssim_loss(x.numpy(), y.numpy(), None if v is None else v.numpy())
        yd = y.clone().requires_grad_(True)
        tl = torch_ssim_loss(x, yd, v)
        tl.backward()
        assert abs(loss.item() - ol) <= RTOL_F32 * max(ol, 1e-6), (loss.item(), ol)
        assert abs(loss.item() - tl.item()) <= 2 * RTOL_F32 * max(tl.item(), 1e-6)
        if np.abs(og).max() > 0:
            assert_close(npy(yc.grad), 2.0 * og, tol=RTOL_F32, what="SSIM d/dy vs oracle")
            assert_close(npy(yc.grad), 2.0 * yd.grad.numpy(), tol=3 * RTOL_F32, what="SSIM d/dy vs torch")


def test_ssim_mixed_loss_runs_end_to_end(ops, syn):
    """loss_functions.photometric_ssim_reconstruction_loss: (1-alpha) L1 + alpha SSIM; gradients equal the sum of the parts."""
    import inverse_warp as iw
    import loss_functions as lf
    d = syn.stereo_temporal_batch(2, 32, 104, seed=5)
    t = {k: v.cuda() for k, v in d.items()}
    for k in ("depth", "T_2to1", "T_R2L"):
        t[k].requires_grad_(True)
    args = (t["img_R2"], t["img_R1"], t["img_L2"], t["depth"], t["T_2to1"], t["T_R2L"], t["intrinsics"], t["intrinsics_inv"])
    loss = lf.photometric_ssim_reconstruction_loss(*args, alpha=0.85)
    loss.backward()
    got = {k: t[k].grad.clone() for k in ("depth", "T_2to1", "T_R2L")}
    for k in got:
        t[k].grad = None
    l1 = lf.photometric_reconstruction_loss(*args)
    s = 0
    for src, pose in (("img_R1", "T_2to1"), ("img_L2", "T_R2L")):
        w = iw.inverse_warp(t[src], t["depth"], t[pose], t["intrinsics"], t["intrinsics_inv"])
        s = s + ops.ssim_loss(t["img_R2"], w, (w.detach() != 0).any(1).to(torch.uint8))
    ref = 0.15 * l1 + 0.85 * s
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-6 * abs(ref.item())
    assert 0.0 < s.item() < 2.0
    for k in got:
        assert_close(npy(got[k]), npy(t[k].grad), tol=1e-6, what=k)


def test_producer_glue_matches_separate_torch_passes(ops, syn):
    """DVF_FLAG_DISPARITY / img_scale: the reciprocal of unsupervise.py:99 and the 0.004 * img scalings of
    :101 folded into the launch give the loss and gradients of the separate torch passes (which are exact fp32 element-wise
    ops): loss and d/d disparity at 1e-5, validity-sensitive terms identical."""
    import loss_functions as lf
    import loss_functions_sfm as sfm
    B, H, W = 3, 32, 104
    d = syn.stereo_temporal_batch(B, H, W, seed=31)
    t = {k: v.cuda() for k, v in d.items()}
    raw = {k: (t[k] / 0.004).contiguous() for k in ("img_R2", "img_R1", "img_L2")}     # uint8-range images
    inv = (1.0 / t["depth"] - 1e-4).unsqueeze(1).contiguous()
    a = inv.clone().requires_grad_(True)
    pa, pb = t["T_2to1"].clone().requires_grad_(True), t["T_R2L"].clone().requires_grad_(True)
    depth = (1 / (a + 1e-4)).squeeze(1)
    ref = lf.photometric_reconstruction_loss(0.004 * raw["img_R2"], 0.004 * raw["img_R1"], 0.004 * raw["img_L2"], depth, pa, pb,
                                             t["intrinsics"], t["intrinsics_inv"])
    ref.backward()
    b = inv.clone().requires_grad_(True)
    qa, qb = t["T_2to1"].clone().requires_grad_(True), t["T_R2L"].clone().requires_grad_(True)
    got = lf.photometric_reconstruction_loss_fused_inputs(raw["img_R2"], raw["img_R1"], raw["img_L2"], b, qa, qb,
                                                          t["intrinsics"], t["intrinsics_inv"])
    got.backward()
    assert abs(got.item() - ref.item()) <= 1e-6 * abs(ref.item())
    assert_close(npy(b.grad), npy(a.grad), tol=RTOL_F32, what="d inv_depth")
    assert_close(npy(qa.grad), npy(pa.grad), tol=RTOL_F32, what="d T_2to1")
    assert_close(npy(qb.grad), npy(pb.grad), tol=RTOL_F32, what="d T_R2L")
    # multi-scale form (train.py:188, eps 0)
    disps = [(1.0 / syn.depth(B, H >> s, W >> s, 40 + s)).cuda().unsqueeze(1) for s in range(2)]
    pose = torch.stack([t["T_2to1"], t["T_R2L"]], 1)
    x = [v.clone().requires_grad_(True) for v in disps]
    r2 = sfm.photometric_reconstruction_loss(t["img_R2"], [t["img_R1"], t["img_L2"]], t["intrinsics"], t["intrinsics_inv"],
                                             [1 / v for v in x], [None, None], pose)
    r2.backward()
    y = [v.clone().requires_grad_(True) for v in disps]
    g2 = sfm.photometric_reconstruction_loss(t["img_R2"], [t["img_R1"], t["img_L2"]], t["intrinsics"], t["intrinsics_inv"], y,
                                             [None, None], pose, disparity_eps=0.0)
    g2.backward()
    assert abs(g2.item() - r2.item()) <= 1e-6 * abs(r2.item())
    for u, v in zip(x, y):
        assert_close(npy(v.grad), npy(u.grad), tol=RTOL_F32, what="d disparity (multi-scale)")


def test_fused_terms_exchange_single_rank(ops, syn):
    """dvf_loss_desc.peer_terms on ONE GPU (world size 1): the kernel's epilogue stores the loss terms into the exchange buffer
    obtained from dvf_b200.dist.PeerTerms (symmetric memory / CUDA IPC mapping); the N > 1 path is exercised by bench.py under
    torchrun, which checks it against an NCCL all-reduce (exchange_check_rel_err)."""
    import os
    import socket
    import torch.distributed as dist
    from dvf_b200.dist import PeerTerms
    created = False
    if not dist.is_initialized():
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        B, H, W, L, V, S = 4, 32, 104, 2, 2, 2
        pt = PeerTerms(S, L * V, torch.device("cuda", 0))
        assert pt.world == 1 and pt.how in ("symmetric_memory", "cuda_ipc")
        plans = []
        for k in range(S):
            plans += _plans(ops, syn, B, H, W, L, V, 1, True, peer_terms=pt.slot_ptrs(k), peer_rank=0)
        for p in plans:
            p.launch()
        torch.cuda.synchronize()
        for k, p in enumerate(plans):
            assert np.array_equal(npy(pt.gathered(k)[0]), npy(p.terms)), "terms stored by the kernel epilogue"
    finally:
        if created:
            dist.destroy_process_group()


# ---- REF_CUDA arithmetic profile: the reference run with torch-CUDA eager on this GPU is the oracle ------------
@pytest.fixture
def ref_cuda(ops):
    ops.ARITHMETIC = "ref_cuda"
    try:
        yield ops
    finally:
        ops.ARITHMETIC = "ref_cpu"


@pytest.mark.parametrize("padding", ["zeros", "border"])
@pytest.mark.parametrize("kind,shape", [("kitti", (8, 64, 208)), ("tiny", (8, 64, 208)), ("large", (8, 64, 208)),
                                        ("kitti", (4, 128, 416))])
def test_ref_cuda_profile_projection_and_warp(ref_cuda, syn, kind, shape, padding):
    """ops.ARITHMETIC = 'ref_cuda': P = K @ pose_vec2mat(pose) and the warped image carry the bits of the reference's torch
    operator sequence executed by torch-CUDA eager (oracle/torch_port.py on this GPU), where the default profile
    reproduces torch-CPU: libdevice sin / cos and FMA-chain tiny matmuls in the pose chain (DVF_ROT_REF_CUDA), division
    by the scalar w-1 as a reciprocal multiply and corner-difference bilinear weights per pixel (DVF_FLAG_REF_CUDA)."""
    import inverse_warp as iw
    from oracle import torch_port as tp
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W = shape
    d = syn.stereo_temporal_batch(B, H, W, seed=5, temporal=kind)
    t = {k: v.cuda() for k, v in d.items()}
    for pose in (t["T_2to1"], t["T_R2L"]):
        P_ref = t["intrinsics"] @ tp.pose_matrix(pose)
        _, P, _ = ref_cuda.pose_proj_fwd(pose, t["intrinsics"], None, 1, "euler", [1.0])
        assert np.array_equal(npy(P[0]), npy(P_ref)), "projection matrices: torch-CUDA's bits"
        w_ref = tp.warp(t["img_R1"], t["depth"], pose, t["intrinsics"], t["intrinsics_inv"], padding_mode=padding)
        w = iw.inverse_warp(t["img_R1"], t["depth"], pose, t["intrinsics"], t["intrinsics_inv"], padding_mode=padding)
        assert np.array_equal(npy(w), npy(w_ref)), "warped image: torch-CUDA's bits"
    # and the default profile is NOT torch-CUDA's (otherwise this test would prove nothing)
    ref_cuda.ARITHMETIC = "ref_cpu"
    w_cpu_profile = iw.inverse_warp(t["img_R1"], t["depth"], t["T_2to1"], t["intrinsics"], t["intrinsics_inv"], padding_mode=padding)
    w_ref = tp.warp(t["img_R1"], t["depth"], t["T_2to1"], t["intrinsics"], t["intrinsics_inv"], padding_mode=padding)
    assert not np.array_equal(npy(w_cpu_profile), npy(w_ref))
    assert float((w_cpu_profile - w_ref).abs().max()) <= 2e-4


@pytest.mark.parametrize("pad", ["zeros", "border"])
def test_ref_cuda_profile_golden(ref_cuda, syn, pad):
    """the committed torch-CUDA fixture (tests/golden/ref_cuda_warp.npz, also what pins the CPU oracle's restatement):
    projection matrices from the pose and warped images through the drop-in, bit for bit"""
    import inverse_warp as iw
    from helpers import golden
    g = golden("ref_cuda_warp")
    d = syn.stereo_temporal_batch(int(g["B"]), int(g["H"]), int(g["W"]), seed=int(g["seed"]))
    t = {k: v.cuda() for k, v in d.items()}
    for tag, pose in (("temporal", "T_2to1"), ("stereo", "T_R2L")):
        _, P, _ = ref_cuda.pose_proj_fwd(t[pose], t["intrinsics"], None, 1, "euler", [1.0])
        assert np.array_equal(npy(P[0]), g["P_" + tag])
        w = iw.inverse_warp(t["img_R1"], t["depth"], t[pose], t["intrinsics"], t["intrinsics_inv"], padding_mode=pad)
        assert np.array_equal(npy(w), g[f"warped_{tag}_{pad}"])


def test_ref_cuda_profile_loss_through_the_dropin(ref_cuda, syn):
    """the multi-scale loss and its gradients under the torch-CUDA profile (three-launch form) against torch-CUDA autograd"""
    import loss_functions_sfm as sfm
    from oracle import torch_port as tp
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W, L = 4, 64, 208, 3
    d = syn.stereo_temporal_batch(B, H, W, seed=8)
    t = {k: v.cuda() for k, v in d.items()}
    pose0 = torch.stack([t["T_2to1"], t["T_R2L"]], 1)
    depth0 = [torch.nn.functional.avg_pool2d(t["depth"].unsqueeze(1), 1 << s).contiguous() for s in range(L)]
    expl0 = [torch.sigmoid(torch.randn(B, 2, H >> s, W >> s, device="cuda", generator=torch.Generator("cuda").manual_seed(s)))
             for s in range(L)]

    def run(fn):
        pose = pose0.clone().requires_grad_(True)
        depths = [x.clone().requires_grad_(True) for x in depth0]
        masks = [x.clone().requires_grad_(True) for x in expl0]
        loss = fn(t["img_R2"], [t["img_R1"], t["img_L2"]], t["intrinsics"], t["intrinsics_inv"], depths, masks, pose)
        loss.backward()
        return loss.detach(), pose.grad, [x.grad for x in depths], [x.grad for x in masks]

    l_ref, gp_ref, gd_ref, ge_ref = run(lambda *a: tp.loss_multi_scale(*a))
    l_gpu, gp, gd, ge = run(sfm.photometric_reconstruction_loss)
    assert abs(float(l_gpu) - float(l_ref)) <= RTOL_F32 * abs(float(l_ref))
    assert_close(npy(gp), npy(gp_ref), what="d pose")
    # per-pixel gradients: torch-CUDA's own rounding of the chain differs from torch-CPU's (which the kernels follow) at the
    # 1e-5 level -- measured 1.7e-5 of max|g| on this input
    for a, b in zip(gd + ge, gd_ref + ge_ref):
        assert_close(npy(a), npy(b), tol=1e-4, what="d depth / d mask")
