"""Times every stand-alone operator of libdvf_b200.so (everything except the fused loss, see kernel_time.py) on
C2-shaped inputs (B=64, 3x128x416) with CUDA events; inputs rotate over enough sets to exceed the 126 MB L2.
Prints one table row per kernel: us/launch, algorithmic GB/s, fraction of the measured copy peak.
usage: python profiles/all_kernels_time.py [--peak GBps]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from dvf_b200 import ops, synthetic as syn

ap = argparse.ArgumentParser(); ap.add_argument("--peak", type=float, default=0.0); ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
peak = a.peak
if not peak:
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6545.0
dev = torch.device("cuda"); B, C, H, W = a.batch, 3, 128, 416; HW = H * W
SETS = 6


def timed(name, fns, nbytes, iters=200):
    """each call is captured into a CUDA graph so that the figure is device time, not Python / allocator time"""
    for f in fns: f()
    torch.cuda.synchronize()
    graphs = []
    for f in fns:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = f()
        graphs.append((g, keep))
    fns = [g.replay for g, _ in graphs]
    for f in fns: f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fns[i % len(fns)]()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    gbps = nbytes / us * 1e-3
    print(f"| {name} | {us:.1f} | {nbytes / 1e6:.1f} | {gbps:.0f} | {gbps / peak * 100:.0f} % |")


sets = []
for k in range(SETS):
    d = syn.stereo_temporal_batch(B, H, W, seed=100 + k)
    t = {n: v.to(dev) for n, v in d.items()}
    P = ops.pose_proj_fwd(t["T_R2L"].unsqueeze(1).contiguous(), t["intrinsics"], t["intrinsics_inv"], 1, "euler", [1.0])[1][0]
    t["P"] = P.reshape(B, 3, 4).contiguous()
    t["gout"] = torch.randn(B, C, H, W, device=dev)
    t["cam"] = ops.pixel2cam(t["depth"], t["intrinsics_inv"])
    t["expl"] = syn.explainability(B, 2, H, W, 7 + k).to(dev)
    t["T44"] = torch.eye(4, device=dev).repeat(B, 1, 1, 1); t["T44"][:, 0, :3, 3] = t["T_R2L"][:, :3]
    t["K4"] = torch.stack([t["intrinsics"][:, 0, 0], t["intrinsics"][:, 1, 1], t["intrinsics"][:, 0, 2], t["intrinsics"][:, 1, 2]], 1).reshape(B, 4, 1, 1).contiguous()
    t["pts"] = ops.GeoTransform.apply(t["depth"].unsqueeze(1), t["T44"], t["K4"])
    t["xy"] = ops.PinHoleProject.apply(t["pts"], t["K4"])
    t["g2"] = torch.randn(B, 2, H, W, device=dev); t["g3"] = torch.randn(B, 3, H, W, device=dev)
    t["se3"] = torch.randn(B, 6, 1, 1, device=dev) * 0.1
    sets.append(t)
print(f"B={B} {C}x{H}x{W}, {SETS} rotating input sets, peak {peak:.0f} GB/s\n| kernel | us | alg. MB | GB/s | of peak |\n|---|---|---|---|---|")
px = B * HW
from dvf_b200 import _lib
lib = _lib.load(); st = lambda: torch.cuda.current_stream().cuda_stream
timed("inverse_warp fwd (zeros)", [lambda t=t: ops.inverse_warp_fwd_P(t["img_R1"], t["depth"], t["P"], t["intrinsics_inv"], "zeros") for t in sets], px * (4 + 4 * C + 4 * C))
timed("inverse_warp fwd (border)", [lambda t=t: ops.inverse_warp_fwd_P(t["img_R1"], t["depth"], t["P"], t["intrinsics_inv"], "border") for t in sets], px * (4 + 4 * C + 4 * C))
timed("inverse_warp bwd (d depth, d P)", [lambda t=t: ops.inverse_warp_bwd_P(t["gout"], t["img_R1"], t["depth"], t["P"], t["intrinsics_inv"], "zeros", need_gimg=False) for t in sets], px * (4 + 4 * C + 4 * C + 4))
timed("inverse_warp bwd (+ d img scatter)", [lambda t=t: ops.inverse_warp_bwd_P(t["gout"], t["img_R1"], t["depth"], t["P"], t["intrinsics_inv"], "zeros", need_gimg=True) for t in sets], px * (4 + 4 * C + 4 * C + 4 + 8 * C), iters=60)
timed("pixel2cam", [lambda t=t: ops.pixel2cam(t["depth"], t["intrinsics_inv"]) for t in sets], px * 16)
R = [t["P"][:, :, :3].contiguous() for t in sets]; T = [t["P"][:, :, 3:].contiguous() for t in sets]
timed("cam2pixel", [lambda t=t, r=r, tr=tr: ops.cam2pixel(t["cam"], r, tr, "zeros") for t, r, tr in zip(sets, R, T)], px * 20)
sizes = [(H >> s, W >> s) for s in range(4)]
timed("area_pyramid (4 levels)", [lambda t=t: ops.area_pyramid(t["img_R2"], sizes) for t in sets], px * C * 4 * (1 + 1 / 4 + 1 / 16 + 1 / 64))
dl = [[syn.depth(B, h, w, 3 + k).unsqueeze(1).to(dev) for (h, w) in sizes] for k in range(SETS)]
def reg(fn, maps):
    m = [x.detach().requires_grad_(True) for x in maps]
    fn(m).backward()
def reg_abi(kind, maps):
    """the C entry alone (value + unit gradients), no autograd scaling pass"""
    from dvf_b200._lib import dvf_reg_level
    L = len(maps); levels = (dvf_reg_level * L)(); gs = [torch.empty_like(m) for m in maps]
    for l, (m, g) in enumerate(zip(maps, gs)):
        levels[l] = dvf_reg_level(m.data_ptr(), g.data_ptr(), m.numel() // (m.shape[-1] * m.shape[-2]), m.shape[-2], m.shape[-1], 1.0)
    out = torch.empty(1, device=dev); ws = ops.workspace(lib.dvf_reg_workspace_bytes(levels, L), dev, ("t", kind, L, maps[0].shape))
    fn = lib.dvf_smooth_loss if kind == "smooth" else lib.dvf_explainability_loss
    _lib.check(fn(levels, L, out.data_ptr(), ws.data_ptr(), ws.numel(), st()), kind)
    return gs, out
timed("smooth_loss value+grad (4 scales), C entry", [lambda m=m: reg_abi("smooth", m) for m in dl], px * (1 + 1 / 4 + 1 / 16 + 1 / 64) * 8)
timed("smooth_loss through autograd (+ scaling pass)", [lambda m=m: reg(ops.smooth_loss, m) for m in dl], px * (1 + 1 / 4 + 1 / 16 + 1 / 64) * 8, iters=60)
el = [[t["expl"]] for t in sets]
timed("explainability_loss value+grad, C entry", [lambda m=m: reg_abi("expl", m) for m in el], px * 2 * 8)
timed("explainability_loss through autograd", [lambda m=m: reg(ops.explainability_loss, m) for m in el], px * 2 * 8, iters=60)
timed("se3_exp fwd", [lambda t=t: ops.SE3Exp.apply(t["se3"]) for t in sets], B * (24 + 128))
timed("caffe geo_transform fwd", [lambda t=t: ops.GeoTransform.apply(t["depth"].unsqueeze(1), t["T44"], t["K4"]) for t in sets], px * 16)
timed("caffe pin_hole fwd", [lambda t=t: ops.PinHoleProject.apply(t["pts"], t["K4"]) for t in sets], px * 20)
timed("caffe inverse_warp fwd", [lambda t=t: ops.PixelWarp.apply(t["img_R1"], t["xy"]) for t in sets], px * (8 + 8 * C))
def cbwd(t):
    N = B
    gd = torch.empty(B, 1, H, W, device=dev); gT = torch.empty(B, 16, device=dev); gK = torch.empty(B, 4, device=dev)
    _lib.check(lib.dvf_caffe_geo_bwd(t["g3"].data_ptr(), t["depth"].data_ptr(), t["T44"].data_ptr(), t["K4"].data_ptr(), N, H, W, gd.data_ptr(), gT.data_ptr(), gK.data_ptr(), st()), "geo_bwd")
timed("caffe geo_transform bwd", [lambda t=t: cbwd(t) for t in sets], px * 20)
def pbwd(t):
    gp = torch.empty(B, 3, H, W, device=dev); gK = torch.empty(B, 4, device=dev)
    _lib.check(lib.dvf_caffe_pinhole_bwd(t["g2"].data_ptr(), t["pts"].data_ptr(), t["K4"].data_ptr(), B, H, W, gp.data_ptr(), gK.data_ptr(), st()), "ph_bwd")
timed("caffe pin_hole bwd", [lambda t=t: pbwd(t) for t in sets], px * 32)
def wbwd(t, gi):
    gu = torch.empty(B, C, H, W, device=dev) if gi else None; gxy = torch.empty(B, 2, H, W, device=dev)
    _lib.check(lib.dvf_caffe_warp_bwd(t["gout"].data_ptr(), t["img_R1"].data_ptr(), t["xy"].data_ptr(), B, C, H, W, None if gu is None else gu.data_ptr(), gxy.data_ptr(), st()), "w_bwd")
timed("caffe inverse_warp bwd (d coords)", [lambda t=t: wbwd(t, False) for t in sets], px * (8 + 8 * C + 8))
timed("caffe inverse_warp bwd (+ d img)", [lambda t=t: wbwd(t, True) for t in sets], px * (8 + 8 * C + 8 + 8 * C), iters=60)
