"""Host-side cost of the drop-in entries: tiny maps (GPU time negligible), wall clock per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from dvf_b200 import ops, synthetic as syn
import loss_functions_sfm as sfm, loss_functions as lf
import argparse
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=2); ap.add_argument("--height", type=int, default=16); ap.add_argument("--width", type=int, default=52)
a = ap.parse_args()
B, H, W, L = a.batch, a.height, a.width, 4
dev = torch.device("cuda")
d = syn.stereo_temporal_batch(B, H, W, seed=1)
t = {k: v.to(dev) for k, v in d.items()}
depths = [syn.depth(B, H >> s, W >> s, 5 + s).unsqueeze(1).to(dev) for s in range(L)]
pose = t["T_R2L"].unsqueeze(1).contiguous()
def step_sfm():
    dl = [x.detach().requires_grad_(True) for x in depths]; p = pose.detach().requires_grad_(True)
    loss = sfm.photometric_reconstruction_loss(t["img_R2"], [t["img_L2"]], t["intrinsics"], t["intrinsics_inv"], dl, [None] * L, p)
    loss.backward()
def step_fwd_only():
    with torch.no_grad():
        sfm.photometric_reconstruction_loss(t["img_R2"], [t["img_L2"]], t["intrinsics"], t["intrinsics_inv"], depths, [None] * L, pose)
def step_smooth():
    dl = [x.detach().requires_grad_(True) for x in depths]
    lf.smooth_loss(dl, 2.0).backward()
for name, fn in [("sfm photometric loss, 4 scales, fwd+bwd", step_sfm), ("same, forward only (no_grad)", step_fwd_only), ("smooth_loss fwd+bwd", step_smooth)]:
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter(); N = 300
    for _ in range(N): fn()
    torch.cuda.synchronize(); us = (time.perf_counter() - t0) / N * 1e6
    print(f"{name}: {us:.0f} us per call (host-bound: maps of {H}x{W})")
