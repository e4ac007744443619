"""Tuning aid: time the feature-reconstruction loss (C4 shape: 64-ch maps at 32x104, V=2, gradients to all maps)
through the autograd entry (NCHW fp32 generic-C kernel)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from dvf_b200 import ops, synthetic as syn
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=128); ap.add_argument("--channels", type=int, default=64)
ap.add_argument("--no-map-grads", action="store_true")
a = ap.parse_args()
B, C, H, W, V = a.batch, a.channels, 32, 104, 2
dev = torch.device("cuda")
f = [x.to(dev) for x in syn.features(B, C, H, W, 3, n=3)]
depth = syn.depth(B, H, W, 4).to(dev)
pose = torch.stack([syn.pose(B, "kitti", 5), syn.pose(B, "stereo", 6)], 1).to(dev)
K, Kinv = [x.to(dev) for x in syn.intrinsics(B, H, W)]
def step():
    tg, s0, s1 = [x.detach().requires_grad_(not a.no_map_grads) for x in f]
    d = depth.detach().requires_grad_(True); p = pose.detach().requires_grad_(True)
    loss, _ = ops.fused_photo_loss([tg], [[s0, s1]], [d], p, K, Kinv)
    loss.backward()
    return loss
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 50; e0.record()
for _ in range(N): step()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / N * 1e3
wpx = B * V * H * W
bytes_alg = wpx * (C * 4 * (1 if a.no_map_grads else 2)) + (wpx // V) * (8 + C * 4 * (1 if a.no_map_grads else 2))
print(f"features B={B} C={C} map_grads={not a.no_map_grads}: {us:.1f} us/step (autograd entry), {wpx / us * 1e-3:.2f} G wpx/s, {bytes_alg / us * 1e-3:.0f} GB/s algorithmic")
# channels-last variants (NHWC kernel)
for tdt in (torch.float32, torch.bfloat16):
    g = [x.to(tdt).contiguous(memory_format=torch.channels_last) for x in f]
    def step2():
        tg, s0, s1 = [x.detach().requires_grad_(not a.no_map_grads) for x in g]
        d = depth.detach().requires_grad_(True); p = pose.detach().requires_grad_(True)
        loss, _ = ops.fused_photo_loss([tg], [[s0, s1]], [d], p, K, Kinv)
        loss.backward()
    for _ in range(5): step2()
    torch.cuda.synchronize(); e0.record()
    for _ in range(N): step2()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / N * 1e3
    e = 2 if tdt == torch.bfloat16 else 4
    bytes_alg = wpx * (C * e + (0 if a.no_map_grads else C * 4)) + (wpx // V) * (8 + C * e + (0 if a.no_map_grads else C * 4))
    print(f"features NHWC {tdt} B={B} C={C} map_grads={not a.no_map_grads}: {us:.1f} us/step, {wpx / us * 1e-3:.2f} G wpx/s, {bytes_alg / us * 1e-3:.0f} GB/s algorithmic")
