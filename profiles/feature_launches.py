"""ncu target: ONE feature-loss step (C4 shape) per dtype through the autograd entry, so that a launch list shows
every kernel a step launches (loss kernel, fills, scaling pass, dtype conversions)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from dvf_b200 import ops, synthetic as syn
B, C, H, W, V = 128, 64, 32, 104, 2
dev = torch.device("cuda")
f = syn.features(B, C, H, W, 3, n=3)
depth = syn.depth(B, H, W, 4).to(dev)
pose = torch.stack([syn.pose(B, "kitti", 5), syn.pose(B, "stereo", 6)], 1).to(dev)
K, Kinv = [x.to(dev) for x in syn.intrinsics(B, H, W)]
for tdt in (torch.float32, torch.bfloat16):
    g = [x.to(dev).to(tdt).contiguous(memory_format=torch.channels_last) for x in f]
    for it in range(3):
        if it == 2:
            torch.cuda.synchronize(); print("=== step", tdt, flush=True)
            torch.cuda.nvtx.range_push(f"step_{tdt}")
        tg, s0, s1 = [x.detach().requires_grad_(True) for x in g]
        d = depth.detach().requires_grad_(True); p = pose.detach().requires_grad_(True)
        loss, _ = ops.fused_photo_loss([tg], [[s0, s1]], [d], p, K, Kinv)
        (0.5 * loss).backward()
        if it == 2:
            torch.cuda.synchronize(); torch.cuda.nvtx.range_pop()
