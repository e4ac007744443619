"""REF_CUDA probe: how far is the reference's torch operator sequence run with torch-CUDA eager from
the same sequence on torch-CPU (the parity target of this repo), and which arithmetic forms does torch-CUDA use?
Prints, for a KITTI-shaped batch: bitwise agreement of P = K @ pose_mat, of the sampling grid, of the warped image and of
the validity mask between torch-CPU, torch-CUDA and libdvf_b200; and which un-normalisation form grid_sample's CUDA kernel
matches bit for bit.  usage: python profiles/ref_cuda_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "depth-vo-feat_b200"))
import numpy as np, torch, torch.nn.functional as F
from dvf_b200 import ops, synthetic as syn
from oracle import torch_port as tp

B, H, W = 16, 128, 416
d = syn.stereo_temporal_batch(B, H, W, seed=7)
def run(dev):
    t = {k: v.to(dev) for k, v in d.items()}
    P = t["intrinsics"] @ tp.pose_matrix(t["T_2to1"])
    w = tp.warp(t["img_R1"], t["depth"], t["T_2to1"], t["intrinsics"], t["intrinsics_inv"])
    return P.cpu(), w.cpu()
P_cpu, w_cpu = run("cpu")
P_gpu, w_gpu = run("cuda")
t = {k: v.cuda() for k, v in d.items()}
import inverse_warp as iw
w_dvf = iw.inverse_warp(t["img_R1"], t["depth"], t["T_2to1"], t["intrinsics"], t["intrinsics_inv"]).cpu()
eq = lambda a, b: float((a.view(torch.int32) == b.view(torch.int32)).float().mean())
v = lambda w: (w != 0).any(1)
print(f"P = K @ pose_mat      torch-CUDA vs torch-CPU: {eq(P_gpu, P_cpu) * 100:.2f} % of entries bit-identical")
print(f"warped image          torch-CUDA vs torch-CPU: {eq(w_gpu, w_cpu) * 100:.3f} % of values bit-identical, "
      f"max |diff| {float((w_gpu - w_cpu).abs().max()):.3e}, masks differ at {int((v(w_gpu) != v(w_cpu)).sum())} of {B * H * W} pixels")
print(f"warped image          libdvf_b200 vs torch-CPU: {eq(w_dvf, w_cpu) * 100:.3f} % bit-identical, masks differ at "
      f"{int((v(w_dvf) != v(w_cpu)).sum())} pixels")
# un-normalisation form of grid_sample on CUDA: feed a grid, recover the cell by sampling a ramp image
g = torch.Generator().manual_seed(1)
xn = (torch.rand(1, 1, 200000, 1, generator=g) * 2.2 - 1.1)
grid = torch.cat([xn, torch.zeros_like(xn)], -1).cuda()
ramp = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).cuda()
s = F.grid_sample(ramp, grid, padding_mode="border", align_corners=False).view(-1).cpu()   # = clipped ix
x = xn.view(-1)
forms = {"fma(x+1, W/2, -0.5)": torch.addcmul(torch.tensor(-0.5), x + 1, torch.tensor(W / 2.0)),   # rounded product + add on the CPU
         "((x+1)*W - 1) / 2": ((x + 1) * W - 1) / 2}
x64 = x.double()
forms["fma exact (fp64 then rounded)"] = ((x64 + 1).float().double() * (W / 2.0) - 0.5).float()
for name, ix in forms.items():
    ixc = ix.clamp(0, W - 1)
    inside = (ixc > 0) & (ixc < W - 1)
    print(f"grid_sample CUDA un-normalise == {name:32s}: {float((s[inside] == ixc[inside]).float().mean()) * 100:.3f} % of in-range samples")
