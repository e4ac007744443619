"""cProfile of the sfm drop-in loss on tiny maps (host-bound): where the Python time of a call goes."""
import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from dvf_b200 import ops, synthetic as syn
import loss_functions_sfm as sfm
B, H, W, L = 2, 16, 52, 4
dev = torch.device("cuda")
d = syn.stereo_temporal_batch(B, H, W, seed=1)
t = {k: v.to(dev) for k, v in d.items()}
depths = [syn.depth(B, H >> s, W >> s, 5 + s).unsqueeze(1).to(dev) for s in range(L)]
pose = t["T_R2L"].unsqueeze(1).contiguous()
def step():
    dl = [x.detach().requires_grad_(True) for x in depths]; p = pose.detach().requires_grad_(True)
    loss = sfm.photometric_reconstruction_loss(t["img_R2"], [t["img_L2"]], t["intrinsics"], t["intrinsics_inv"], dl, [None] * L, p)
    loss.backward()
for _ in range(30): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:7000])
