"""Profiling aid: a few launches of the regulariser C entries (smooth / explainability) on C2-shaped maps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from dvf_b200 import ops, synthetic as syn, _lib
from dvf_b200._lib import dvf_reg_level
lib = _lib.load(); dev = torch.device("cuda"); B, H, W = 64, 128, 416
def run(kind, maps):
    L = len(maps); levels = (dvf_reg_level * L)(); gs = [torch.empty_like(m) for m in maps]
    for l, (m, g) in enumerate(zip(maps, gs)):
        levels[l] = dvf_reg_level(m.data_ptr(), g.data_ptr(), m.numel() // (m.shape[-1] * m.shape[-2]), m.shape[-2], m.shape[-1], 1.0)
    out = torch.empty(1, device=dev); ws = ops.workspace(lib.dvf_reg_workspace_bytes(levels, L), dev, ("t", kind, L, maps[0].shape))
    fn = lib.dvf_smooth_loss if kind == "smooth" else lib.dvf_explainability_loss
    _lib.check(fn(levels, L, out.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), kind)
    return out
d = [syn.depth(B, H >> s, W >> s, 3 + s).unsqueeze(1).to(dev) for s in range(4)]
e = [syn.explainability(B, 2, H, W, 7).to(dev)]
for _ in range(3):
    run("smooth", d); run("expl", e)
torch.cuda.synchronize()
