"""Tuning aid: time the fused loss kernel on the C2 workload (optionally forward-only) with CUDA events.
usage: python profiles/kernel_time.py [--no-grad] [--views V] [--expl] [--batch B]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from dvf_b200 import ops, synthetic as syn
from dvf_b200.plan import FusedLossPlan
ap = argparse.ArgumentParser(); ap.add_argument("--no-grad", action="store_true"); ap.add_argument("--views", type=int, default=1)
ap.add_argument("--expl", action="store_true"); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--sets", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda"); H, W, L = bench.H, bench.W, bench.LEVELS
sizes = [(H >> s, W >> s) for s in range(L)]; ds = [float(1 << s) for s in range(L)]
host = bench.make_inputs(a.batch, 1000)
plans = []
for k in range(a.sets):
    roll = lambda t: torch.roll(t, k, 0).contiguous().to(dev)
    tg = ops.area_pyramid(roll(host["tgt"]), sizes); sr = ops.area_pyramid(roll(host["src"]), sizes)
    pose = roll(host["pose"]).repeat(1, a.views, 1).contiguous()
    if a.views > 1: pose[:, 1:] = syn.pose(a.batch, "kitti", 5).to(dev).unsqueeze(1)
    expl = [syn.explainability(a.batch, a.views, h, w, 7 + i).to(dev) for i, (h, w) in enumerate(sizes)] if a.expl else None
    plans.append(FusedLossPlan(tg, [[s] * a.views for s in sr], [roll(x) for x in host["depths"]], pose, roll(host["K"]),
                               roll(host["Kinv"]), expl_levels=expl, downscales=ds, need_grad=not a.no_grad))
graphs = []
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for p in plans: graphs.append(p.capture())
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
for i in range(300): graphs[i % len(graphs)].replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); N = 3000
for i in range(N): graphs[i % len(graphs)].replay()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / N * 1e3
print(f"grad={not a.no_grad} views={a.views} expl={a.expl} B={a.batch}: {us:.2f} us/launch, {plans[0].warped_px / us * 1e-3:.1f} G wpx/s, "
      f"{plans[0].algorithmic_bytes() / us * 1e-3:.0f} GB/s algorithmic, lib={os.environ.get('DVF_LIB_NAME','default')}")
