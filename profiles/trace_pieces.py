"""Experiment aid (needs a -DDVF_TRACE build, e.g. DVF_LIB_NAME=libexpT.so): where does the time of a CTA's FIRST piece go?
Per-CTA globaltimer stamps: 0 kernel start, 1 TMA issued, 2 matrices ready, 3 first chunk landed, 4 pixel loop done,
5 predecessor waited (PDL), 6 reduction + tickets done; slot 7 = pixels of that piece.
usage: DVF_LIB_NAME=libexpT.so python profiles/trace_pieces.py [--config C2] [--no-pdl]"""
import argparse, ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
ap = argparse.ArgumentParser(); ap.add_argument("--config", default="C2"); ap.add_argument("--no-pdl", action="store_true")
ap.add_argument("--overhead", type=int, default=0)
a = ap.parse_args(); sys.argv = [sys.argv[0]]
args = bench.parse(); args.config = a.config
from dvf_b200 import _lib, plan as _plan
if a.overhead:
    _orig = _plan.FusedLossPlan.__init__
    def _init(self, *x, **k):
        k.setdefault("piece_overhead", a.overhead); _orig(self, *x, **k)
    _plan.FusedLossPlan.__init__ = _init
dev = torch.device("cuda", 0); wl = bench.WORKLOADS[a.config]
Bl, Bg = bench.local_batch(wl, 1)
steps = bench.build_steps(wl, Bl, Bg, bench.make_inputs(wl, Bl, 1000), dev, args.sets, pdl=not a.no_pdl)
r = bench.Runner(steps, args.graph_steps, 1, False)
r.spin(200); r.run(64); torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.lib_path()); n = 592
buf = (ctypes.c_ulonglong * (8 * n))()
assert lib.dvf_debug_trace_read(buf, n) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(n, 8).astype(np.int64)
px = t[:, 7]; t0 = t[:, 0].min()
names = ["start->tma issued", "matrices", "first chunk wait", "pixel loop", "pdl wait", "reduce+tickets"]
d = np.diff(t[:, :7], axis=1) / 1e3
print(f"{a.config} pdl={not a.no_pdl}: CTA start spread {(t[:,0].max()-t0)/1e3:.1f} us, last first-piece end {(t[:,6].max()-t0)/1e3:.1f} us")
for i, nme in enumerate(names):
    print(f"  {nme:20s} median {np.median(d[:, i]):7.2f} us   p90 {np.percentile(d[:, i], 90):7.2f}   max {d[:, i].max():7.2f}")
upx = d[:, 3] / np.maximum(px, 1) * 256
print(f"  pixel loop per 256-px unit: median {np.median(upx):.2f} us (pieces of {np.median(px):.0f} px median)")
