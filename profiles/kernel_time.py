"""Tuning aid: time the fused loss launch(es) of a bench workload with CUDA events (same graphs as bench.py).
usage: [DVF_LIB_NAME=libexp.so] python profiles/kernel_time.py [--config C2] [--no-pdl] [--batch B] [--ctas N] [--overhead U]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2"); ap.add_argument("--no-pdl", action="store_true"); ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--ctas", type=int, default=0); ap.add_argument("--overhead", type=int, default=0); ap.add_argument("--n", type=int, default=3200)
a = ap.parse_args()
sys.argv = [sys.argv[0]]
args = bench.parse(); args.config = a.config; args.no_pdl = a.no_pdl; args.batch = a.batch
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
wl = bench.WORKLOADS[a.config]
Bl, Bg = bench.local_batch(wl, 1, a.batch)
host = bench.make_inputs(wl, Bl, 1000)
if a.ctas or a.overhead:   # tuning fields of the descriptor
    from dvf_b200 import plan as _plan
    _orig = _plan.FusedLossPlan.__init__
    def _init(self, *x, **k):
        k.setdefault("ctas_per_sm", a.ctas); k.setdefault("piece_overhead", a.overhead); _orig(self, *x, **k)
    _plan.FusedLossPlan.__init__ = _init
steps = bench.build_steps(wl, Bl, Bg, host, dev, args.sets, pdl=not a.no_pdl)
r = bench.Runner(steps, args.graph_steps, 1, False)
r.spin(300)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); r.run(a.n); e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / a.n * 1e3
print(f"{a.config} B={Bl} pdl={not a.no_pdl} ctas={a.ctas} overhead={a.overhead}: {us:.2f} us/step, {steps[0].warped_px / us * 1e-3:.1f} G wpx/s, "
      f"{steps[0].bytes / us * 1e-3:.0f} GB/s algorithmic, lib={os.environ.get('DVF_LIB_NAME', 'default')}")
