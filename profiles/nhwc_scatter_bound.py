"""How much of the channels-last feature kernel is the source-map gradient scatter?  Times the C4 feature-loss launch
(B=128, 64-channel bf16 maps at 32x104, V=2; bench.py's C4 inputs, 4 rotating sets, graphs of chained launches)
  full      : d tgt (bf16) + d src (fp32 red.v4 into zero-filled maps, fill included)
  no-fill   : the same launch without the 218 MB zero-fill (numbers wrong, time only)
  no-scatter: d src not requested (no reductions, no fill) -- the bound of ANY rearrangement of the scatter
  no-maps   : neither map gradient (d depth / d pose only)
usage: python profiles/nhwc_scatter_bound.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from dvf_b200 import _lib

sys.argv = [sys.argv[0]]
args = bench.parse(); args.config = "C4"
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
wl = bench.WORKLOADS["C4"]
Bl, Bg = bench.local_batch(wl, 1, 0)
host = bench.make_inputs(wl, Bl, 1000)


def run(tag, mutate):
    steps = bench.build_steps(wl, Bl, Bg, host, dev, args.sets, pdl=True)
    for s in steps:
        s.plans = s.plans[1:]            # the feature plan alone
        mutate(s.plans[0])
    r = bench.Runner(steps, args.graph_steps, 1, False)
    r.spin(200)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 800
    torch.cuda.synchronize(); e0.record(); r.run(n); e1.record(); torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per launch", flush=True)


def no_fill(p):
    p.desc.flags &= ~_lib.FLAG_ZERO_GSRC


def no_scatter(p):
    no_fill(p)
    for l in range(p.L):
        for v in range(p.V):
            p.levels[l].gsrc[v] = None


def no_maps(p):
    no_scatter(p)
    for l in range(p.L):
        p.levels[l].gtgt = None


run("full", lambda p: None)
run("no-fill", no_fill)
run("no-scatter", no_scatter)
run("no-maps", no_maps)
