"""Can the 218 MB zero-fill of the accumulated source-map gradients (C4) hide behind the kernels if a COPY ENGINE does it?
A C4 step = image loss launch + feature loss launch.  Baseline: DVF_FLAG_ZERO_GSRC (cudaMemsetAsync in stream order right
before the feature kernel: a fill kernel, 49 us).  Variant: the maps of input set k are cleared by a device-to-device copy
from a zero buffer on a side stream as soon as set k's feature kernel has finished (its consumer would sit there), i.e.
three steps ahead of their next use, overlapping the loss kernels of the other sets.
usage: python profiles/c4_ce_fill.py"""
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
N_GRAPH, N = 16, 50


def measure(tag, side_fill, memset_side=False):
    steps = bench.build_steps(wl, Bl, Bg, host, dev, args.sets, pdl=True)
    main, side = torch.cuda.Stream(), torch.cuda.Stream()
    zeros = None
    if side_fill:
        for s in steps:
            s.plans[1].desc.flags &= ~_lib.FLAG_ZERO_GSRC
            for g in s.plans[1].gsrc[0]:
                g.zero_()
        zeros = [torch.zeros_like(g) for g in steps[0].plans[1].gsrc[0]]
    torch.cuda.synchronize()
    used, filled = {}, {}

    def body(i):
        k = i % len(steps)
        st = steps[k]
        st.plans[0].launch()
        if side_fill and k in filled:
            torch.cuda.current_stream().wait_event(filled.pop(k))
        st.plans[1].launch()
        if side_fill:
            ev = torch.cuda.Event(); ev.record(); used[k] = ev
            side.wait_event(ev)
            with torch.cuda.stream(side):
                for g, z in zip(st.plans[1].gsrc[0], zeros):
                    if memset_side:
                        g.zero_()
                    else:
                        g.copy_(z, non_blocking=True)
                e2 = torch.cuda.Event(); e2.record(side); filled[k] = e2

    with torch.cuda.stream(main):
        for i in range(len(steps)):
            body(i)
        main.wait_stream(side)
        torch.cuda.synchronize()
        used.clear(); filled.clear()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main, capture_error_mode="thread_local"):
            for i in range(N_GRAPH):
                body(i)
            if side_fill:
                torch.cuda.current_stream().wait_stream(side)
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(N):
            g.replay()
        e1.record(); torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1) / (N * N_GRAPH) * 1e3:.1f} us per C4 step", flush=True)
    if side_fill:   # the maps must be what the in-stream fill produces
        ref = bench.build_steps(wl, Bl, Bg, host, dev, 1, pdl=False)[0]
        ref.launch(); torch.cuda.synchronize()
        k0 = steps[0].plans[1]
        g.replay(); torch.cuda.synchronize()
        # after a whole graph every set was cleared again behind its last use: run set 0 once more by hand
        steps[0].plans[0].launch(); steps[0].plans[1].launch(); torch.cuda.synchronize()
        err = max(float((a - b).abs().max()) for a, b in zip(k0.gsrc[0], ref.plans[1].gsrc[0]))
        print(f"   max |d src - reference fill| = {err:.3e}", flush=True)


measure("in-stream memset (DVF_FLAG_ZERO_GSRC)", False)
measure("side-stream device-to-device copy of zeros", True)
measure("side-stream fill kernel (tensor.zero_)", True, memset_side=True)
