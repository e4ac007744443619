#!/usr/bin/env python
"""Benchmark of the fused inverse-warp + reconstruction-loss hot path (BASELINE.json metric:
warped px/s of the fused warp+loss fwd+bwd, and % of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]            # the reference's CPU PyTorch path (port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY 8d "C2"): 4-scale stereo photometric loss fwd+bwd, batch 64
per GPU at 128x416 (levels 128x416, 64x208, 32x104, 16x52), one source view (the rectified-stereo pose),
fp32 NCHW, synthetic KITTI-shaped inputs, image pyramids prebuilt (they are inputs of the path, 8d).
A step = ONE launch of dvf_photo_loss_fused_pose (pose_vec2mat + projection in the kernel prologue, warp + loss +
all gradients over all levels, pose backward in the epilogue).  Steps rotate over several distinct input sets whose total
size exceeds L2; one CUDA graph holds one round of them (--single-step-graphs: one graph per step).
Multi-GPU: batch sharded, B=64 per rank (weak scaling), no data-path collective; one NCCL all-reduce of the
loss terms closes the timed region (logging exchange).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "depth-vo-feat_b200")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "warped_px_per_s_fused_warp_loss_fwd_bwd"
UNIT = "warped px/s"
H, W, LEVELS = 128, 416, 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="batch per GPU")
    ap.add_argument("--sets", type=int, default=4, help="distinct input sets rotated through (working set > L2)")
    ap.add_argument("--prewarm-ms", type=float, default=400.0, help="untimed clock ramp before the W warm-up steps")
    ap.add_argument("--roofline-ms", type=float, default=1500.0, help="length of the dominant-kernel timing loop")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--cpu-batch", type=int, default=8, help="batch of the bounded CPU sample")
    ap.add_argument("--single-step-graphs", action="store_true", help="one CUDA graph per step instead of one per round of --sets steps")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--iid-depth", action="store_true", help="stress case: iid-noise depth instead of the smooth field")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
def make_inputs(B, seed, smooth=True):
    """CPU tensors of one C2 batch: full-resolution target/source images, depth per level, stereo pose, K."""
    from dvf_b200 import synthetic as syn
    tgt, src = syn.images(B, 3, H, W, seed + 1, smooth=True, n=2)
    depths = [syn.depth(B, H >> s, W >> s, seed + 10 + s, smooth=smooth) for s in range(LEVELS)]
    pose = syn.pose(B, "stereo", seed).unsqueeze(1)          # [B,1,6]
    K, Kinv = syn.intrinsics(B, H, W)
    return dict(tgt=tgt, src=src, depths=depths, pose=pose, K=K, Kinv=Kinv)


def warped_px(B, V=1):
    return B * V * sum((H >> s) * (W >> s) for s in range(LEVELS))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.marks, self.proc = [], [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self, t0, t1):
        self.marks.append((t0, t1))

    def finish(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            loaded = any(a - 0.03 <= ts <= b + 0.03 for a, b in self.marks)
            try:
                if loaded:
                    sm.append(float(f[0]))
                    power.append(float(f[2]))
                    for n, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(n)
                smax = float(f[1])
            except ValueError:
                continue
        sm.sort()
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples_under_load": len(sm), "power_w_max": (max(power) if power else None)}


# ------------------------------------------------------------------------------------------------
# reference / cpu baseline (oracle/torch_port.py restates the reference's torch op sequence)
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(B):
    from oracle import torch_port as ref
    d = make_inputs(B, seed=4242)
    tgt, src, K, Kinv = d["tgt"], d["src"], d["K"], d["Kinv"]

    def step():
        depths = [x.unsqueeze(1).clone().requires_grad_(True) for x in d["depths"]]
        pose = d["pose"].clone().requires_grad_(True)
        loss = ref.loss_multi_scale(tgt, [src], K, Kinv, depths, [None] * LEVELS, pose)
        loss.backward()
        return float(loss.detach())

    return step


def cpu_baseline(seconds, B):
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    step = cpu_reference_step_fn(B)
    step(); step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds or n >= 200:
            break
    return {"value": warped_px(B) * n / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} fwd+bwd iterations of the same 4-scale stereo loss at batch {B} ({warped_px(B)} warped px each), "
                      f"torch {torch.__version__} CPU, {cores} threads, oracle/torch_port.py"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    B = args.cpu_batch
    step = cpu_reference_step_fn(B)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    val = warped_px(B) * args.steps / el
    sample = (f"each step = fwd+bwd of the 4-scale stereo loss on a bounded batch of {B} (not {args.batch}); "
              f"torch {torch.__version__} CPU, {cores} threads; reference op sequence restated in oracle/torch_port.py "
              f"(the reference checkout is not present on the GPU box)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "4-scale stereo photometric loss fwd+bwd, 128x416, CPU bounded sample", "batch": B,
                   "levels": LEVELS, "views": 1},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's CUDA path
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from dvf_b200 import ops
    from dvf_b200.plan import FusedLossPlan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    sizes = [(H >> s, W >> s) for s in range(LEVELS)]
    ds = [float(1 << s) for s in range(LEVELS)]

    # ---- inputs: one CPU draw per rank, rotated into `sets` distinct device copies ----------------
    host = make_inputs(B, seed=1000 + rank, smooth=not args.iid_depth)
    plans, graphs, loss_graphs = [], [], []
    for k in range(args.sets):
        roll = lambda t: torch.roll(t, shifts=k, dims=0).contiguous().to(dev)   # noqa: E731
        tgt_pyr = ops.area_pyramid(roll(host["tgt"]), sizes)
        src_pyr = ops.area_pyramid(roll(host["src"]), sizes)
        plan = FusedLossPlan(tgt_pyr, [[s] for s in src_pyr], [roll(x) for x in host["depths"]], roll(host["pose"]),
                             roll(host["K"]), roll(host["Kinv"]), downscales=ds)
        plans.append(plan)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    round_graph = None
    with torch.cuda.stream(side):
        for p in plans:
            graphs.append(p.capture())
            loss_graphs.append(p.capture(loss_only=True))
        if args.sets > 1 and not args.single_step_graphs:
            # one graph = one ROUND of `sets` steps (one launch per input set), as a training iteration captured whole
            # would hold them: consecutive steps are kernel -> kernel edges inside the graph instead of separate graph
            # launches (~2 us of launch gap per step less).  Still one launch per step, K launches for K steps.
            round_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(round_graph):
                for p in plans:
                    p.launch()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()

    def run_steps(n, singles):
        """n steps over the rotating input sets: whole rounds from the round graph, the rest one graph per step"""
        i = 0
        if round_graph is not None:
            for _ in range(n // args.sets):
                round_graph.replay()
            i = (n // args.sets) * args.sets
        for k in range(i, n):
            singles[k % args.sets].replay()
    wpx_step = plans[0].warped_px
    bytes_launch = plans[0].algorithmic_bytes()
    set_bytes = bytes_launch
    assert wpx_step == warped_px(B)

    sampler = ClockSampler(local) if rank == 0 else None

    def spin(graph_list, ms):
        t_end = time.perf_counter() + ms / 1e3
        i = 0
        while time.perf_counter() < t_end:
            for _ in range(32):
                graph_list[i % len(graph_list)].replay()
                i += 1
            torch.cuda.synchronize()
        return i

    # ---- value: K steps, device-timed, max over ranks ----------------------------------------------
    t_load0 = time.time()
    spin(graphs, args.prewarm_ms)
    for i in range(args.warmup):
        graphs[i % args.sets].replay()
    acc_terms = torch.zeros_like(plans[0].terms)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(args.steps, graphs)
    if world > 1:
        acc_terms.copy_(plans[(args.steps - 1) % args.sets].terms)
        dist.all_reduce(acc_terms)     # the only exchange of the path: <= 16 floats of loss terms
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = wpx_step * world * args.steps / (ms_total * 1e-3)

    # ---- roofline: the dominant kernel alone, CUDA events on its launch stream --------------------
    n_roof = max(50, int(args.roofline_ms * 1e-3 / max(ms_total * 1e-3 / args.steps, 1e-6)))
    n_roof = min(n_roof, 200000)
    for i in range(10):
        loss_graphs[i % args.sets].replay()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    r0.record()
    run_steps(n_roof, loss_graphs)
    r1.record()
    torch.cuda.synchronize()
    t_load1 = time.time()
    if sampler:
        sampler.mark(t_load0, t_load1)
    kernel_ms = r0.elapsed_time(r1) / n_roof
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bytes_launch / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(REPO, "profiles", "traffic.json"))).get("photo_loss_kernel_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "dvf::photo_loss_c3x2_kernel<1,zeros,noexpl,grad,tma> (dvf_photo_loss_fused_pose)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": bytes_launch, "bytes_per_warped_px": bytes_launch / wpx_step,
                "kernel_us": kernel_ms * 1e3, "launches_timed": n_roof, "frac_of_nominal_8000": achieved / 8000.0}

    # ---- e2e: public drop-in API, inputs from pinned host memory every step ------------------------
    e2e = None
    if not args.no_e2e:
        import loss_functions_sfm as sfm
        pin = lambda t: t.contiguous().pin_memory()   # noqa: E731
        h_t, h_s = pin(host["tgt"]), pin(host["src"])
        h_d = [pin(x.unsqueeze(1)) for x in host["depths"]]
        h_p, h_K, h_Ki = pin(host["pose"]), pin(host["K"]), pin(host["Kinv"])
        h2d = sum(t.numel() * 4 for t in [h_t, h_s, h_p, h_K, h_Ki] + h_d)

        # Inputs of step i+1 are uploaded on a copy stream while step i computes (the usual prefetching loader loop);
        # every step's inputs cross PCIe once, inside the timed region, and every step ends with a D2H read of its loss.
        copy_stream = torch.cuda.Stream()

        def upload():
            with torch.cuda.stream(copy_stream):
                up = lambda t: t.to(dev, non_blocking=True)   # noqa: E731
                bufs = dict(tgt=up(h_t), src=up(h_s), depths=[up(x) for x in h_d], pose=up(h_p), K=up(h_K), Ki=up(h_Ki))
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return bufs, ev

        def e2e_step(cur):
            bufs, ev = cur
            nxt = upload()                                   # step i+1's H2D overlaps this step's kernels
            cs = torch.cuda.current_stream()
            cs.wait_event(ev)
            for t in [bufs["tgt"], bufs["src"], bufs["pose"], bufs["K"], bufs["Ki"]] + bufs["depths"]:
                t.record_stream(cs)
            depths = [x.requires_grad_(True) for x in bufs["depths"]]
            pose = bufs["pose"].requires_grad_(True)
            loss = sfm.photometric_reconstruction_loss(bufs["tgt"], [bufs["src"]], bufs["K"], bufs["Ki"], depths,
                                                       [None] * LEVELS, pose)
            loss.backward()
            loss.item()             # device -> host read of the step's result
            return nxt

        cur = upload()
        for _ in range(3):
            cur = e2e_step(cur)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_a = time.time()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(args.e2e_steps):
            cur = e2e_step(cur)
        q1.record()
        torch.cuda.synchronize()
        if sampler:
            sampler.mark(t_a, time.time())
        ems = torch.tensor([q0.elapsed_time(q1)], device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": wpx_step * world * args.e2e_steps / (float(ems.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": args.e2e_steps,
               "api": "loss_functions_sfm.photometric_reconstruction_loss(...) + loss.backward() (includes the area pyramid); "
                      "pinned-host inputs of step i+1 uploaded on a copy stream while step i computes, loss.item() every step"}

    clocks = sampler.finish() if sampler else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.cpu_seconds, args.cpu_batch)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: 4-scale stereo photometric loss fwd+bwd, batch 64/GPU at 128x416 (BASELINE configs[1])",
                       "batch_per_gpu": B, "global_batch": B * world, "levels": LEVELS, "views": 1, "layout": "NCHW fp32",
                       "depth_field": "iid-noise (stress)" if args.iid_depth else "smooth (17x17 box-filtered disparity)",
                       "pyramid": "prebuilt inputs (SURVEY 8d)", "parallelism": f"batch-sharded dp{world}",
                       "l2_policy": f"{args.sets} rotating input sets, {args.sets * set_bytes / 1e6:.0f} MB > 126 MB L2",
                       "step": "ONE launch: dvf_photo_loss_fused_pose (pose_vec2mat+projection, warp+loss+all gradients over 4 levels, pose backward); "
                               + ("CUDA graph per step" if round_graph is None else f"CUDA graphs of {args.sets} consecutive steps (one per input set)")},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": plans[0].n_launches * args.steps,
            "warped_px_per_step_per_gpu": wpx_step,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
