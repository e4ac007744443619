#!/usr/bin/env python
"""Benchmark of the fused inverse-warp + reconstruction-loss hot path (BASELINE.json metric:
warped px/s of the fused warp+loss fwd+bwd, and % of the HBM roofline).

    python bench.py [--config C2] [--gpus N] [--steps K] [--warmup W]     # this repo's CUDA path
    python bench.py --impl reference [--config C2] [--steps K] [--warmup W]  # the reference's CPU PyTorch path (port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs):
    C1      configs[0]: loss_functions.photometric_reconstruction_loss (2 views, 1 scale) on 4 x 3 x 128 x 416
    C2      configs[1]: 4-scale stereo photometric loss fwd+bwd, batch 64 per GPU at 128 x 416, 1 view  [default]
    C3      configs[2]: stereo + temporal depth-odometry loss: 2 views (temporal pose at PoseExpNet scale + stereo),
            explainability masks, 4 scales, GLOBAL batch 256 sharded over the GPUs (strong scaling)
    C4      configs[3]: photometric loss (2 views, 128 x 416) + feature reconstruction loss on 64-channel maps at 1/4
            resolution (bf16 channels-last, gradients to all three maps), GLOBAL batch 128
    C5loss  configs[4], the warp+loss part: 256 x 832, 4 scales, 2 views + masks, batch 64 per GPU (512 on 8)
A step = the fused launch(es) of one loss evaluation with all gradients (pose_vec2mat + projection in the kernel
prologue, warp + loss + gradients over all levels and views, pose backward in the epilogue).  Image pyramids are
prebuilt inputs.  Steps rotate over several distinct input sets whose total size exceeds L2; CUDA graphs
hold --graph-steps consecutive steps.
Multi-GPU: batch sharded, no data-path collective.  The loss terms (<= 16 floats) are exchanged EVERY step: by default
fused into the loss kernel, whose epilogue stores the rank's terms into every peer's buffer over NVLink (an all-gather by
peer-to-peer stores, dvf_b200.dist.PeerTerms; --exchange nccl: an NCCL all-reduce on a side stream inside the captured
graph, which costs ~15 us per step because its kernel competes with the persistent loss kernel for SM slots).  The main value keeps the workload of --config at every N (C2: fixed batch per GPU =
weak scaling, so that the driver's per-N efficiency compares like with like); when --config is C2 the line also carries
`strong_c3`: the C3 workload at a FIXED global batch of 256 on the same N GPUs.
"""
from __future__ import annotations

import argparse
import hashlib
import glob
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

WORKLOADS = {
    "C1": dict(H=128, W=416, levels=1, views=2, expl=False, batch=4, batch_is="per_gpu", temporal="kitti",
               title="C1: loss_functions.photometric_reconstruction_loss fwd+bwd, 2 views, 4x3x128x416 (BASELINE configs[0])"),
    "C2": dict(H=128, W=416, levels=4, views=1, expl=False, batch=64, batch_is="per_gpu", temporal=None,
               title="C2: 4-scale stereo photometric loss fwd+bwd, batch 64/GPU at 128x416 (BASELINE configs[1])"),
    "C3": dict(H=128, W=416, levels=4, views=2, expl=True, batch=256, batch_is="global", temporal="tiny",
               title="C3: stereo+temporal depth-odometry loss fwd+bwd, 2 views + explainability masks, 4 scales, "
                     "global batch 256 at 128x416 (BASELINE configs[2])"),
    "C4": dict(H=128, W=416, levels=1, views=2, expl=False, batch=128, batch_is="global", temporal="kitti",
               feature=dict(C=64, h=32, w=104, dtype="bf16"),
               title="C4: photometric loss (2 views, 128x416) + feature reconstruction loss on 64-ch bf16 channels-last maps "
                     "at 32x104 with gradients to all maps, global batch 128 (BASELINE configs[3])"),
    "C5loss": dict(H=256, W=832, levels=4, views=2, expl=True, batch=64, batch_is="per_gpu", temporal="kitti",
                   title="C5loss: warp+loss part of the high-res step, 256x832, 4 scales, 2 views + masks, batch 64/GPU "
                         "(BASELINE configs[4]; the conv nets stay on cuDNN and are not part of the path)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the workload's batch (per GPU or global, as the workload defines it)")
    ap.add_argument("--sets", type=int, default=4, help="distinct input sets rotated through (working set > L2)")
    ap.add_argument("--graph-steps", type=int, default=64, help="consecutive steps per CUDA graph (at most)")
    ap.add_argument("--prewarm-ms", type=float, default=400.0, help="untimed clock ramp before the W warm-up steps")
    ap.add_argument("--roofline-ms", type=float, default=1500.0, help="length of the dominant-kernel timing loop")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--cpu-batch", type=int, default=0, help="batch of the CPU legs (0 = the GPU arm's per-GPU batch)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip strong_c3 / unfused_gpu")
    ap.add_argument("--ctas-per-sm", type=int, default=0, help="tuning: dvf_loss_desc.ctas_per_sm of the image-loss plans (0 = automatic)")
    ap.add_argument("--no-pdl", action="store_true", help="plain launches instead of programmatic dependent launches (A/B)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl", "none"],
                    help="N > 1, every step: p2p = the loss kernel stores its terms into every peer's buffer over NVLink (fused "
                         "all-gather, dvf_b200.dist.PeerTerms; falls back to nccl if no peer mapping is available); nccl = "
                         "all-reduce on a side stream inside the graph; none = no exchange (A/B)")
    ap.add_argument("--iid-depth", action="store_true", help="stress case: iid-noise depth instead of the smooth field")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
def local_batch(wl, world, override=0):
    """(images per rank, global batch); the sharding rule is dvf_b200.dist.shard_range (equal contiguous slices here)"""
    from dvf_b200.dist import shard_range
    b = override or wl["batch"]
    if wl["batch_is"] == "global":
        if b % world:
            raise SystemExit(f"global batch {b} does not divide over {world} GPUs")
        lo, hi = shard_range(b, 0, world)
        return hi - lo, b
    return b, b * world


def make_inputs(wl, B, seed, smooth=True):
    """CPU tensors of one batch: full-resolution target / source images, depth per level, poses [B,V,6], K
    (+ explainability masks per level, + feature maps and their depth for C4)."""
    from dvf_b200 import synthetic as syn
    H, W, L, V = wl["H"], wl["W"], wl["levels"], wl["views"]
    imgs = syn.images(B, 3, H, W, seed + 1, smooth=True, n=1 + V)
    depths = [syn.depth(B, H >> s, W >> s, seed + 10 + s, smooth=smooth) for s in range(L)]
    poses = [syn.pose(B, wl["temporal"], seed + 3)] if V == 2 else []
    poses.append(syn.pose(B, "stereo", seed + 4))
    K, Kinv = syn.intrinsics(B, H, W)
    d = dict(tgt=imgs[0], srcs=imgs[1:], depths=depths, pose=torch.stack(poses, 1).contiguous(), K=K, Kinv=Kinv)
    if wl["expl"]:
        d["expl"] = [syn.explainability(B, V, H >> s, W >> s, seed + 20 + s) for s in range(L)]
    f = wl.get("feature")
    if f:
        d["feat"] = syn.features(B, f["C"], f["h"], f["w"], seed + 30, n=1 + V)
        d["feat_depth"] = syn.depth(B, f["h"], f["w"], seed + 40, smooth=smooth)
        d["feat_K"], d["feat_Kinv"] = syn.intrinsics(B, f["h"], f["w"])
    return d


def warped_px(wl, B):
    n = B * wl["views"] * sum((wl["H"] >> s) * (wl["W"] >> s) for s in range(wl["levels"]))
    f = wl.get("feature")
    if f:
        n += B * wl["views"] * f["h"] * f["w"]
    return n


def config_dict(name, wl, world, args):
    """The `config` object of the JSON line -- identical for the b200 and the reference arm."""
    Bl, Bg = local_batch(wl, world, args.batch)
    return {"workload": wl["title"], "name": name, "batch_per_gpu": Bl, "global_batch": Bg, "H": wl["H"], "W": wl["W"],
            "levels": wl["levels"], "views": wl["views"], "explainability_masks": wl["expl"],
            "feature_maps": wl.get("feature"), "layout": "NCHW fp32 images" + (", NHWC bf16 feature maps" if wl.get("feature") else ""),
            "depth_field": "iid-noise (stress)" if args.iid_depth else "smooth (17x17 box-filtered disparity)",
            "pyramid": "prebuilt inputs (SURVEY 8d)", "parallelism": f"batch-sharded dp{world}",
            "scaling": "strong" if wl["batch_is"] == "global" else "weak"}


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
# the reference's torch operator sequence (oracle/torch_port.py restates it; used on the CPU as the reference arm /
# cpu_baseline and on the GPU as the "unfused stock PyTorch" comparison)
# ------------------------------------------------------------------------------------------------
def torch_port_step_fn(wl, B, device="cpu"):
    from oracle import torch_port as ref
    d = make_inputs(wl, B, seed=4242)
    mv = lambda t: t.to(device)   # noqa: E731
    tgt, srcs, K, Kinv = mv(d["tgt"]), [mv(s) for s in d["srcs"]], mv(d["K"]), mv(d["Kinv"])
    depths0 = [mv(x) for x in d["depths"]]
    pose0 = mv(d["pose"])
    expl0 = [mv(x) for x in d["expl"]] if wl["expl"] else None
    f = wl.get("feature")
    if f:
        feat0 = [mv(x) for x in d["feat"]]
        fdepth0, fK, fKi = mv(d["feat_depth"]), mv(d["feat_K"]), mv(d["feat_Kinv"])
    L, V = wl["levels"], wl["views"]

    def step():
        pose = pose0.clone().requires_grad_(True)
        if L == 1 and V == 2:   # loss_functions.py:7-20
            depth = depths0[0].clone().requires_grad_(True)
            loss = ref.loss_two_view(tgt, srcs[0], srcs[1], depth, pose[:, 0], pose[:, 1], K, Kinv)
            if f:               # unsupervise.py:104-111
                feats = [x.clone().requires_grad_(True) for x in feat0]
                fdepth = fdepth0.clone().requires_grad_(True)
                loss = loss + ref.loss_two_view(feats[0], feats[1], feats[2], fdepth, pose[:, 0], pose[:, 1], fK, fKi)
        else:                   # loss_functions_sfm.py:9-46
            depths = [x.unsqueeze(1).clone().requires_grad_(True) for x in depths0]
            masks = [x.clone().requires_grad_(True) for x in expl0] if expl0 else [None] * L
            loss = ref.loss_multi_scale(tgt, srcs, K, Kinv, depths, masks, pose)
        loss.backward()
        return loss.detach()

    return step


def cpu_baseline(wl, seconds, B):
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    step = torch_port_step_fn(wl, B)
    step(); step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds or n >= 200:
            break
    out = {"value": warped_px(wl, B) * n / el, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{n} fwd+bwd iterations of the same workload at batch {B} ({warped_px(wl, B)} warped px each), "
                     f"torch {torch.__version__} CPU, {cores} threads, oracle/torch_port.py"}
    # SURVEY 8(d) also asks for a single-thread figure: a few iterations at a batch that keeps this leg to seconds
    B1 = max(1, min(B, 8))
    torch.set_num_threads(1)
    try:
        step1 = torch_port_step_fn(wl, B1)
        step1()
        n1, t0 = 0, time.perf_counter()
        while True:
            step1()
            n1 += 1
            el1 = time.perf_counter() - t0
            if el1 >= min(seconds, 4.0) or n1 >= 20:
                break
        out["one_thread"] = {"value": warped_px(wl, B1) * n1 / el1, "unit": UNIT, "cores": 1,
                             "sample": f"{n1} iterations at batch {B1}"}
    finally:
        torch.set_num_threads(cores)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wl = WORKLOADS[args.config]
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    Bl, _ = local_batch(wl, world, args.batch)
    B = args.cpu_batch or Bl
    step = torch_port_step_fn(wl, B)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        float(step())
    el = time.perf_counter() - t0
    val = warped_px(wl, B) * args.steps / el
    sample = (f"each step = fwd+bwd of the workload on one rank's batch of {B} on the host cores; "
              f"torch {torch.__version__} CPU, {cores} threads; reference op sequence restated in oracle/torch_port.py "
              f"(the reference checkout is not present on the GPU box)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if wl["batch_is"] == "global" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.config, wl, world, args),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's CUDA path
# ------------------------------------------------------------------------------------------------
class Step:
    """The fused launch(es) of one loss evaluation on one resident input set."""

    def __init__(self, plans):
        self.plans = plans
        self.warped_px = sum(p.warped_px for p in plans)
        self.bytes = sum(p.algorithmic_bytes() for p in plans)
        self.n_launches = sum(p.n_launches for p in plans)
        self.dominant = max(plans, key=lambda p: p.algorithmic_bytes())
        self.n_terms = sum(p.terms.numel() for p in plans)
        self.chained = all(bool(p.desc.flags & 64) for p in plans)   # DVF_FLAG_PDL_CHAINED

    def copy_terms_into(self, buf):
        o = 0
        for p in self.plans:
            buf[o:o + p.terms.numel()].copy_(p.terms)
            o += p.terms.numel()

    def launch(self, edge=False):
        """edge: first or last step of a chain of dependent launches -- nobody to overlap with on one side, so the launch takes
        the full grid (the others are sized for overlap, DVF_FLAG_PDL_CHAINED)"""
        for p in self.plans:
            if self.chained:
                p.set_chained(not edge)
            p.launch()


def build_steps(wl, B, Bg, host, dev, sets, pdl, peers=None, rank=0, ctas_per_sm=0):
    """peers: None, or a list that receives one dvf_b200.dist.PeerTerms per plan of a step (fused exchange of the terms)"""
    from dvf_b200 import ops
    from dvf_b200.plan import FusedLossPlan
    from dvf_b200.dist import PeerTerms

    def peer_kw(j, n_terms, k):
        if peers is None:
            return {}
        if len(peers) <= j:
            peers.append(PeerTerms(sets, n_terms, dev))
        return dict(peer_terms=peers[j].slot_ptrs(k), peer_rank=rank)
    H, W, L, V = wl["H"], wl["W"], wl["levels"], wl["views"]
    sizes = [(H >> s, W >> s) for s in range(L)]
    ds = [float(1 << s) for s in range(L)]
    steps = []
    for k in range(sets):
        roll = lambda t: torch.roll(t, shifts=k, dims=0).contiguous().to(dev)   # noqa: E731
        tgt_pyr = ops.area_pyramid(roll(host["tgt"]), sizes)
        src_pyrs = [ops.area_pyramid(roll(s), sizes) for s in host["srcs"]]
        pose, K, Kinv = roll(host["pose"]), roll(host["K"]), roll(host["Kinv"])
        expl = [roll(x) for x in host["expl"]] if wl["expl"] else None
        plans = [FusedLossPlan(tgt_pyr, [[sp[l] for sp in src_pyrs] for l in range(L)], [roll(x) for x in host["depths"]],
                               pose, K, Kinv, expl_levels=expl, downscales=ds, global_batch=Bg, pdl=pdl,
                               pdl_chained=pdl and not wl.get("feature"), ctas_per_sm=ctas_per_sm, **peer_kw(0, L * V, k))]
        f = wl.get("feature")
        if f:
            cl = lambda t: roll(t).to(torch.bfloat16 if f["dtype"] == "bf16" else torch.float32).contiguous(   # noqa: E731
                memory_format=torch.channels_last)
            feats = [cl(x) for x in host["feat"]]
            plans.append(FusedLossPlan([feats[0]], [feats[1:]], [roll(host["feat_depth"])], pose, roll(host["feat_K"]),
                                       roll(host["feat_Kinv"]), map_grads=True, global_batch=Bg, pdl=pdl, **peer_kw(1, V, k)))
        steps.append(Step(plans))
    return steps


class Runner:
    """CUDA graphs over the rotating input sets: one graph holds up to `graph_steps` consecutive steps (a run of n steps is
    n // graph_steps replays of the long graph plus ONE graph with the remainder, captured ahead of the timed region by
    prepare()).  N > 1 with --exchange nccl: the loss terms of every step are all-reduced on a side stream inside the graph
    (step i's exchange overlaps step i+1's kernel; the graph joins the side stream at its end)."""

    def __init__(self, steps, graph_steps, world, allreduce, only=None):
        import torch.distributed as dist
        self.steps, self.world = steps, world
        self.n_sets = len(steps)
        self.graph_steps = max(self.n_sets, (graph_steps // self.n_sets) * self.n_sets)
        self.exchange = "none"
        def launch(s, edge):
            if only is None:
                s.launch(edge)
                return
            p = only(s)
            if s.chained:
                p.set_chained(not edge)
            p.launch()
        self.main = torch.cuda.Stream()
        self.main.wait_stream(torch.cuda.current_stream())
        xs = torch.cuda.Stream()
        self.do_ar = do_ar = world > 1 and allreduce
        # term buffers of the exchange: one per step of the long graph (an all-reduce may still be in flight when the
        # same input set is launched again)
        self.xbuf = ([torch.zeros(steps[0].n_terms, device=steps[0].plans[0].terms.device) for _ in range(self.graph_steps)]
                     if do_ar else None)
        self.xs = xs
        copied = self._copied = {}

        def body(i, edge=False):
            st = steps[i % self.n_sets]
            cur = torch.cuda.current_stream()
            if do_ar and (i - self.n_sets) in copied:
                cur.wait_event(copied.pop(i - self.n_sets))   # this input set's terms have left for the exchange
            launch(st, edge)
            if do_ar:
                # copy + all-reduce on the side stream: the loss kernels stay back to back on the main stream
                xs.wait_stream(cur)
                with torch.cuda.stream(xs):
                    st.copy_terms_into(self.xbuf[i % self.graph_steps])
                    ev = torch.cuda.Event()
                    ev.record(xs)
                    copied[i] = ev
                    dist.all_reduce(self.xbuf[i % self.graph_steps])
        self._body = body
        self.graphs = {}
        self.eager = False
        with torch.cuda.stream(self.main):
            for i in range(self.n_sets):         # untimed: first launches, NCCL communicator warm-up
                body(i)
            self.main.wait_stream(xs)
            torch.cuda.synchronize()
            copied.clear()
        try:
            self._graph(self.graph_steps)
            self.exchange = ("NCCL all-reduce of the loss terms every step, on a side stream inside the CUDA graph"
                             if do_ar else "none")
        except Exception as e:   # NCCL refused capture: eager per-step exchange, still every step
            if not do_ar:
                raise
            torch.cuda.synchronize()
            copied.clear()
            self.eager = True
            self.exchange = f"NCCL all-reduce every step, eager on a side stream (graph capture failed: {type(e).__name__})"
        torch.cuda.current_stream().wait_stream(self.main)
        torch.cuda.synchronize()

    def _graph(self, n_steps):
        g = self.graphs.get(n_steps)
        if g is None:
            torch.cuda.synchronize()
            with torch.cuda.stream(self.main):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.main, capture_error_mode="thread_local"):
                    for i in range(n_steps):
                        self._body(i, i == 0 or i == n_steps - 1)
                    if self.do_ar:
                        torch.cuda.current_stream().wait_stream(self.xs)
                self._copied.clear()
            torch.cuda.current_stream().wait_stream(self.main)
            torch.cuda.synchronize()
            self.graphs[n_steps] = g
        return g

    def prepare(self, n):
        """capture whatever run(n) will replay and launch it once (the first launch of a graph uploads it), outside any timed
        region"""
        if self.eager:
            return
        for k in ({self.graph_steps} if n >= self.graph_steps else set()) | ({n % self.graph_steps} - {0}):
            fresh = k not in self.graphs
            g = self._graph(k)
            if fresh:
                g.replay()
                torch.cuda.synchronize()

    def run(self, n):
        if self.eager:
            for i in range(n):
                self._body(i)
            return
        q, r = divmod(n, self.graph_steps)
        long = self._graph(self.graph_steps)
        for _ in range(q):
            long.replay()
        if r:
            self._graph(r).replay()

    def spin(self, ms):
        t_end = time.perf_counter() + ms / 1e3
        while time.perf_counter() < t_end:
            self.run(2 * self.graph_steps)
            torch.cuda.synchronize()


def timed(runner, n, world, dev):
    import torch.distributed as dist
    runner.prepare(n)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    runner.run(n)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


LOSS_KERNEL_SOURCES = ("dvf_loss*", "dvf_math*", "dvf_tma.cuh", "dvf_reduce.cuh", "dvf_pose.cuh", "dvf_internal.h")


def csrc_sha16():
    """hash of the sources the fused loss kernels are built from: profiles/traffic.json records it at capture time"""
    h = hashlib.sha256()
    files = sorted({f for pat in LOSS_KERNEL_SOURCES for f in glob.glob(os.path.join(PKG, "csrc", pat))})
    for f in files:
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def measure_workload(name, wl, args, world, rank, dev, sampler, steps_n, warmup_n, roofline=True):
    """value (K steps, device-timed, max over ranks) and, on request, the dominant kernel's roofline."""
    Bl, Bg = local_batch(wl, world, args.batch if name == args.config else 0)
    host = make_inputs(wl, Bl, seed=1000 + rank, smooth=not args.iid_depth)
    pdl = args.sets >= 2 and not args.no_pdl
    mode, peers = (args.exchange if world > 1 else "none"), None
    if mode == "p2p":
        try:
            peers = []
            steps = build_steps(wl, Bl, Bg, host, dev, args.sets, pdl, peers=peers, rank=rank, ctas_per_sm=args.ctas_per_sm)
        except RuntimeError as e:
            if "peer-to-peer" not in str(e):
                raise
            mode, peers = "nccl", None
    if peers is None:
        steps = build_steps(wl, Bl, Bg, host, dev, args.sets, pdl, ctas_per_sm=args.ctas_per_sm)
    runner = Runner(steps, args.graph_steps, world, mode == "nccl")
    if mode == "p2p":
        runner.exchange = (f"fused into the loss kernel: every step its epilogue stores the rank's loss terms into every peer's buffer "
                           f"over NVLink (all-gather by peer-to-peer stores, mapping via {peers[0].how}); no collective launch")
    wpx = steps[0].warped_px
    assert wpx == warped_px(wl, Bl)
    t0 = time.time()
    runner.prepare(steps_n)          # capture the graphs of the timed region now: the GPU must not idle between warm-up and timing
    runner.prepare(max(warmup_n, 3))
    runner.spin(args.prewarm_ms)
    runner.run(max(warmup_n, 3))
    ms_total = timed(runner, steps_n, world, dev)
    xcheck = None
    if world > 1 and mode != "none":
        # what every rank holds for the LAST step against an NCCL all-reduce of the same terms (outside the timed region)
        import torch.distributed as dist
        last = (steps_n - 1) % args.sets
        mine = torch.cat([p.terms for p in steps[last].plans]).clone()
        ref = mine.clone()
        dist.all_reduce(ref)
        if mode == "p2p":
            got = torch.cat([pt.gathered(last).sum(0) for pt in peers])
        else:
            got = ref   # (buffers rotate inside the graphs; the nccl mode is the A/B arm, not the checked one)
        xcheck = float(((got - ref).abs().max() / ref.abs().max()).item())
    out = {"value": wpx * world * steps_n / (ms_total * 1e-3), "ms_per_step": ms_total / steps_n, "steps": steps_n,
           "warped_px_per_step_per_gpu": wpx, "gpu_launches": steps[0].n_launches * steps_n,
           "exchange": runner.exchange, "exchange_check_rel_err": xcheck, "set_bytes": steps[0].bytes, "batch_per_gpu": Bl, "global_batch": Bg}
    roof = None
    if roofline:
        dom = Runner(steps, args.graph_steps, 1, False, only=lambda s: s.dominant)
        n_roof = min(200000, max(50, int(args.roofline_ms * 1e-3 / max(ms_total * 1e-3 / steps_n, 1e-6))))
        dom.prepare(n_roof)
        dom.run(16)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        r0.record()
        dom.run(n_roof)
        r1.record()
        torch.cuda.synchronize()
        kernel_ms = r0.elapsed_time(r1) / n_roof
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        p = steps[0].dominant
        bytes_launch = p.algorithmic_bytes()
        achieved = bytes_launch / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_note = None, "no ncu capture recorded for this workload"
        try:
            tj = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))
            ent = tj.get("workloads", {}).get(name)
            if ent:
                if ent.get("csrc_sha16") == csrc_sha16():
                    traffic, traffic_note = ent["dram_bytes_per_launch"], ent.get("source")
                else:
                    traffic_note = "stale: profiles/traffic.json was captured with other kernel sources"
        except Exception:
            pass
        kname = ("dvf::photo_loss_nhwc_kernel" if p.layout == 1 else "dvf::photo_loss_c3x2_kernel") + \
            f"<V={p.V}, zeros, {'expl' if p.inputs[6] is not None else 'noexpl'}, grad> (dvf_photo_loss_fused_pose)"
        roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_note,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": bytes_launch, "bytes_per_warped_px": bytes_launch / p.warped_px,
                "kernel_us": kernel_ms * 1e3, "launches_timed": n_roof, "frac_of_nominal_8000": achieved / 8000.0}
    if sampler:
        sampler.mark(t0, time.time())
    return out, roof, host, steps


def e2e_leg(name, wl, args, world, dev, host, sampler):
    """Public drop-in API with HOST inputs: every step uploads its inputs from pinned memory (the upload of step i+1
    overlaps step i, the usual prefetching loader loop) and ends with a device -> host read of the loss."""
    import torch.distributed as dist
    import loss_functions as lf
    import loss_functions_sfm as sfm
    L, V = wl["levels"], wl["views"]
    pin = lambda t: t.contiguous().pin_memory()   # noqa: E731
    two_view = (L == 1 and V == 2)
    hb = dict(tgt=pin(host["tgt"]), srcs=[pin(s) for s in host["srcs"]], pose=pin(host["pose"]), K=pin(host["K"]),
              Ki=pin(host["Kinv"]), depths=[pin(x if two_view else x.unsqueeze(1)) for x in host["depths"]])
    if wl["expl"]:
        hb["expl"] = [pin(x) for x in host["expl"]]
    if wl.get("feature"):
        hb["feat"] = [pin(x) for x in host["feat"]]
        hb["fdepth"], hb["fK"], hb["fKi"] = pin(host["feat_depth"]), pin(host["feat_K"]), pin(host["feat_Kinv"])

    def flat(d):
        for v in d.values():
            if isinstance(v, list):
                yield from v
            else:
                yield v
    h2d = sum(t.numel() * t.element_size() for t in flat(hb))
    copy_stream = torch.cuda.Stream()

    def upload():
        with torch.cuda.stream(copy_stream):
            up = lambda t: t.to(dev, non_blocking=True)   # noqa: E731
            bufs = {k: ([up(x) for x in v] if isinstance(v, list) else up(v)) for k, v in hb.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return bufs, ev

    def step(cur):
        bufs, ev = cur
        nxt = upload()
        cs = torch.cuda.current_stream()
        cs.wait_event(ev)
        for t in flat(bufs):
            t.record_stream(cs)
        pose = bufs["pose"].requires_grad_(True)
        depths = [x.requires_grad_(True) for x in bufs["depths"]]
        if two_view:
            loss = lf.photometric_reconstruction_loss(bufs["tgt"], bufs["srcs"][0], bufs["srcs"][1], depths[0], pose[:, 0],
                                                      pose[:, 1], bufs["K"], bufs["Ki"])
            if "feat" in bufs:
                feats = [x.requires_grad_(True) for x in bufs["feat"]]
                fd = bufs["fdepth"].requires_grad_(True)
                loss = loss + lf.photometric_reconstruction_loss(feats[0], feats[1], feats[2], fd, pose[:, 0], pose[:, 1],
                                                                 bufs["fK"], bufs["fKi"])
        else:
            masks = [x.requires_grad_(True) for x in bufs["expl"]] if "expl" in bufs else [None] * L
            loss = sfm.photometric_reconstruction_loss(bufs["tgt"], bufs["srcs"], bufs["K"], bufs["Ki"], depths, masks, pose)
        loss.backward()
        loss.item()             # device -> host read of the step's result
        return nxt

    cur = upload()
    for _ in range(3):
        cur = step(cur)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_a = time.time()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record()
    for _ in range(args.e2e_steps):
        cur = step(cur)
    q1.record()
    torch.cuda.synchronize()
    if sampler:
        sampler.mark(t_a, time.time())
    ems = torch.tensor([q0.elapsed_time(q1)], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    Bl = host["tgt"].shape[0]
    return {"value": warped_px(wl, Bl) * world * args.e2e_steps / (float(ems.item()) * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": args.e2e_steps,
            "ms_per_step": float(ems.item()) / args.e2e_steps,
            "api": ("loss_functions.photometric_reconstruction_loss" if two_view else
                    "loss_functions_sfm.photometric_reconstruction_loss") +
                   "(...) + loss.backward() through the drop-in modules (area pyramid included); fp32 inputs in pinned host memory, "
                   "step i+1's upload on a copy stream while step i computes, loss.item() every step"}


def unfused_gpu_leg(wl, B, dev, seconds=3.0):
    """The reference's own torch operator sequence (oracle/torch_port.py) on the SAME GPU, inputs resident: what
    fusing buys over stock PyTorch-CUDA eager."""
    step = torch_port_step_fn(wl, B, device=dev)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    while True:
        for _ in range(5):
            step()
        n += 5
        torch.cuda.synchronize()
        if time.perf_counter() - t0 > seconds or n >= 500:
            break
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return {"value": warped_px(wl, B) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": n, "batch": B,
            "what": f"oracle/torch_port.py (the reference's torch op sequence, unfused) with torch {torch.__version__} CUDA eager on the same "
                    "GPU, inputs resident, fwd+bwd (includes F.interpolate pyramids, as the reference's loss does)"}


def run_b200(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        t = torch.zeros(16, device=dev)
        dist.all_reduce(t)          # communicator warm-up, untimed
        torch.cuda.synchronize()
    name, wl = args.config, WORKLOADS[args.config]
    sampler = ClockSampler(local) if rank == 0 else None

    main, roof, host, steps = measure_workload(name, wl, args, world, rank, dev, sampler, args.steps, args.warmup)
    e2e = None if args.no_e2e else e2e_leg(name, wl, args, world, dev, host, sampler)
    Bl = main["batch_per_gpu"]
    del steps
    extras = {}
    if not args.no_extras and name == "C2":
        # strong scaling on C3 at a fixed global batch of 256 (BASELINE configs[2]), same run, same GPUs
        c3 = WORKLOADS["C3"]
        s_out, s_roof, _, s_steps = measure_workload("C3", c3, args, world, rank, dev, sampler, args.steps, args.warmup)
        del s_steps
        extras["strong_c3"] = {"metric": METRIC, "value": s_out["value"], "unit": UNIT, "scaling": "strong", "n_gpus": world,
                               "ms_per_step": s_out["ms_per_step"], "steps": s_out["steps"], "global_batch": s_out["global_batch"],
                               "batch_per_gpu": s_out["batch_per_gpu"], "exchange": s_out["exchange"],
                               "exchange_check_rel_err": s_out["exchange_check_rel_err"],
                               "workload": c3["title"], "roofline": s_roof}
    if not args.no_extras and rank == 0:
        try:
            extras["unfused_gpu"] = unfused_gpu_leg(wl, Bl, dev)
        except Exception as e:   # e.g. out of memory on the large shapes: the fused path does not depend on it
            extras["unfused_gpu"] = {"unavailable": f"{type(e).__name__}: {str(e)[:120]}"}
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    clocks = sampler.finish() if sampler else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(wl, args.cpu_seconds, args.cpu_batch or min(Bl, 64))

    if rank == 0:
        cfg = config_dict(name, wl, world, args)
        out = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "detail": {"l2_policy": f"{args.sets} rotating input sets, {args.sets * main['set_bytes'] / 1e6:.0f} MB > 126 MB L2",
                       "step": f"{main['gpu_launches'] // args.steps} launch(es) of dvf_photo_loss_fused_pose per step (pose_vec2mat + "
                               f"projection, warp + loss + all gradients over all levels and views, pose backward); CUDA graphs of "
                               f"{args.graph_steps} consecutive steps" + ("" if (args.no_pdl or args.sets < 2) else
                               "; consecutive steps work on disjoint input sets and are chained by programmatic dependent launch "
                               "(DVF_FLAG_PDL | DVF_FLAG_PDL_CHAINED): step i+1's CTAs start on the SM slots step i leaves or frees, "
                               "grids sized for overlap (~64 units per CTA)"),
                       "exchange": main["exchange"], "exchange_check_rel_err": main["exchange_check_rel_err"]},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": main["gpu_launches"], "warped_px_per_step_per_gpu": main["warped_px_per_step_per_gpu"],
        }
        out.update(extras)
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
