"""CPU, world_size 2 over gloo: the batch-sharding rule of dvf_b200.dist reproduces the un-sharded loss and
gradients (checked with the CPU oracle -- no GPU in this test)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "depth-vo-feat_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from dvf_b200 import dist as ddist, synthetic as syn
    from oracle import cpu_oracle as O
    H, W = 16, 52
    d = syn.stereo_temporal_batch(B, H, W, seed=3)
    b0, b1 = ddist.shard_range(B, rank, world)
    sl = slice(b0, b1)
    P = np.stack([O.project(d["intrinsics"].numpy(), O.pose_vec2mat(d[k].numpy())) for k in ("T_2to1", "T_R2L")], 1)
    r = O.photo_loss_P(d["img_R2"].numpy()[sl], [d["img_R1"].numpy()[sl], d["img_L2"].numpy()[sl]], d["depth"].numpy()[sl],
                       np.ascontiguousarray(P[sl]), d["intrinsics_inv"].numpy()[sl])
    w = ddist.local_weight(B, rank, world)
    terms = ddist.all_reduce_terms(torch.from_numpy(r["terms"]), w)
    gdepth = torch.zeros(B, H, W)
    gdepth[sl] = torch.from_numpy(r["gdepth"]) * w
    dist.all_reduce(gdepth)   # only to compare against the un-sharded result in rank 0
    if rank == 0:
        full = O.photo_loss_P(d["img_R2"].numpy(), [d["img_R1"].numpy(), d["img_L2"].numpy()], d["depth"].numpy(), P,
                              d["intrinsics_inv"].numpy())
        out.put((terms.numpy(), full["terms"], gdepth.numpy(), full["gdepth"]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_loss_equals_full_batch():
    world, B = 2, 5   # uneven split: 3 + 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    terms, full_terms, gdepth, full_gdepth = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    np.testing.assert_allclose(terms, full_terms, rtol=1e-12)
    assert np.abs(gdepth - full_gdepth).max() <= 1e-6 * np.abs(full_gdepth).max()


def test_shard_ranges_cover_batch():
    from dvf_b200 import dist as ddist
    for B in (1, 5, 64, 257):
        for world in (1, 2, 3, 8):
            spans = [ddist.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert abs(sum(ddist.local_weight(B, r, world) for r in range(world)) - 1.0) < 1e-12
