"""Batch sharding of the reconstruction loss across the GPUs of one node (one process per GPU).

Every image is independent (per-image K, pose, depth, sources); the only cross-sample coupling of
loss_functions.py:13 / loss_functions_sfm.py:33 is the `mean`, whose denominator B*C*H*W is known in
advance.  So each rank runs the same fused kernel on its contiguous slice of the batch and

    loss_global = sum_r (B_r / B) * loss_r ,     d loss_global / d x_r = (B_r / B) * d loss_r / d x_r

No data-path collective is needed; the loss terms are all-reduced (<= 16 floats) for logging only.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous [begin, end) slice of the batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(global_batch, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def local_weight(global_batch: int, rank: int, world: int) -> float:
    """Factor that turns a rank-local mean loss (and its gradients) into its share of the global mean."""
    b, e = shard_range(global_batch, rank, world)
    return (e - b) / float(global_batch)


def all_reduce_terms(local_terms: torch.Tensor, weight: float, group=None, async_op: bool = False):
    """Global loss terms = sum over ranks of weight_r * local_terms_r (one tiny all-reduce, NCCL or gloo)."""
    buf = local_terms.detach() * weight
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return (buf, work) if async_op else buf
    return (buf, None) if async_op else buf


class PeerTerms:
    """Exchange buffers for the loss terms, written by the loss kernel itself (dvf_loss_desc.peer_terms).

    Every rank owns `slots` buffers of [world][n_terms] floats in memory that its peers of the same node can address
    over NVLink.  A loss launch with slot s stores the launching rank's terms into row `rank` of slot s on EVERY rank
    -- an all-gather fused into the kernel's epilogue: no collective launch, no extra kernel competing with the
    persistent loss kernel for SM slots (an NCCL all-reduce per step costs ~15 us next to a 52 us kernel on 2 GPUs,
    profiles/r2_summary.md).  `gathered(s).sum(0)` is the global term vector; rows written by peers are valid once their
    launches have completed (stream / event order plus a barrier, or simply one step late for logging).

    Collective constructor (all ranks of `group` call it).  Peer mappings come from torch's symmetric memory, or, where
    that is unavailable, from CUDA IPC handles exchanged through the process group; raises if neither works (callers
    then fall back to an NCCL all-reduce).
    """

    def __init__(self, slots: int, n_terms: int, device: torch.device, group=None):
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.slots, self.n_terms = int(slots), int(n_terms)
        self.slot_elems = self.world * self.n_terms
        n = self.slots * self.slot_elems
        self.how = None
        self._keep = []
        errs = []
        for how in ("symmetric_memory", "cuda_ipc"):
            try:
                getattr(self, "_init_" + how)(n, device, group)
                self.how = how
                break
            except Exception as e:   # noqa: BLE001 -- any failure means "try the next mechanism"
                errs.append(f"{how}: {type(e).__name__}: {e}")
        ok = torch.tensor([1 if self.how else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            raise RuntimeError("no peer-to-peer mapping available: " + "; ".join(errs))
        self.local.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)

    def _init_symmetric_memory(self, n, device, group):
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(n, dtype=torch.float32, device=device)
        h = symm.rendezvous(t, group=dist.group.WORLD if group is None else group)
        self.local = t
        self.ptrs = [int(p) for p in h.buffer_ptrs]
        self._keep.append(h)

    def _init_cuda_ipc(self, n, device, group):
        t = torch.zeros(n, dtype=torch.float32, device=device)
        handle = t.untyped_storage()._share_cuda_()
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        self.local = t
        self.ptrs = []
        for r, hnd in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(t.data_ptr())
                continue
            st = torch.UntypedStorage._new_shared_cuda(*hnd)
            peer = torch.empty(0, dtype=torch.float32, device=st.device).set_(st)[:n]
            # make the peer's memory addressable from kernels on this device (torch enables peer access on first copy)
            torch.empty(1, device=device).copy_(peer[:1])
            self._keep.append(peer)
            self.ptrs.append(peer.data_ptr())

    def slot_ptrs(self, slot: int):
        """device pointers of slot `slot` in every rank's buffer (entry q = rank q's), for FusedLossPlan(peer_terms=...)"""
        off = (slot % self.slots) * self.slot_elems * 4
        return [p + off for p in self.ptrs]

    def gathered(self, slot: int) -> torch.Tensor:
        """[world, n_terms] view of this rank's copy of slot `slot` (row r = rank r's terms)"""
        s = slot % self.slots
        return self.local[s * self.slot_elems:(s + 1) * self.slot_elems].view(self.world, self.n_terms)
