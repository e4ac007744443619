"""Static-shape execution plan for the fused reconstruction loss: allocate once, launch many times.

A training loop with fixed shapes binds its tensors once and then issues ONE launch per step through the
C ABI -- dvf_photo_loss_fused_pose: pose_vec2mat and projection in the kernel prologue, warp + loss + all
gradients for every level and view, pose backward in the epilogue -- optionally replayed from a CUDA graph.
(The three-launch form dvf_pose_proj_fwd / dvf_photo_loss_fused / dvf_pose_proj_bwd is kept for callers
that supply their own projection matrices.)  No autograd bookkeeping, no allocation, no host
synchronisation.  Gradients are those of `sum of loss terms` (upstream gradient 1), exactly what loss.backward() yields for the reference's
photometric_reconstruction_loss (loss_functions_sfm.py:9-46 / loss_functions.py:7-20).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import PADDING, ROTATION, dvf_level, dvf_loss_desc, dvf_pose_args


class FusedLossPlan:
    def __init__(self, tgt_levels: Sequence[torch.Tensor], src_levels: Sequence[Sequence[torch.Tensor]],
                 depth_levels: Sequence[torch.Tensor], pose: torch.Tensor, K: torch.Tensor, Kinv: torch.Tensor,
                 expl_levels: Optional[Sequence[torch.Tensor]] = None, downscales: Optional[Sequence[float]] = None,
                 rotation_mode: str = "euler", padding_mode: str = "zeros", need_grad: bool = True,
                 map_grads: bool = False, global_batch: Optional[int] = None, align_corners: bool = False,
                 upstream: Optional[torch.Tensor] = None, fused_pose: bool = True, use_tma: bool = True,
                 piece_overhead: int = 0, ctas_per_sm: int = 0, pdl: bool = False, pdl_chained: bool = False,
                 peer_terms: Optional[Sequence[int]] = None, peer_rank: int = 0, bf16_grads: bool = True,
                 disparity_eps: Optional[float] = None, img_scale: float = 1.0):
        """Maps: dense NCHW fp32 (images, any C) or channels-last fp32 / bf16 feature maps ([B,C,H,W] tensors in
        torch.channels_last memory).  map_grads: also produce d/d tgt and d/d src (fp32, layout of the maps).
        global_batch: this plan holds B of the global_batch images of a sharded batch (dvf_loss_desc.mean_batch).
        upstream: device scalar multiplying every gradient (dvf_loss_desc.upstream).
        pdl: programmatic dependent launch (DVF_FLAG_PDL): back-to-back launches of plans with DISJOINT buffers overlap.
        pdl_chained: this plan is launched in a chain of such launches (DVF_FLAG_PDL_CHAINED): grid sized for overlap.
        peer_terms: device pointers (as mapped in this process) of every rank's [n_peers][L*V] exchange buffer; the kernel
        epilogue stores this rank's loss terms into row peer_rank of each (dvf_b200.dist.PeerTerms)."""
        from .ops import _nhwc_ok
        self.lib = _lib.load()
        dev = pose.device
        L, V = len(depth_levels), pose.shape[1]
        B, Cc = tgt_levels[0].shape[0], tgt_levels[0].shape[1]
        self.B, self.C, self.V, self.L = B, Cc, V, L
        self.rotation = ROTATION[rotation_mode]
        self.inputs = (list(tgt_levels), [list(s) for s in src_levels], list(depth_levels), pose, K, Kinv,
                       None if expl_levels is None else list(expl_levels), upstream)   # keep alive
        maps = list(tgt_levels) + [s for lv in src_levels for s in lv]
        if all(_nhwc_ok(t) for t in maps) and len({t.dtype for t in maps}) == 1:
            layout, dtype = _lib.NHWC, (_lib.BF16 if maps[0].dtype == torch.bfloat16 else _lib.F32)
        else:
            layout, dtype = _lib.NCHW, _lib.F32
            for t in maps:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        for t in list(depth_levels) + [pose, K, Kinv] + ([] if expl_levels is None else list(expl_levels)):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        self.layout, self.dtype = layout, dtype
        self.elem_bytes = 2 if dtype == _lib.BF16 else 4
        ds = [1.0] * L if downscales is None else [float(x) for x in downscales]
        self.ds = (C.c_float * L)(*ds)
        self.P = torch.empty(L, B * V, 3, 4, device=dev)
        self.Kinv_s = torch.empty(L, B, 3, 3, device=dev)
        self.terms = torch.empty(L * V, device=dev)
        self.gP = torch.empty(L, B * V, 3, 4, device=dev) if need_grad else None
        self.gpose = torch.empty(B, V, 6, device=dev) if need_grad else None
        self.gdepth = [torch.empty_like(d) for d in depth_levels] if need_grad else None
        self.gexpl = ([torch.empty(B, V, e.shape[2], e.shape[3], device=dev) for e in expl_levels]
                      if (need_grad and expl_levels is not None) else None)
        self.map_grads = bool(map_grads and need_grad)
        self.grad_bf16 = bool(bf16_grads) and layout == _lib.NHWC and dtype == _lib.BF16
        gdt = torch.bfloat16 if self.grad_bf16 else torch.float32
        self.gtgt = [torch.empty_like(t, dtype=gdt) for t in tgt_levels] if self.map_grads else None
        self.gsrc = ([[torch.empty_like(t, dtype=torch.float32) for t in lv] for lv in src_levels]
                     if self.map_grads else None)   # zero-filled by the entry on every launch (DVF_FLAG_ZERO_GSRC)
        self.levels = (dvf_level * L)()
        for l in range(L):
            lv = self.levels[l]
            lv.H, lv.W = depth_levels[l].shape[1], depth_levels[l].shape[2]
            lv.depth, lv.tgt = depth_levels[l].data_ptr(), tgt_levels[l].data_ptr()
            for v in range(V):
                lv.src[v] = src_levels[l][v].data_ptr()
            lv.P, lv.Kinv = self.P[l].data_ptr(), self.Kinv_s[l].data_ptr()
            if expl_levels is not None:
                e = expl_levels[l]
                lv.expl, lv.expl_bstride = e.data_ptr(), e.shape[1] * e.shape[2] * e.shape[3]
            if need_grad:
                lv.gdepth, lv.gP = self.gdepth[l].data_ptr(), self.gP[l].data_ptr()
                if self.gexpl is not None:
                    lv.gexpl = self.gexpl[l].data_ptr()
                if self.map_grads:
                    lv.gtgt = self.gtgt[l].data_ptr()
                    for v in range(V):
                        lv.gsrc[v] = self.gsrc[l][v].data_ptr()
        flags = (_lib.FLAG_ALIGN_CORNERS if align_corners else 0) | (_lib.FLAG_ZERO_GSRC if self.map_grads else 0) | \
                (0 if use_tma else _lib.FLAG_NO_TMA) | (_lib.FLAG_PDL if pdl else 0) | \
                (_lib.FLAG_PDL_CHAINED if (pdl and pdl_chained) else 0)
        self.desc = dvf_loss_desc(B, Cc, V, L, dtype, layout, PADDING[padding_mode], flags, int(global_batch or 0), _lib.BF16 if self.grad_bf16 else _lib.F32,
                                  int(piece_overhead), int(ctas_per_sm), None if upstream is None else upstream.data_ptr(), None,
                                  0, 0, None, float(disparity_eps or 0.0), float(img_scale))
        if disparity_eps is not None:
            self.desc.flags |= _lib.FLAG_DISPARITY   # depth_levels hold disparities; gdepth = d/d disparity
        if peer_terms:
            self._peer_arr = (C.c_void_p * len(peer_terms))(*[int(x) for x in peer_terms])   # keep alive
            self.desc.n_peers, self.desc.peer_rank, self.desc.peer_terms = len(peer_terms), int(peer_rank), self._peer_arr
        n = self.lib.dvf_photo_loss_workspace_bytes(C.byref(self.desc), self.levels)
        if n == 0:
            raise _lib.DvfError("dvf_photo_loss_workspace_bytes rejected the shapes")
        self.ws = torch.zeros(n, dtype=torch.uint8, device=dev)   # private workspace: plans may be in flight together
        self.warped_px = sum(B * V * d.shape[1] * d.shape[2] for d in depth_levels)
        self.need_grad = need_grad
        # single-launch form: pose -> P in the kernel prologue, d pose in its epilogue
        self.pose_args = dvf_pose_args(pose.data_ptr(), K.data_ptr(), Kinv.data_ptr(), self.ds, self.rotation, 0,
                                       self.gpose.data_ptr() if need_grad else None)
        self.fused_pose = bool(fused_pose)   # False: three-launch form (debug / A-B timing)
        self.n_launches = 1 if self.fused_pose else (3 if need_grad else 2)

    # -- the three launches ---------------------------------------------------------------------
    def launch_pose_fwd(self, stream: int):
        pose, K, Kinv = self.inputs[3], self.inputs[4], self.inputs[5]
        _lib.check(self.lib.dvf_pose_proj_fwd(pose.data_ptr(), K.data_ptr(), Kinv.data_ptr(), self.B, self.V, self.rotation,
                                              self.ds, self.L, None, self.P.data_ptr(), self.Kinv_s.data_ptr(), stream),
                   "dvf_pose_proj_fwd")

    def launch_loss(self, stream: int):
        _lib.check(self.lib.dvf_photo_loss_fused(C.byref(self.desc), self.levels, self.terms.data_ptr(), self.ws.data_ptr(),
                                                 self.ws.numel(), stream), "dvf_photo_loss_fused")

    def launch_pose_bwd(self, stream: int):
        pose, K = self.inputs[3], self.inputs[4]
        _lib.check(self.lib.dvf_pose_proj_bwd(self.gP.data_ptr(), None, pose.data_ptr(), K.data_ptr(), self.B, self.V,
                                              self.rotation, self.ds, self.L, self.gpose.data_ptr(), stream),
                   "dvf_pose_proj_bwd")

    def set_chained(self, chained: bool):
        """switch DVF_FLAG_PDL_CHAINED for the following launches (a chain's first and last launch have no neighbour on one
        side: with the full grid they ramp the GPU up and down faster)"""
        if self.desc.flags & _lib.FLAG_PDL:
            if chained:
                self.desc.flags |= _lib.FLAG_PDL_CHAINED
            else:
                self.desc.flags &= ~_lib.FLAG_PDL_CHAINED

    def launch_fused(self, stream: int):
        """ONE launch: pose_vec2mat + projection, warp + loss + all gradients, pose backward."""
        _lib.check(self.lib.dvf_photo_loss_fused_pose(C.byref(self.desc), self.levels, C.byref(self.pose_args),
                                                      self.terms.data_ptr(), self.ws.data_ptr(), self.ws.numel(), stream),
                   "dvf_photo_loss_fused_pose")

    def launch(self, stream: Optional[int] = None, fused_pose: Optional[bool] = None):
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        if self.fused_pose if fused_pose is None else fused_pose:
            self.launch_fused(st)
            return
        self.launch_pose_fwd(st)
        self.launch_loss(st)
        if self.need_grad:
            self.launch_pose_bwd(st)

    def capture(self, loss_only: bool = False) -> torch.cuda.CUDAGraph:
        """Capture one step (or only the fused loss kernel) into a CUDA graph."""
        self.launch()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if loss_only and not self.fused_pose:
                self.launch_loss(torch.cuda.current_stream().cuda_stream)
            else:
                self.launch()   # fused-pose form: the step IS the one loss kernel
        return g

    def loss(self) -> torch.Tensor:
        return self.terms.sum()

    # -- algorithmic (compulsory) HBM bytes of one fused-loss launch, SURVEY 8(d) ------------------
    def algorithmic_bytes(self) -> int:
        e = self.elem_bytes
        g = 4 if self.need_grad else 0
        per_tpx = 4 + self.C * e + g                                      # depth + target (+ d depth), once per target pixel
        per_wpx = self.C * e                                              # source texels, once per warped pixel
        if self.map_grads:
            per_tpx += self.C * (2 if self.grad_bf16 else 4)              # d target, written once
            per_wpx += self.C * 4                                         # d source (fp32, accumulated), one write per texel
        if self.inputs[6] is not None:
            per_wpx += 4 + g                                              # explainability read (+ its gradient)
        tpx = self.warped_px // self.V
        return tpx * per_tpx + self.warped_px * per_wpx
