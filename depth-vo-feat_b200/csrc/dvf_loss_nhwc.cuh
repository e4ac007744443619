// dvf_loss_nhwc.cuh -- fused reconstruction loss for FEATURE MAPS in channels-last layout ([B,H,W,C], fp32 or bf16).
//
// Same operator as dvf_loss_kernel.cuh (loss_functions.py:7-20 applied to FeatExtractor features,
// unsupervise.py:104-109: gradients flow to all three maps).  With C = 32..64 channels the per-pixel work is
// dominated by the channel loop, so the mapping changes: a pixel is owned by a GROUP of lanes (C / vec lanes,
// vec = 16 bytes of channels); every bilinear tap is ONE coalesced 16-byte load per lane (the whole C-vector of
// a texel is contiguous), the validity mask / d(ix,iy) sums are sub-warp shuffle reductions, the target-map
// gradient is a plain vector store and the source-map gradient scatter is ONE vector reduction per lane and
// tap (red.global.add.v4.f32, sm_90+) instead of 4*C scalar atomics per pixel.  The coordinate chain is the
// scalar exact one (dvf_math.cuh), evaluated redundantly by the lanes of a group.
// Arithmetic is fp32 throughout; bf16 inputs are widened on load (geometry stays fp32), gradients to the maps
// are produced in fp32 NHWC buffers.
#pragma once
#include <cuda_bf16.h>

#include "dvf_loss_kernel.cuh"

namespace dvf {

template <bool kBf16>
struct VecIO;
template <>
struct VecIO<false> {   // 4 fp32 channels per lane
  static constexpr int kVec = 4;
  static __device__ __forceinline__ void load(const void* base, size_t elem, bool pred, float (&v)[4]) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pred) t = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <>
struct VecIO<true> {    // 8 bf16 channels per lane
  static constexpr int kVec = 8;
  static __device__ __forceinline__ void load(const void* base, size_t elem, bool pred, float (&v)[8]) {
    uint4 t = make_uint4(0u, 0u, 0u, 0u);
    if (pred) t = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + elem));
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {   // bf16 -> fp32 is a 16-bit shift
      v[2 * q] = __uint_as_float(w[q] << 16);
      v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
    }
  }
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// C = lanes_per_px * kVec;  lanes_per_px in {1,2,4,8,16,32}
template <int kV, bool kZeros, bool kBf16>
__global__ void __launch_bounds__(kLossThreads, 4) photo_loss_nhwc_kernel(const __grid_constant__ LossParams prm) {
  constexpr int kVec = VecIO<kBf16>::kVec;
  __shared__ __align__(16) float s_P[kV][12];
  __shared__ __align__(16) float s_M[12];

  int l = 0;
  while (l + 1 < prm.n_levels && (int)blockIdx.x >= prm.lv[l + 1].block_begin) ++l;
  const LevelDev& lv = prm.lv[l];
  const int rel = (int)blockIdx.x - lv.block_begin;
  const int b = rel / lv.blocks_per_image;
  const int chunk = rel - b * lv.blocks_per_image;
  const int tid = threadIdx.x, lane = tid & 31;
  const int H = lv.H, W = lv.W, HW = lv.HW, C = prm.C;
  const int lpp = C / kVec;                 // lanes per pixel
  const int slot = lane & (lpp - 1);        // my slot in the group
  const int ch0 = slot * kVec;               // my first channel
  const Geo geo = lv.geo;
  const bool need_grad = prm.need_grad != 0;
  const bool allow_fast = lv.allow_fast != 0;
  const bool has_expl = lv.expl != nullptr;
  const float inv_n = lv.inv_n;

  load_matrices<kV>(prm, lv, b, s_P, s_M);
  __syncthreads();

  float acc[kV][kRedSlots];
#pragma unroll
  for (int v = 0; v < kV; ++v)
#pragma unroll
    for (int k = 0; k < kRedSlots; ++k) acc[v][k] = 0.0f;
  float M[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) M[k] = s_M[k];

  const size_t img_px = (size_t)b * HW;     // pixel offset of this image; element offset = px * C
  const float* depth_b = lv.depth + img_px;
  float* gtgt_b = lv.gtgt ? lv.gtgt + img_px * C : nullptr;
  const int px_per_iter = kLossThreads / lpp;
  const int px_begin = chunk * lv.px_per_cta;
  const int px_end = min(px_begin + lv.px_per_cta, HW);

  for (int base = px_begin; base < px_end; base += px_per_iter) {
    const int idx = base + tid / lpp;
    const bool live = idx < px_end;
    const int idc = live ? idx : px_end - 1;
    Cam cam;
    {
      const int i = (int)fastdiv((uint32_t)idc, lv.divW);
      pixel_to_cam(M, ld_stream(depth_b + idc), i, idc - i * W, cam);
    }
    float tg[kVec], gt[kVec];
    VecIO<kBf16>::load(lv.tgt, (img_px + idc) * C + ch0, live, tg);
#pragma unroll
    for (int c = 0; c < kVec; ++c) gt[c] = 0.0f;
    float gd = 0.0f;

#pragma unroll
    for (int v = 0; v < kV; ++v) {
      float P[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) P[k] = s_P[v][k];
      Proj pr;
      Loc L;
      const bool fast = project<false, kZeros>(P, cam, geo, pr) && allow_fast;
      if (__builtin_expect(!fast, 0)) pr = project_exact<kZeros>(&s_P[v][0], cam, &lv.geo);
      locate<kZeros>(pr.xn, pr.yn, H, W, geo, L);
      const bool bnw = L.bnw && live, bne = L.bne && live, bsw = L.bsw && live, bse = L.bse && live;
      const size_t o_nw = (img_px + (size_t)(L.y0 * W + L.x0)) * C + ch0;     // only dereferenced under the tap predicates
      float a0[kVec], a1[kVec], a2[kVec], a3[kVec];
      VecIO<kBf16>::load(lv.src[v], o_nw, bnw, a0);
      VecIO<kBf16>::load(lv.src[v], o_nw + C, bne, a1);
      VecIO<kBf16>::load(lv.src[v], o_nw + (size_t)W * C, bsw, a2);
      VecIO<kBf16>::load(lv.src[v], o_nw + (size_t)W * C + C, bse, a3);
      const float ex = (has_expl && live) ? ld_stream(lv.expl + (size_t)b * lv.expl_bstride + (size_t)v * HW + idc) : 1.0f;
      const float wnw = mul(L.s, L.e), wne = mul(L.s, L.w), wsw = mul(L.n, L.e), wse = mul(L.n, L.w);

      float d0[kVec], d1[kVec];
      bool any = false;
#pragma unroll
      for (int c = 0; c < kVec; ++c) {
        const float wv = bilerp(a0[c], a1[c], a2[c], a3[c], wnw, wne, wsw, wse);
        any |= (wv != 0.0f);
        d0[c] = sub(tg[c], wv);
        d1[c] = has_expl ? mul(d0[c], ex) : d0[c];
      }
      // value-based mask over ALL channels of the pixel: OR across the lanes of the group
      for (int o = 1; o < lpp; o <<= 1) any |= (__shfl_xor_sync(0xffffffffu, (int)any, o) != 0);
      float lsum = 0.0f;
#pragma unroll
      for (int c = 0; c < kVec; ++c) lsum += fabsf(d1[c]);
      acc[v][12] += any ? lsum : 0.0f;        // every lane adds its own channels

      if (need_grad) {
        float gx = 0.0f, gy = 0.0f, ge = 0.0f, g[kVec];
#pragma unroll
        for (int c = 0; c < kVec; ++c) {
          const float gd1 = signed_unit(d1[c], inv_n, any);
          g[c] = has_expl ? mul(gd1, ex) : gd1;
          ge += gd1 * d0[c];
          gt[c] += g[c];
          bilerp_grad(a0[c], a1[c], a2[c], a3[c], L, -g[c], gx, gy);
        }
        for (int o = 1; o < lpp; o <<= 1) {   // sums over all channels of the pixel
          gx += __shfl_xor_sync(0xffffffffu, gx, o);
          gy += __shfl_xor_sync(0xffffffffu, gy, o);
          ge += __shfl_xor_sync(0xffffffffu, ge, o);
        }
        if (lv.gsrc[v] && any) {               // scatter: one 16-byte reduction per tap per kVec/4 quad
          float* gs = lv.gsrc[v];
#pragma unroll
          for (int q = 0; q < kVec; q += 4) {
            if (bnw) red_add_v4(gs + o_nw + q, -g[q] * wnw, -g[q + 1] * wnw, -g[q + 2] * wnw, -g[q + 3] * wnw);
            if (bne) red_add_v4(gs + o_nw + C + q, -g[q] * wne, -g[q + 1] * wne, -g[q + 2] * wne, -g[q + 3] * wne);
            if (bsw) red_add_v4(gs + o_nw + (size_t)W * C + q, -g[q] * wsw, -g[q + 1] * wsw, -g[q + 2] * wsw, -g[q + 3] * wsw);
            if (bse) red_add_v4(gs + o_nw + (size_t)W * C + C + q, -g[q] * wse, -g[q + 1] * wse, -g[q + 2] * wse, -g[q + 3] * wse);
          }
        }
        ChainGrad cg;
        chain_backward<false>(P, cam, pr, L, gx, gy, geo, cg);
        if (__builtin_expect(!fast, 0)) cg = chain_backward_exact(&s_P[v][0], cam, pr, L, gx, gy, &lv.geo);
        if (slot == 0 && live) {               // one lane per pixel owns the per-pixel outputs
          if (lv.gexpl) st_stream(lv.gexpl + ((size_t)b * kV + v) * HW + idx, ge);
          gd = add(gd, cg.gdepth);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[v][r * 4 + k] = fmaf(cg.gq[r], cam.cam[k], acc[v][r * 4 + k]);
            acc[v][r * 4 + 3] += cg.gq[r];
          }
        }
      }
    }  // views
    if (need_grad && live) {
      if (slot == 0 && lv.gdepth) st_stream(lv.gdepth + img_px + idx, gd);
      if (gtgt_b) {
        float* q = gtgt_b + (size_t)idx * C + ch0;
#pragma unroll
        for (int c = 0; c < kVec; c += 4) *reinterpret_cast<float4*>(q + c) = make_float4(gt[c], gt[c + 1], gt[c + 2], gt[c + 3]);
      }
    }
  }
  reduce_and_finish<kV, kLossThreads>(acc, prm, lv, l, rel, b, C);
}

template <int kV, bool kZeros>
void launch_loss_nhwc(const LossParams& prm, int blocks, bool bf16, cudaStream_t st);

}  // namespace dvf
