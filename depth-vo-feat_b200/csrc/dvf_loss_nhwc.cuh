// dvf_loss_nhwc.cuh -- fused reconstruction loss for FEATURE MAPS in channels-last layout ([B,H,W,C], fp32 or bf16).
//
// Same operator as dvf_loss_kernel.cuh (loss_functions.py:7-20 applied to FeatExtractor features,
// unsupervise.py:104-109: gradients flow to all three maps).  With C = 32..64 channels the per-pixel work is
// dominated by the channel loop, so the mapping changes: a pixel is owned by a GROUP of lanes (C / vec lanes,
// vec = 16 bytes of channels); every bilinear tap is ONE coalesced 16-byte load per lane (the whole C-vector of
// a texel is contiguous), the validity mask / d(ix,iy) sums are sub-warp shuffle reductions, the target-map
// gradient is a plain vector store and the source-map gradient scatter is ONE vector reduction per lane and
// tap (red.global.add.v4.f32, sm_90+) instead of 4*C scalar atomics per pixel.  The coordinate chain is the
// scalar exact one (dvf_math.cuh), evaluated once per pixel by an owner lane and broadcast to the group.
// Arithmetic is fp32 throughout; bf16 inputs are widened on load (geometry stays fp32), gradients to the maps
// are produced in fp32 NHWC buffers, or in bf16 ones for bf16 maps when the caller asks for it (grad_dtype).
#pragma once
#include <cuda_bf16.h>

#include "dvf_loss_kernel.cuh"

namespace dvf {

// one 16-byte vector of a texel's channels: 4 fp32 or 8 bf16 values per lane, loaded raw and widened at use
template <bool kBf16>
struct VecIO;
template <>
struct VecIO<false> {
  static constexpr int kVec = 4;
  static __device__ __forceinline__ void widen(const uint4& t, float (&v)[4]) {
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  }
};
template <>
struct VecIO<true> {
  static constexpr int kVec = 8;
  static __device__ __forceinline__ void widen(const uint4& t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {   // bf16 -> fp32 is a 16-bit shift
      v[2 * q] = __uint_as_float(w[q] << 16);
      v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
    }
  }
};
// The channels of a lane.  fp32 maps: 4 consecutive channels (one 16-byte vector).  bf16 maps: TWO groups of 4
// consecutive channels, C/2 apart (two 8-byte vectors), so that the fp32 gradient of each group is again one 16-byte
// vector and the lanes of a pixel write contiguous 16-byte pieces -- with 8 consecutive bf16 channels per lane every
// vector reduction filled only half of each 32-byte sector and the bf16 kernel was slower than the fp32 one.
// base + byte_off: 32-bit offset inside one image (ONE wide multiply-add per address); zeros if !pred.
// Predicated load INTO the zeroed registers (written in PTX: the C form makes nvcc load into temporaries and copy
// them over under the predicate, four extra instructions per tap).
template <bool kBf16>
__device__ __forceinline__ uint4 ld_lane_vec(const char* base, int byte_off, int half_b, unsigned pred) {
  uint4 t = make_uint4(0u, 0u, 0u, 0u);
  if (kBf16) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 a, b;\n"
        "setp.ne.u32 p, %7, 0;\n"
        "mad.wide.s32 a, %5, 1, %4;\n"
        "mad.wide.s32 b, %6, 1, a;\n"
        "@p ld.global.nc.v2.u32 {%0, %1}, [a];\n"
        "@p ld.global.nc.v2.u32 {%2, %3}, [b];\n"
        "}\n"
        : "+r"(t.x), "+r"(t.y), "+r"(t.z), "+r"(t.w)
        : "l"(base), "r"(byte_off), "r"(half_b), "r"(pred));
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 a;\n"
        "setp.ne.u32 p, %6, 0;\n"
        "mad.wide.s32 a, %5, 1, %4;\n"
        "@p ld.global.nc.v4.u32 {%0, %1, %2, %3}, [a];\n"
        "}\n"
        : "+r"(t.x), "+r"(t.y), "+r"(t.z), "+r"(t.w)
        : "l"(base), "r"(byte_off), "r"(pred));
  }
  return t;
}

// 16-byte vector reduction at base + byte_off, executed under `pred` (a predicated instruction, not a branch)
__device__ __forceinline__ void red_add_v4(char* base, int byte_off, unsigned pred, float a, float b, float c, float d) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 a;\n"
      "setp.ne.u32 p, %6, 0;\n"
      "mad.wide.s32 a, %1, 1, %0;\n"
      "@p red.global.add.v4.f32 [a], {%2, %3, %4, %5};\n"
      "}\n" ::"l"(base), "r"(byte_off), "f"(a), "f"(b), "f"(c), "f"(d), "r"(pred)
      : "memory");
}

// bf16 target-map gradient (dvf_loss_desc.grad_dtype = DVF_BF16): four channels leave as two packed bf16x2 words.
// (The source-map gradients stay fp32: they are ACCUMULATED, and bf16x2 reductions -- red.global.add.noftz.v2.bf16x2,
// tried -- round after every one of the up to 4 contributions of a texel: 1.17e-2 of max|g| on the C4 shape, outside the
// 1e-2 class of bf16 results.)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// sums of (x, y[, z]) over aligned groups of kLanes lanes: straight-line xor butterfly (same order as a loop over
// q = 1, 2, 4, ...: results are bit-identical for every group size)
template <int kLanes>
__device__ __forceinline__ void group_sum(float& x, float& y, float& z, bool with_z) {
#pragma unroll
  for (int q = 1; q < kLanes; q <<= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, q);
    y += __shfl_xor_sync(0xffffffffu, y, q);
    if (with_z) z += __shfl_xor_sync(0xffffffffu, z, q);
  }
}
__device__ __forceinline__ void group_sum_dispatch(int lpp_shift, float& x, float& y, float& z, bool with_z) {
  switch (lpp_shift) {   // warp-uniform
    case 0: break;
    case 1: group_sum<2>(x, y, z, with_z); break;
    case 2: group_sum<4>(x, y, z, with_z); break;
    case 3: group_sum<8>(x, y, z, with_z); break;
    case 4: group_sum<16>(x, y, z, with_z); break;
    default: group_sum<32>(x, y, z, with_z); break;
  }
}

// C = lanes_per_px * kVec;  lanes_per_px in {1,2,4,8,16,32}
//
// A warp walks its share of the CTA's pixel run 32 pixels at a time, in three phases:
//   A  every lane OWNS one of the 32 pixels and evaluates the exact coordinate chain for it once per view
//      (pixel2cam, cam2pixel, bilinear cell) -- not once per lane of the channel group as before, which made
//      the kernel instruction-bound (9.5 k thread-instructions per warped pixel at C = 64);
//   B  the warp then visits the pixels 32/lpp at a time: the cell (offset, tap predicates, weights, mask weight)
//      is broadcast from the owner lane by shuffles, every lane handles its kVec channels of all four taps,
//      the value-based mask and the d(ix,iy) sums are folded over the group, the source-map gradient leaves as
//      one 16-byte reduction per lane and tap, and the folded sums are shuffled back to the owner;
//   C  the owner lanes run the backward of the coordinate chain for their pixel (depth gradient, dL/dP sums).
#ifndef DVF_NHWC_MINBLOCKS
#define DVF_NHWC_MINBLOCKS 4
#endif
template <int kV, bool kZeros, bool kBf16>
__global__ void __launch_bounds__(kLossThreads, DVF_NHWC_MINBLOCKS) photo_loss_nhwc_kernel(const __grid_constant__ LossParams prm) {
  constexpr int kVec = VecIO<kBf16>::kVec;
  constexpr unsigned kFull = 0xffffffffu;
  __shared__ __align__(16) float s_P[kV][12];
  __shared__ __align__(16) float s_M[12];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = prm.C;
  const int lpp = C / kVec;                 // lanes per pixel (power of two)
  const int lpp_shift = __ffs(lpp) - 1;
  const int pps = 32 >> lpp_shift;          // pixels the warp handles per step
  const int grp = lane >> lpp_shift;        // my channel group
  const int ch0 = (lane & (lpp - 1)) * 4;      // first channel of my (first) group of 4; bf16: second group at + C/2
  const unsigned grp_mask = lpp == 32 ? 0xffffffffu : ((1u << lpp) - 1u);
  const bool need_grad = prm.need_grad != 0;
  const bool grad_bf16 = prm.grad_bf16 != 0;   // target-map gradient in bf16 (bf16 maps only)

  // balanced split (same unit list and bookkeeping as the image kernel, dvf_loss_kernel.cuh): this CTA owns units
  // [w, w_end) and walks them (image, level) by (image, level)
  int w = split_start((int)blockIdx.x, prm);
  const int w_end = split_start((int)blockIdx.x + 1, prm);
  pdl_let_successor_start(prm);
  while (w < w_end) {
  const int b = (int)fastdiv((uint32_t)w, prm.div_upi);
  int l = 0;
  while (l + 1 < prm.n_levels && w - b * prm.units_per_image_all >= prm.lv[l + 1].unit_base) ++l;
  const LevelDev& lv = prm.lv[l];
  const int real0 = b * prm.units_per_image_all + lv.unit_base + prm.piece_overhead;   // first pixel-carrying unit
  const int real1 = real0 + lv.units_per_image;
  const int piece_end = min(w_end, real1);
  const int k0 = max(w, real0) - real0, k1 = piece_end - real0;   // units [k0, k1) of image b at level l
  const bool piece_starts_here = w <= real0, piece_ends_here = piece_end == real1;
  w = piece_end;
  if (k1 <= k0) continue;   // only overhead units fell into my range (CTA-uniform)
  const int first_cta = piece_starts_here ? (int)blockIdx.x : cta_of_unit(real0, prm);
  const int last_cta = piece_ends_here ? (int)blockIdx.x : cta_of_unit(real1 - 1, prm);
  const int n_parts = last_cta - first_cta + 1;
  const int part = (int)blockIdx.x - first_cta;
  const int H = lv.H, W = lv.W, HW = lv.HW;
  const Geo geo = lv.geo;
  const bool allow_fast = lv.allow_fast != 0;
  const bool has_expl = lv.expl != nullptr;
  const float inv_n = prm.upstream ? mul(lv.inv_n, __ldg(prm.upstream)) : lv.inv_n;

  load_matrices<kV>(prm, lv, b, s_P, s_M);
  __syncthreads();

  // Registers are the occupancy limit of this kernel, so the per-pixel state lives in shared memory: the cell of
  // every owned pixel (phase A -> B), the folded sums going back (B -> C) and the dL/dP sums (per warp).
  constexpr int kWarps = kLossThreads / 32;
  __shared__ __align__(16) float s_cell[kWarps][kV][32][8];   // {offset, predicates, w, e}, {n, s, mask weight, -}
  __shared__ __align__(16) float s_back[kWarps][kV][32][4];   // {sum gx, sum gy, sum ge, -}
  __shared__ float s_wacc[kWarps][kV][12];
  float acc_loss[kV];
#pragma unroll
  for (int v = 0; v < kV; ++v) {
    acc_loss[v] = 0.0f;
    if (lane < 12) s_wacc[warp][v][lane] = 0.0f;
  }
  __syncwarp();

  const size_t img_px = (size_t)b * HW;     // pixel offset of this image; element offset = px * C
  const float* depth_b = lv.depth + img_px;
  // per-image bases; inside an image every byte offset fits 32 bits (HW * C * 4 < 2^31, checked on the host)
  constexpr int kEsz = kBf16 ? 2 : 4;
  const char* const tgt_b = static_cast<const char*>(static_cast<const void*>(lv.tgt)) + img_px * C * kEsz;
  const char* src_bb[kV];
  float* gsrc_b[kV];
#pragma unroll
  for (int v = 0; v < kV; ++v) {
    src_bb[v] = static_cast<const char*>(static_cast<const void*>(lv.src[v])) + img_px * C * kEsz;
    gsrc_b[v] = lv.gsrc[v] ? lv.gsrc[v] + img_px * C : nullptr;
    asm volatile("" : "+l"(src_bb[v]), "+l"(gsrc_b[v]));   // keep the bases in registers (nvcc re-derives them per load otherwise)
  }
  float* gtgt_b = lv.gtgt ? lv.gtgt + (grad_bf16 ? img_px * C / 2 : img_px * C) : nullptr;
  const char* tgt_bb = tgt_b;
  asm volatile("" : "+l"(tgt_bb), "+l"(gtgt_b));
  const int row_b = W * C * kEsz, px_b = C * kEsz, ch_b = ch0 * kEsz, half_b = (C / 2) * kEsz;   // byte strides of the maps
  const int px_begin = k0 * kUnitPx;
  const int px_end = min(k1 * kUnitPx, HW);

  for (int base = px_begin + warp * 32; base < px_end; base += kLossThreads) {
    // ---- phase A: my own pixel -------------------------------------------------------------------
    const int idx = base + lane;
    const bool live = idx < px_end;
    const int idc = live ? idx : px_end - 1;
    Cam cam;
    {
      float M[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) M[k] = s_M[k];
      const int i = (int)fastdiv((uint32_t)idc, lv.divW);
      float dv = ld_stream(depth_b + idc);
      if (prm.disparity) dv = depth_of_disp(dv, prm.disp_eps);
      pixel_to_cam(M, dv, i, idc - i * W, cam);
    }
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
      // y0 * W + x0 of the cell (only dereferenced under the tap predicates nw | ne << 1 | sw << 2 | se << 3, which
      // are false for pixels past the run)
      const unsigned pk = live ? ((L.bnw ? 1u : 0u) | (L.bne ? 2u : 0u) | (L.bsw ? 4u : 0u) | (L.bse ? 8u : 0u)) : 0u;
      const float ex = (has_expl && live) ? ld_stream(lv.expl + (size_t)b * lv.expl_bstride + (size_t)v * HW + idc) : 1.0f;
      float4* cell = reinterpret_cast<float4*>(&s_cell[warp][v][lane][0]);
      cell[0] = make_float4(__int_as_float(L.y0 * W + L.x0), __uint_as_float(pk), L.w, L.e);
      cell[1] = make_float4(L.n, L.s, ex, __int_as_float(L.x0));
    }
    __syncwarp();

    // ---- phase B: channel work, pps pixels per step ------------------------------------------------
    // every load of a step goes out before any arithmetic: the target vector and the four taps of each view
    struct StepRegs {
      uint4 raw_t;
      uint4 raw[kV][4];
      int o_b[kV];
      unsigned pkv[kV];
      float4 cw[kV];   // {offset, predicates, w, e}
    };
    auto issue_loads = [&](int s, StepRegs& R) {
      const int p = s * pps + grp;
      const int pidx = base + p;
      const bool plive = pidx < px_end;
      R.raw_t = ld_lane_vec<kBf16>(tgt_bb, (plive ? pidx : px_end - 1) * px_b + ch_b, half_b, plive ? 1u : 0u);
#pragma unroll
      for (int v = 0; v < kV; ++v) {
        R.cw[v] = *reinterpret_cast<const float4*>(&s_cell[warp][v][p][0]);
        R.o_b[v] = __float_as_int(R.cw[v].x) * px_b + ch_b;
        R.pkv[v] = __float_as_uint(R.cw[v].y);
        R.raw[v][0] = ld_lane_vec<kBf16>(src_bb[v], R.o_b[v], half_b, R.pkv[v] & 1u);
        R.raw[v][1] = ld_lane_vec<kBf16>(src_bb[v], R.o_b[v] + px_b, half_b, R.pkv[v] & 2u);
        R.raw[v][2] = ld_lane_vec<kBf16>(src_bb[v], R.o_b[v] + row_b, half_b, R.pkv[v] & 4u);
        R.raw[v][3] = ld_lane_vec<kBf16>(src_bb[v], R.o_b[v] + row_b + px_b, half_b, R.pkv[v] & 8u);
      }
    };
    // (software-pipelining the loads of step s+1 under the arithmetic of step s needs ~170 registers = 3 CTAs per SM
    // and came out slower: 262 vs 258 us fp32, 274 vs 247 us bf16)
    for (int s = 0; s < lpp; ++s) {
      const int p = s * pps + grp;            // owner lane of the pixel my group handles in this step
      const int pidx = base + p;
      const bool plive = pidx < px_end;       // uniform over the group
      StepRegs R;
      issue_loads(s, R);
      const uint4 raw_t = R.raw_t;
      uint4 (&raw)[kV][4] = R.raw;
      int (&o_b)[kV] = R.o_b;
      unsigned (&pkv)[kV] = R.pkv;
      float4 (&cw)[kV] = R.cw;
      float tg[kVec], gt[kVec];
      VecIO<kBf16>::widen(raw_t, tg);
#pragma unroll
      for (int c = 0; c < kVec; ++c) gt[c] = 0.0f;

#pragma unroll
      for (int v = 0; v < kV; ++v) {
        const float4 c1 = *reinterpret_cast<const float4*>(&s_cell[warp][v][p][4]);   // {n, s, mask weight, -}
        Loc L;   // only the weights are used below
        L.w = cw[v].z;
        L.e = cw[v].w;
        L.n = c1.x;
        L.s = c1.y;
        const float ex = c1.z;
        const unsigned pk = pkv[v];
        float a0[kVec], a1[kVec], a2[kVec], a3[kVec];
        VecIO<kBf16>::widen(raw[v][0], a0);
        VecIO<kBf16>::widen(raw[v][1], a1);
        VecIO<kBf16>::widen(raw[v][2], a2);
        VecIO<kBf16>::widen(raw[v][3], a3);
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
        // value-based mask over ALL channels of the pixel: OR across the lanes of the group (one vote)
        any = ((__ballot_sync(kFull, any) >> (lane & ~(lpp - 1))) & grp_mask) != 0u;
        float lsum = 0.0f;
#pragma unroll
        for (int c = 0; c < kVec; ++c) lsum += fabsf(d1[c]);
        acc_loss[v] += any ? lsum : 0.0f;       // every lane adds its own channels

        if (need_grad) {
          if (kBf16) {
            // 8 channels per lane: do not keep the 32 widened tap values alive across the vote -- widen them again
            // from the raw vectors (a shift / a mask each); the empty asm stops the compiler from merging the copies
#pragma unroll
            for (int t = 0; t < 4; ++t)
              asm volatile("" : "+r"(raw[v][t].x), "+r"(raw[v][t].y), "+r"(raw[v][t].z), "+r"(raw[v][t].w));
            VecIO<kBf16>::widen(raw[v][0], a0);
            VecIO<kBf16>::widen(raw[v][1], a1);
            VecIO<kBf16>::widen(raw[v][2], a2);
            VecIO<kBf16>::widen(raw[v][3], a3);
          }
          float gx = 0.0f, gy = 0.0f, ge = 0.0f, g[kVec];
#pragma unroll
          for (int c = 0; c < kVec; ++c) {
            const float gd1 = signed_unit(d1[c], inv_n, any);
            g[c] = has_expl ? mul(gd1, ex) : gd1;
            if (has_expl) ge += gd1 * d0[c];
            gt[c] += g[c];
            bilerp_grad(a0[c], a1[c], a2[c], a3[c], L, -g[c], gx, gy);
          }
          group_sum_dispatch(lpp_shift, gx, gy, ge, has_expl);   // sums over all channels of the pixel
          if (gsrc_b[v]) {   // warp-uniform
            // Scatter: one 16-byte reduction per tap and quad of channels.  Neighbouring lane groups hold neighbouring
            // target pixels; where the sampling position advances by exactly one texel (same row, next column -- the usual
            // case) the east taps of a group are the west taps of the next one: it hands its two east contributions over
            // (shuffles) and the neighbour folds them into its own west reductions -- 4 -> ~2 reductions per pixel and view.
            const int o_cell = __float_as_int(cw[v].x), x0_cell = __float_as_int(c1.w);
            const int o_nx = __shfl_down_sync(kFull, o_cell, lpp), x_nx = __shfl_down_sync(kFull, x0_cell, lpp);
            const bool any_nx = __shfl_down_sync(kFull, (int)any, lpp) != 0;
            const bool give = (lane + lpp < 32) && any && any_nx && o_nx == o_cell + 1 && x_nx == x0_cell + 1;
            const bool take = (__shfl_up_sync(kFull, (int)give, lpp) != 0) && lane >= lpp;
            const int o_e = kBf16 ? o_b[v] * 2 : o_b[v];   // byte offset in the fp32 gradient map
            char* const gs = static_cast<char*>(static_cast<void*>(gsrc_b[v]));
#pragma unroll
            for (int q = 0; q < kVec; q += 4) {
              float cnw[4], cne[4], csw[4], cse[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                cnw[k] = -g[q + k] * wnw; cne[k] = -g[q + k] * wne; csw[k] = -g[q + k] * wsw; cse[k] = -g[q + k] * wse;
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float r_ne = __shfl_up_sync(kFull, cne[k], lpp), r_se = __shfl_up_sync(kFull, cse[k], lpp);
                if (take) {
                  cnw[k] += r_ne;
                  csw[k] += r_se;
                }
              }
              const int oq = o_e + (q / 4) * (C / 2) * 4;   // second group of 4 channels: C/2 further
              const unsigned on = any ? 1u : 0u, east = (any && !give) ? 1u : 0u;
              red_add_v4(gs, oq, (pk & 1u) * on, cnw[0], cnw[1], cnw[2], cnw[3]);
              red_add_v4(gs, oq + C * 4, (pk & 2u) * east, cne[0], cne[1], cne[2], cne[3]);
              red_add_v4(gs, oq + W * C * 4, (pk & 4u) * on, csw[0], csw[1], csw[2], csw[3]);
              red_add_v4(gs, oq + W * C * 4 + C * 4, (pk & 8u) * east, cse[0], cse[1], cse[2], cse[3]);
            }
          }
          // hand the folded sums to the owner lane of the pixel
          if ((lane & (lpp - 1)) == 0) *reinterpret_cast<float4*>(&s_back[warp][v][p][0]) = make_float4(gx, gy, ge, 0.0f);
        }
      }  // views
      if (kBf16 && grad_bf16 && need_grad && plive && gtgt_b) {
        char* q = static_cast<char*>(static_cast<void*>(gtgt_b)) + (size_t)pidx * px_b + ch_b;
#pragma unroll
        for (int c = 0; c < kVec; c += 4)
          *reinterpret_cast<uint2*>(q + (c / 4) * half_b) = make_uint2(pack_bf16x2(gt[c], gt[c + 1]), pack_bf16x2(gt[c + 2], gt[c + 3]));
      } else if (need_grad && plive && gtgt_b) {
        float* q = gtgt_b + (size_t)pidx * C + ch0;
#pragma unroll
        for (int c = 0; c < kVec; c += 4)
          *reinterpret_cast<float4*>(q + (c / 4) * (C / 2)) = make_float4(gt[c], gt[c + 1], gt[c + 2], gt[c + 3]);
      }
    }  // steps

    // ---- phase C: backward of the coordinate chain for my own pixel ------------------------------------
    __syncwarp();
    if (need_grad) {
      float gd = 0.0f;
#pragma unroll
      for (int v = 0; v < kV; ++v) {
        const float4 back = *reinterpret_cast<const float4*>(&s_back[warp][v][lane][0]);
        float P[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) P[k] = s_P[v][k];
        Proj pr;
        Loc L;
        const bool fast = project<false, kZeros>(P, cam, geo, pr) && allow_fast;   // same values as in phase A
        if (__builtin_expect(!fast, 0)) pr = project_exact<kZeros>(&s_P[v][0], cam, &lv.geo);
        locate<kZeros>(pr.xn, pr.yn, H, W, geo, L);
        ChainGrad cg;
        chain_backward<false>(P, cam, pr, L, back.x, back.y, geo, cg);
        if (__builtin_expect(!fast, 0)) cg = chain_backward_exact(&s_P[v][0], cam, pr, L, back.x, back.y, &lv.geo);
        if (live && lv.gexpl) st_stream(lv.gexpl + ((size_t)b * kV + v) * HW + idx, back.z);
        gd = add(gd, live ? cg.gdepth : 0.0f);
        // dL/dP of the 32 pixels: folded over the warp, added to the warp's sums (one owner lane per entry)
        float t[kRedSlots];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int k = 0; k < 3; ++k) t[r * 4 + k] = live ? cg.gq[r] * cam.cam[k] : 0.0f;
          t[r * 4 + 3] = live ? cg.gq[r] : 0.0f;
        }
        t[12] = t[13] = t[14] = t[15] = 0.0f;
        const float tot = butterfly16(t, lane);
        const int slot = butterfly_slot(lane);
        if ((lane & 1) == 0 && slot < 12) s_wacc[warp][v][slot] += tot;
      }
      if (live && lv.gdepth) {
        if (prm.disparity) gd = gdisp_of_gdepth(gd, depth_of_disp(ld_stream(depth_b + idx), prm.disp_eps));
        st_stream(lv.gdepth + img_px + idx, gd);
      }
    }
  }
  __syncwarp();
  float acc[kV][kRedSlots];
#pragma unroll
  for (int v = 0; v < kV; ++v) {
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[v][k] = lane == 0 ? s_wacc[warp][v][k] : 0.0f;
    acc[v][12] = acc_loss[v];
    acc[v][13] = acc[v][14] = acc[v][15] = 0.0f;
  }
  reduce_and_finish<kV, kLossThreads>(acc, prm, lv, l, part, n_parts, b, C);
  }  // pieces of this CTA
  pdl_wait_predecessor(prm);
}

template <int kV, bool kZeros>
void launch_loss_nhwc(const LossParams& prm, int blocks, bool bf16, cudaStream_t st);

}  // namespace dvf
