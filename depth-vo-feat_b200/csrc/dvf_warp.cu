// dvf_warp.cu -- the drop-in inverse_warp operator (materialised warped image) and its backward,
// plus the reference's stand-alone pixel2cam / cam2pixel.
//
// Replaces pytorch_version/inverse_warp.py: pixel2cam :26-40, cam2pixel :43-74, inverse_warp
// :160-193 (and F.grid_sample's grid_sampler_2d forward/backward, which the reference calls at
// :191).  The reference runs a K=3 bmm, ~25 elementwise kernels and a sampler per call; here one
// kernel per direction: thread = target pixel, every per-pixel load/store is coalesced, the 4
// bilinear taps per channel are read-only gathers served by L1/L2.
#include "dvf_internal.h"
#include "dvf_math.cuh"
#include "dvf_math2.cuh"
#include "dvf_reduce.cuh"

namespace dvf {

struct WarpParams {
  int B, C, H, W, HW;
  FastDiv divW;
  Geo geo;
  int allow_fast, zeros_padding;
  int ref_cuda;           // forward: torch-CUDA's per-pixel rounding (DVF_FLAG_REF_CUDA)
  int blocks_per_image;
  int bwd_blocks_per_image, bwd_iters;   // backward: a CTA walks bwd_iters chunks of 2*kThreads pixels (amortises its reduction)
  const float* img;
  const float* depth;
  const float* P;
  const float* Kinv;
  // forward
  float* warped;
  uint8_t* valid;
  // backward
  const float* gout;
  float* gdepth;
  float* gP;
  float* gimg;
  float* partials;        // [B*blocks_per_image][kRedSlots]
  unsigned* img_counter;  // [B]
};

constexpr int kWarpPPT = 2;

// cold, out-of-line exact versions (operands outside the guarded range of the shared-reciprocal divisions)
template <bool kZeros>
__device__ __noinline__ Proj warp_project_exact(const float* P, Cam c, Geo g) {
  Proj o;
  project<true, kZeros>(P, c, g, o);
  return o;
}
static __device__ __noinline__ ChainGrad warp_chain_backward_exact(const float* P, Cam c, Proj p, Loc L, float gx, float gy, Geo g) {
  ChainGrad o;
  chain_backward<true>(P, c, p, L, gx, gy, g, o);
  return o;
}

__device__ __forceinline__ void load_PM(const WarpParams& p, int b, float (&P)[12], float (&M)[9]) {
#pragma unroll
  for (int k = 0; k < 12; ++k) P[k] = __ldg(p.P + b * 12 + k);
#pragma unroll
  for (int k = 0; k < 9; ++k) M[k] = __ldg(p.Kinv + b * 9 + k);
}

// DVF_FLAG_REF_CUDA: the same pixel as torch-CUDA eager rounds it.  Differences to the torch-CPU chain (established on
// the B200 against torch 2.11+cu128, tests/test_gpu_api.py::test_ref_cuda_*): `2*(X/Z)/(w-1)` multiplies by the fp32
// reciprocal of the scalar (ATen BinaryDivTrueKernel.cu: a * (1 / b)), and grid_sample's CUDA kernel forms the weights as
// (x1 - ix) * (y1 - iy), ... from the corner coordinates instead of from w = ix - x0, e = 1 - w (GridSampler.cu).
template <bool kZeros>
__device__ __forceinline__ void ref_cuda_cell(const Proj& pr, const Geo& g, int H, int W, Loc& L, float (&wt)[4]) {
  float xn = sub(mul(add(pr.u, pr.u), g.rW1), 1.0f);   // g.rW1 = RN(1 / (w-1)) = torch's `1.0f / b`
  float yn = sub(mul(add(pr.v, pr.v), g.rH1), 1.0f);
  if (kZeros) {
    xn = fabsf(xn) > 1.0f ? 2.0f : xn;
    yn = fabsf(yn) > 1.0f ? 2.0f : yn;
  }
  locate<kZeros>(xn, yn, H, W, g, L);                  // cell and predicates; the un-normalisation is the same FMA
  float ix = fma_(add(xn, 1.0f), g.halfW, g.offs), iy = fma_(add(yn, 1.0f), g.halfH, g.offs);
  if (!kZeros) {
    ix = fminf(fmaxf(ix, 0.0f), g.fW1);
    iy = fminf(fmaxf(iy, 0.0f), g.fH1);
  }
  const float x0 = floorf(ix), y0 = floorf(iy), x1 = add(x0, 1.0f), y1 = add(y0, 1.0f);
  wt[0] = mul(sub(x1, ix), sub(y1, iy));   // nw
  wt[1] = mul(sub(ix, x0), sub(y1, iy));   // ne
  wt[2] = mul(sub(x1, ix), sub(iy, y0));   // sw
  wt[3] = mul(sub(ix, x0), sub(iy, y0));   // se
}

template <bool kZeros>
__global__ void __launch_bounds__(kThreads) inverse_warp_fwd_kernel(const __grid_constant__ WarpParams p) {
  const int b = blockIdx.x / p.blocks_per_image;
  const int chunk = blockIdx.x - b * p.blocks_per_image;
  float P[12], M[9];
  load_PM(p, b, P, M);
  const int HW = p.HW, W = p.W, H = p.H;
  const float* img_b = p.img + (size_t)b * p.C * HW;
  float* out_b = p.warped + (size_t)b * p.C * HW;
#pragma unroll
  for (int q = 0; q < kWarpPPT; ++q) {
    const int idx = chunk * (kThreads * kWarpPPT) + q * kThreads + threadIdx.x;
    if (idx >= HW) continue;
    const int i = (int)fastdiv((uint32_t)idx, p.divW), j = idx - i * W;
    Cam cam;
    Proj pr;
    Loc L;
    pixel_to_cam(M, ld_stream(p.depth + (size_t)b * HW + idx), i, j, cam);
    const bool fast = project<false, kZeros>(P, cam, p.geo, pr) && (p.allow_fast != 0);
    if (__builtin_expect(!fast, 0)) pr = warp_project_exact<kZeros>(p.P + b * 12, cam, p.geo);
    locate<kZeros>(pr.xn, pr.yn, H, W, p.geo, L);
    float wnw = mul(L.s, L.e), wne = mul(L.s, L.w), wsw = mul(L.n, L.e), wse = mul(L.n, L.w);
    if (p.ref_cuda) {   // launch-uniform
      float wt[4];
      ref_cuda_cell<kZeros>(pr, p.geo, H, W, L, wt);
      wnw = wt[0]; wne = wt[1]; wsw = wt[2]; wse = wt[3];
    }
    const int o_nw = L.y0 * W + L.x0;
    bool any = false;
    for (int c = 0; c < p.C; ++c) {
      const float* pl = img_b + (size_t)c * HW + o_nw;
      const float a0 = L.bnw ? __ldg(pl) : 0.0f, a1 = L.bne ? __ldg(pl + 1) : 0.0f;
      const float a2 = L.bsw ? __ldg(pl + W) : 0.0f, a3 = L.bse ? __ldg(pl + W + 1) : 0.0f;
      const float wv = bilerp(a0, a1, a2, a3, wnw, wne, wsw, wse);
      any |= (wv != 0.0f);
      st_stream(out_b + (size_t)c * HW + idx, wv);
    }
    if (p.valid) p.valid[(size_t)b * HW + idx] = any ? 1 : 0;
  }
}

// kCT: compile-time channel count (3 = images: the channel loops unroll and their loads overlap), 0 = run time
template <bool kZeros, int kCT>
__global__ void __launch_bounds__(kThreads) inverse_warp_bwd_kernel(const __grid_constant__ WarpParams p) {
  __shared__ float s_red[kThreads / 32][kRedSlots];
  __shared__ int s_flag;
  const int b = blockIdx.x / p.bwd_blocks_per_image;
  const int chunk0 = (blockIdx.x - b * p.bwd_blocks_per_image) * p.bwd_iters;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float P[12], M[9];
  load_PM(p, b, P, M);
  const int HW = p.HW, W = p.W, H = p.H, C = kCT ? kCT : p.C;
  const float* img_b = p.img + (size_t)b * C * HW;
  const float* gout_b = p.gout + (size_t)b * C * HW;
  float* gimg_b = p.gimg ? p.gimg + (size_t)b * C * HW : nullptr;
  float acc[kRedSlots];
#pragma unroll
  for (int k = 0; k < kRedSlots; ++k) acc[k] = 0.0f;

  for (int it = 0; it < p.bwd_iters * kWarpPPT; ++it) {
    const int idx_raw = chunk0 * (kThreads * kWarpPPT) + it * kThreads + tid;
    const bool live = idx_raw < HW;
    if (__all_sync(0xffffffffu, !live)) break;   // warp-uniform: the lanes of a warp stay together for the shuffles below
    const int idx = live ? idx_raw : HW - 1;
    const int i = (int)fastdiv((uint32_t)idx, p.divW), j = idx - i * W;
    Cam cam;
    Proj pr;
    Loc L;
    pixel_to_cam(M, ld_stream(p.depth + (size_t)b * HW + idx), i, j, cam);
    const bool fast = project<false, kZeros>(P, cam, p.geo, pr) && (p.allow_fast != 0);
    if (__builtin_expect(!fast, 0)) pr = warp_project_exact<kZeros>(p.P + b * 12, cam, p.geo);
    locate<kZeros>(pr.xn, pr.yn, H, W, p.geo, L);
    const int o_nw = L.y0 * W + L.x0;
    const float wnw = mul(L.s, L.e), wne = mul(L.s, L.w), wsw = mul(L.n, L.e), wse = mul(L.n, L.w);
    // Scatter of d img, warp-aggregated: neighbouring lanes are neighbouring target pixels, and where the sampling
    // positions advance by one texel (the usual case: smooth depth) the east taps of a lane are the west taps of the next
    // one.  Such a lane hands its two east contributions to its neighbour (one shuffle each) instead of issuing atomics for
    // them: 4 -> ~2 atomics per pixel and channel.
    bool give = false, take = false;
    if (gimg_b) {
      // same row, next column (comparing linear offsets would alias (W-1, r) + 1 with (-1, r+1))
      const int x_next = __shfl_down_sync(0xffffffffu, L.x0, 1), y_next = __shfl_down_sync(0xffffffffu, L.y0, 1);
      const bool live_next = __shfl_down_sync(0xffffffffu, (int)live, 1) != 0;
      give = lane < 31 && live && live_next && y_next == L.y0 && x_next == L.x0 + 1;
      take = __shfl_up_sync(0xffffffffu, (int)give, 1) != 0 && lane > 0;
    }
    float gx = 0.0f, gy = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float* pl = img_b + (size_t)c * HW + o_nw;
      const float g = live ? ld_stream(gout_b + (size_t)c * HW + idx) : 0.0f;
      const float a0 = L.bnw ? __ldg(pl) : 0.0f, a1 = L.bne ? __ldg(pl + 1) : 0.0f;
      const float a2 = L.bsw ? __ldg(pl + W) : 0.0f, a3 = L.bse ? __ldg(pl + W + 1) : 0.0f;
      bilerp_grad(a0, a1, a2, a3, L, g, gx, gy);
      if (gimg_b) {
        float* gp = gimg_b + (size_t)c * HW + o_nw;
        float c_nw = L.bnw ? mul(wnw, g) : 0.0f, c_sw = L.bsw ? mul(wsw, g) : 0.0f;
        const float c_ne = L.bne ? mul(wne, g) : 0.0f, c_se = L.bse ? mul(wse, g) : 0.0f;
        const float r_ne = __shfl_up_sync(0xffffffffu, c_ne, 1), r_se = __shfl_up_sync(0xffffffffu, c_se, 1);
        if (take) {   // same texels as my west taps (in bounds for both or for neither)
          c_nw += r_ne;
          c_sw += r_se;
        }
        if (live) {
          if (L.bnw) atomicAdd(gp, c_nw);
          if (L.bsw) atomicAdd(gp + W, c_sw);
          if (!give) {
            if (L.bne) atomicAdd(gp + 1, c_ne);
            if (L.bse) atomicAdd(gp + W + 1, c_se);
          }
        }
      }
    }
    ChainGrad cg;
    chain_backward<false>(P, cam, pr, L, gx, gy, p.geo, cg);
    if (__builtin_expect(!fast, 0)) cg = warp_chain_backward_exact(p.P + b * 12, cam, pr, L, gx, gy, p.geo);
    if (live) {
      st_stream(p.gdepth + (size_t)b * HW + idx, cg.gdepth);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[r * 4 + k] = fmaf(cg.gq[r], cam.cam[k], acc[r * 4 + k]);
        acc[r * 4 + 3] += cg.gq[r];
      }
    }
  }

  const float r = butterfly16(acc, lane);
  if ((lane & 1) == 0) s_red[warp][butterfly_slot(lane)] = r;
  __syncthreads();
  if (tid < kRedSlots) {
    float t = 0.0f;
#pragma unroll
    for (int w8 = 0; w8 < kThreads / 32; ++w8) t += s_red[w8][tid];
    __stcg(p.partials + (size_t)blockIdx.x * kRedSlots + tid, t);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = (atomicAdd(p.img_counter + b, 1u) == (unsigned)(p.bwd_blocks_per_image - 1));
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  const float* img_part = p.partials + (size_t)b * p.bwd_blocks_per_image * kRedSlots;
  if (tid < 12 * 8) {
    const int s = tid >> 3;
    const double sum = group8_sum(img_part + s, p.bwd_blocks_per_image, kRedSlots, tid & 7);
    if ((tid & 7) == 0) p.gP[(size_t)b * 12 + s] = (float)sum;
  }
  if (tid == 0) p.img_counter[b] = 0u;
}

// ---- C == 3, packed: a thread owns two adjacent pixels and runs the coordinate chain on (A,B) pairs with FFMA2 /
// FMUL2 / FADD2 (dvf_math2.cuh; every half rounded like the scalar op => same bits as the kernels above), 8-byte
// streaming loads / stores.  Needs HW even and 8-byte aligned depth / output planes (checked on the host).
struct PairCtx {
  Cam2 cam;
  Proj2 pr;
  Loc2 L;
  bool fastA, fastB;
};

template <bool kZeros>
__device__ __forceinline__ void pair_forward(const WarpParams& p, const float (&P)[12], const float (&M)[9], const Geo2& geo2,
                                             float depth_max, int b, int idxA, f2 dep, PairCtx& x) {
  const int W = p.W;
  const int iA = (int)fastdiv((uint32_t)idxA, p.divW), jA = idxA - iA * W;
  const int iB = (int)fastdiv((uint32_t)(idxA + 1), p.divW), jB = idxA + 1 - iB * W;
  pixel_to_cam2(M, dep, make_float2((float)iA, (float)iB), make_float2((float)jA, (float)jB), x.cam);
  x.fastA = fabsf(dep.x) <= depth_max;   // false for NaN
  x.fastB = fabsf(dep.y) <= depth_max;
  project2<kZeros>(P, x.cam, geo2, x.pr);
  if (__builtin_expect(!(x.fastA && x.fastB), 0)) {   // redo the offending lane(s) with exact divisions
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 0 ? x.fastA : x.fastB) continue;
      Cam c1;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        c1.ray[q] = h == 0 ? x.cam.ray[q].x : x.cam.ray[q].y;
        c1.cam[q] = h == 0 ? x.cam.cam[q].x : x.cam.cam[q].y;
      }
      const Proj e = warp_project_exact<kZeros>(p.P + b * 12, c1, p.geo);
      Proj2& pr = x.pr;
      if (h == 0) {
        pr.qz.x = e.qz; pr.nZ.x = -e.Z; pr.u.x = e.u; pr.v.x = e.v; pr.xn.x = e.xn; pr.yn.x = e.yn;
      } else {
        pr.qz.y = e.qz; pr.nZ.y = -e.Z; pr.u.y = e.u; pr.v.y = e.v; pr.xn.y = e.xn; pr.yn.y = e.yn;
      }
    }
  }
  locate2<kZeros>(x.pr.xn, x.pr.yn, p.H, W, p.geo, geo2, x.L);
}

// the four taps of one channel plane for both pixels; pa / pb point at the north-west texel of A / B
__device__ __forceinline__ void pair_taps(const Loc2& L, const float* pa, const float* pb, int W, f2& t00, f2& t01, f2& t10, f2& t11) {
  const float* pa1 = ptr_off(pa, W);
  const float* pb1 = ptr_off(pb, W);
  t00 = make_float2(L.nwA ? __ldg(pa) : 0.0f, L.nwB ? __ldg(pb) : 0.0f);
  t01 = make_float2(L.neA ? __ldg(pa + 1) : 0.0f, L.neB ? __ldg(pb + 1) : 0.0f);
  t10 = make_float2(L.swA ? __ldg(pa1) : 0.0f, L.swB ? __ldg(pb1) : 0.0f);
  t11 = make_float2(L.seA ? __ldg(pa1 + 1) : 0.0f, L.seB ? __ldg(pb1 + 1) : 0.0f);
}

__device__ __forceinline__ float pair_depth_max(const WarpParams& p, const float (&P)[12], const float (&M)[9]) {
  bool ok = p.allow_fast != 0;
#pragma unroll
  for (int k = 0; k < 9; ++k) ok = ok && (fabsf(M[k]) <= 1048576.0f);
#pragma unroll
  for (int k = 0; k < 12; ++k) ok = ok && (fabsf(P[k]) <= 1073741824.0f);
  return ok ? 1073741824.0f : -1.0f;   // same guard as the fused loss kernel (dvf_loss_kernel.cuh)
}

__device__ __forceinline__ f2 ld_stream2(const float* p) {
  f2 v;
  asm volatile("ld.global.cs.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream2(float* p, f2 v) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

template <bool kZeros>
__global__ void __launch_bounds__(kThreads) inverse_warp_fwd_c3x2_kernel(const __grid_constant__ WarpParams p) {
  const int b = blockIdx.x / p.blocks_per_image;
  const int chunk = blockIdx.x - b * p.blocks_per_image;
  const int HW = p.HW, W = p.W;
  const int idxA = chunk * (kThreads * 2) + 2 * threadIdx.x;
  if (idxA >= HW) return;   // HW is even: a pair is never split by the end of the image
  float P[12], M[9];
  load_PM(p, b, P, M);
  const Geo2 geo2 = make_geo2(p.geo);
  const float depth_max = pair_depth_max(p, P, M);
  const float* img_b = p.img + (size_t)b * 3 * HW;
  float* out_b = p.warped + (size_t)b * 3 * HW;
  PairCtx x;
  pair_forward<kZeros>(p, P, M, geo2, depth_max, b, idxA, ld_stream2(p.depth + (size_t)b * HW + idxA), x);
  const Loc2& L = x.L;
  const f2 wnw = mul2(L.s, L.e), wne = mul2(L.s, L.w), wsw = mul2(L.n, L.e), wse = mul2(L.n, L.w);
  const float* pa = ptr_off(img_b, L.y0A * W + L.x0A);
  const float* pb = ptr_off(img_b, L.y0B * W + L.x0B);
  bool anyA = false, anyB = false;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    f2 t00, t01, t10, t11;
    pair_taps(L, pa, pb, W, t00, t01, t10, t11);
    const f2 wv = bilerp2(t00, t01, t10, t11, wnw, wne, wsw, wse);
    anyA |= (wv.x != 0.0f);
    anyB |= (wv.y != 0.0f);
    st_stream2(out_b + (size_t)c * HW + idxA, wv);
    pa = ptr_off(pa, HW);
    pb = ptr_off(pb, HW);
  }
  if (p.valid) *reinterpret_cast<uchar2*>(p.valid + (size_t)b * HW + idxA) = make_uchar2(anyA ? 1 : 0, anyB ? 1 : 0);
}

// kScatter: also accumulate d img (p.gimg, zero-filled by the caller).  A thread's two pixels are horizontal neighbours and
// so are the pairs of neighbouring lanes: where the sampling positions advance by one texel per pixel (smooth depth, the
// usual case) the east taps of a pixel are the west taps of the next one, so their contributions are added in registers
// (inside the pair) or handed over by one shuffle (between lanes) before anything goes to memory: ~1 atomic per pixel, row
// and channel instead of 2 -- and nothing at all changes for the arithmetic of d depth / d P.
template <bool kZeros, bool kScatter>
__global__ void __launch_bounds__(kThreads) inverse_warp_bwd_c3x2_kernel(const __grid_constant__ WarpParams p) {
  __shared__ float s_red[kThreads / 32][kRedSlots];
  __shared__ int s_flag;
  const int b = blockIdx.x / p.bwd_blocks_per_image;
  const int chunk0 = (blockIdx.x - b * p.bwd_blocks_per_image) * p.bwd_iters;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HW = p.HW, W = p.W;
  float acc[kRedSlots];
#pragma unroll
  for (int k = 0; k < kRedSlots; ++k) acc[k] = 0.0f;
  float P[12], M[9];
  load_PM(p, b, P, M);
  const Geo2 geo2 = make_geo2(p.geo);
  const float depth_max = pair_depth_max(p, P, M);
  const float* img_b = p.img + (size_t)b * 3 * HW;
  float* const gimg_b = kScatter ? p.gimg + (size_t)b * 3 * HW : nullptr;
#pragma unroll 1
  for (int it = 0; it < p.bwd_iters; ++it) {
    const int idx_raw = (chunk0 + it) * (kThreads * 2) + 2 * tid;
    const bool live = idx_raw < HW;               // HW is even: a pair is live or dead as a whole
    if (kScatter ? __all_sync(0xffffffffu, !live) : !live) break;   // kScatter: lanes stay together for the shuffles
    const int idxA = live ? idx_raw : HW - 2;
    const float* gout_b = p.gout + (size_t)b * 3 * HW + idxA;
    PairCtx x;
    pair_forward<kZeros>(p, P, M, geo2, depth_max, b, idxA, ld_stream2(p.depth + (size_t)b * HW + idxA), x);
    const Loc2& L = x.L;
    const int oA = L.y0A * W + L.x0A, oB = L.y0B * W + L.x0B;
    const float* pa = ptr_off(img_b, oA);
    const float* pb = ptr_off(img_b, oB);
    bool adjAB = false, give = false, take = false;
    f2 wnw, wne, wsw, wse;
    if (kScatter) {
      // same row, next column (comparing linear offsets would alias (W-1, r) + 1 with (-1, r+1))
      adjAB = L.y0B == L.y0A && L.x0B == L.x0A + 1;             // B's west taps are A's east taps
      const int xA_next = __shfl_down_sync(0xffffffffu, L.x0A, 1), yA_next = __shfl_down_sync(0xffffffffu, L.y0A, 1);
      const bool live_next = __shfl_down_sync(0xffffffffu, (int)live, 1) != 0;
      give = lane < 31 && live && live_next && yA_next == L.y0B && xA_next == L.x0B + 1;   // my B's east = the next lane's A's west
      take = __shfl_up_sync(0xffffffffu, (int)give, 1) != 0 && lane > 0;
      wnw = mul2(L.s, L.e); wne = mul2(L.s, L.w); wsw = mul2(L.n, L.e); wse = mul2(L.n, L.w);
    }
    f2 gx = dup(0.0f), gy = dup(0.0f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      f2 t00, t01, t10, t11;
      pair_taps(L, pa, pb, W, t00, t01, t10, t11);
      const f2 g = ld_stream2(gout_b + (size_t)c * HW);
      bilerp_grad2(t00, t01, t10, t11, L, g, gx, gy);
      if (kScatter) {
        // contributions of the 8 taps (zero where the texel is outside the image)
        const f2 cnw = mul2(wnw, g), cne = mul2(wne, g), csw = mul2(wsw, g), cse = mul2(wse, g);
        float a_nw = L.nwA ? cnw.x : 0.0f, a_sw = L.swA ? csw.x : 0.0f;
        const float a_ne = L.neA ? cne.x : 0.0f, a_se = L.seA ? cse.x : 0.0f;
        const float b_nw = L.nwB ? cnw.y : 0.0f, b_sw = L.swB ? csw.y : 0.0f;
        const float b_ne = L.neB ? cne.y : 0.0f, b_se = L.seB ? cse.y : 0.0f;
        const float r_ne = __shfl_up_sync(0xffffffffu, b_ne, 1), r_se = __shfl_up_sync(0xffffffffu, b_se, 1);
        if (take) {
          a_nw += r_ne;
          a_sw += r_se;
        }
        if (live) {
          float* ga = ptr_off(gimg_b + (size_t)c * HW, oA);
          float* gb = ptr_off(gimg_b + (size_t)c * HW, oB);
          if (L.nwA) atomicAdd(ga, a_nw);
          if (L.swA) atomicAdd(ga + W, a_sw);
          if (adjAB) {   // texel oA+1 == oB: one atomic for A's east and B's west tap (both in bounds or neither)
            if (L.nwB) atomicAdd(gb, a_ne + b_nw);
            if (L.swB) atomicAdd(gb + W, a_se + b_sw);
          } else {
            if (L.neA) atomicAdd(ga + 1, a_ne);
            if (L.seA) atomicAdd(ga + W + 1, a_se);
            if (L.nwB) atomicAdd(gb, b_nw);
            if (L.swB) atomicAdd(gb + W, b_sw);
          }
          if (!give) {
            if (L.neB) atomicAdd(gb + 1, b_ne);
            if (L.seB) atomicAdd(gb + W + 1, b_se);
          }
        }
      }
      pa = ptr_off(pa, HW);
      pb = ptr_off(pb, HW);
    }
    ChainGrad2 cg;
    chain_backward2<kZeros>(P, x.cam, x.pr, L, gx, gy, geo2, cg);
    if (__builtin_expect(!(x.fastA && x.fastB), 0)) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 0 ? x.fastA : x.fastB) continue;
        Cam c1;
        Proj p1;
        Loc L1;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          c1.ray[q] = h == 0 ? x.cam.ray[q].x : x.cam.ray[q].y;
          c1.cam[q] = h == 0 ? x.cam.cam[q].x : x.cam.cam[q].y;
        }
        const Proj2& pr = x.pr;
        p1.qz = h == 0 ? pr.qz.x : pr.qz.y;
        p1.Z = h == 0 ? -pr.nZ.x : -pr.nZ.y;
        p1.rZ = 0.0f;
        p1.u = h == 0 ? pr.u.x : pr.u.y;
        p1.v = h == 0 ? pr.v.x : pr.v.y;
        p1.xn = h == 0 ? pr.xn.x : pr.xn.y;
        p1.yn = h == 0 ? pr.yn.x : pr.yn.y;
        p1.mx = kZeros && (h == 0 ? pr.xn.x : pr.xn.y) == 2.0f;
        p1.my = kZeros && (h == 0 ? pr.yn.x : pr.yn.y) == 2.0f;
        L1.x0 = L1.y0 = 0;
        L1.w = L1.e = L1.n = L1.s = 0.0f;
        L1.bnw = L1.bne = L1.bsw = L1.bse = false;
        L1.gmx = h == 0 ? L.gmx.x : L.gmx.y;
        L1.gmy = h == 0 ? L.gmy.x : L.gmy.y;
        const ChainGrad e = warp_chain_backward_exact(p.P + b * 12, c1, p1, L1, h == 0 ? gx.x : gx.y, h == 0 ? gy.x : gy.y, p.geo);
        if (h == 0) {
          cg.gq[0].x = e.gq[0]; cg.gq[1].x = e.gq[1]; cg.gq[2].x = e.gq[2]; cg.gdepth.x = e.gdepth;
        } else {
          cg.gq[0].y = e.gq[0]; cg.gq[1].y = e.gq[1]; cg.gq[2].y = e.gq[2]; cg.gdepth.y = e.gdepth;
        }
      }
    }
    if (!live) continue;
    st_stream2(p.gdepth + (size_t)b * HW + idxA, cg.gdepth);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const f2 t = mul2(cg.gq[r], x.cam.cam[k]);   // the scalar kernel accumulates with fmaf per pixel; the order
        acc[r * 4 + k] += t.x + t.y;                 // of a sum over pixels is free (dP is compared at 1e-5)
      }
      acc[r * 4 + 3] += cg.gq[r].x + cg.gq[r].y;
    }
  }

  const float r = butterfly16(acc, lane);
  if ((lane & 1) == 0) s_red[warp][butterfly_slot(lane)] = r;
  __syncthreads();
  if (tid < kRedSlots) {
    float t = 0.0f;
#pragma unroll
    for (int w8 = 0; w8 < kThreads / 32; ++w8) t += s_red[w8][tid];
    __stcg(p.partials + (size_t)blockIdx.x * kRedSlots + tid, t);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = (atomicAdd(p.img_counter + b, 1u) == (unsigned)(p.bwd_blocks_per_image - 1));
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  const float* img_part = p.partials + (size_t)b * p.bwd_blocks_per_image * kRedSlots;
  if (tid < 12 * 8) {
    const int s = tid >> 3;
    const double sum = group8_sum(img_part + s, p.bwd_blocks_per_image, kRedSlots, tid & 7);
    if ((tid & 7) == 0) p.gP[(size_t)b * 12 + s] = (float)sum;
  }
  if (tid == 0) p.img_counter[b] = 0u;
}

// ---- stand-alone pixel2cam / cam2pixel ---------------------------------------------------------
__global__ void __launch_bounds__(kThreads) pixel2cam_kernel(const float* __restrict__ depth,
                                                             const float* __restrict__ Kinv, int HW, FastDiv divW,
                                                             int W, float* __restrict__ cam) {
  const int b = blockIdx.y;
  const int idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= HW) return;
  float M[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) M[k] = __ldg(Kinv + b * 9 + k);
  const int i = (int)fastdiv((uint32_t)idx, divW), j = idx - i * W;
  Cam c;
  pixel_to_cam(M, depth[(size_t)b * HW + idx], i, j, c);
#pragma unroll
  for (int k = 0; k < 3; ++k) cam[((size_t)b * 3 + k) * HW + idx] = c.cam[k];
}

__global__ void __launch_bounds__(kThreads) cam2pixel_kernel(const float* __restrict__ cam, const float* __restrict__ rot,
                                                             const float* __restrict__ tr, int H, int W, int HW,
                                                             int zeros, float* __restrict__ grid) {
  const int b = blockIdx.y;
  const int idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= HW) return;
  float c[3], q[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) c[k] = cam[((size_t)b * 3 + k) * HW + idx];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // inverse_warp.py:54-60: both the rotation and the translation are optional
    float v = rot ? dot3(__ldg(rot + b * 9 + k * 3), __ldg(rot + b * 9 + k * 3 + 1), __ldg(rot + b * 9 + k * 3 + 2), c[0], c[1], c[2])
                  : c[k];
    q[k] = tr ? add(v, __ldg(tr + b * 3 + k)) : v;
  }
  const float Z = (q[2] < kMinDepthZ) ? kMinDepthZ : q[2];
  float xn = sub(div(mul(2.0f, div(q[0], Z)), (float)(W - 1)), 1.0f);
  float yn = sub(div(mul(2.0f, div(q[1], Z)), (float)(H - 1)), 1.0f);
  if (zeros) {
    if (xn > 1.0f || xn < -1.0f) xn = 2.0f;
    if (yn > 1.0f || yn < -1.0f) yn = 2.0f;
  }
  reinterpret_cast<float2*>(grid)[(size_t)b * HW + idx] = make_float2(xn, yn);
}

// backward of pixel2cam w.r.t. depth: autograd of `cam = ray * depth.unsqueeze(1)` is (g_cam * ray).sum(1), left to right
__global__ void __launch_bounds__(kThreads) pixel2cam_bwd_kernel(const float* __restrict__ gcam, const float* __restrict__ Kinv,
                                                                 int HW, FastDiv divW, int W, float* __restrict__ gdepth) {
  const int b = blockIdx.y;
  const int idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= HW) return;
  float M[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) M[k] = __ldg(Kinv + b * 9 + k);
  const int i = (int)fastdiv((uint32_t)idx, divW), j = idx - i * W;
  Cam c;
  pixel_to_cam(M, 1.0f, i, j, c);
  float g = mul(gcam[((size_t)b * 3 + 0) * HW + idx], c.ray[0]);
  g = add(g, mul(gcam[((size_t)b * 3 + 1) * HW + idx], c.ray[1]));
  g = add(g, mul(gcam[((size_t)b * 3 + 2) * HW + idx], c.ray[2]));
  gdepth[(size_t)b * HW + idx] = g;
}

// backward of cam2pixel: g_grid [B,H,W,2] -> g_cam [B,3,H,W] (written), g_rot [B,3,3] and g_tr [B,3] (zero-filled by the
// entry, accumulated: one shuffle fold per warp, one atomic per warp and entry).  Same fp32 sequence as autograd's:
// overwritten coordinates pass no gradient, /(w-1), *2, d(X/Z), clamp pass-through, rot^T.
__global__ void __launch_bounds__(kThreads) cam2pixel_bwd_kernel(const float* __restrict__ ggrid, const float* __restrict__ cam,
                                                                 const float* __restrict__ rot, const float* __restrict__ tr,
                                                                 int H, int W, int HW, int zeros, float* __restrict__ gcam,
                                                                 float* __restrict__ grot, float* __restrict__ gtr) {
  const int b = blockIdx.y;
  const int idx = blockIdx.x * kThreads + threadIdx.x;
  const bool live = idx < HW;
  float c[3] = {0.0f, 0.0f, 1.0f}, q[3], gq[3] = {0.0f, 0.0f, 0.0f};
  float R[9] = {1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f};
  if (rot) {
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = __ldg(rot + b * 9 + k);
  }
  if (live) {
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = cam[((size_t)b * 3 + k) * HW + idx];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float v = rot ? dot3(R[k * 3], R[k * 3 + 1], R[k * 3 + 2], c[0], c[1], c[2]) : c[k];
      q[k] = tr ? add(v, __ldg(tr + b * 3 + k)) : v;
    }
    const float Z = (q[2] < kMinDepthZ) ? kMinDepthZ : q[2];
    const float u = div(q[0], Z), v = div(q[1], Z);
    const float xn = sub(div(mul(2.0f, u), (float)(W - 1)), 1.0f);
    const float yn = sub(div(mul(2.0f, v), (float)(H - 1)), 1.0f);
    const float2 g = reinterpret_cast<const float2*>(ggrid)[(size_t)b * HW + idx];
    const float gxn = (zeros && (xn > 1.0f || xn < -1.0f)) ? 0.0f : g.x;
    const float gyn = (zeros && (yn > 1.0f || yn < -1.0f)) ? 0.0f : g.y;
    const float gu = mul(div(gxn, (float)(W - 1)), 2.0f), gv = mul(div(gyn, (float)(H - 1)), 2.0f);
    gq[0] = div(gu, Z);
    gq[1] = div(gv, Z);
    const float gZ = add(mul(-gu, div(u, Z)), mul(-gv, div(v, Z)));
    gq[2] = (q[2] >= kMinDepthZ) ? gZ : 0.0f;
    if (gcam) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        gcam[((size_t)b * 3 + k) * HW + idx] = rot ? dot3(R[k], R[3 + k], R[6 + k], gq[0], gq[1], gq[2]) : gq[k];
    }
  }
  if (grot || gtr) {
    float t[12];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int k = 0; k < 3; ++k) t[r * 4 + k] = live ? gq[r] * c[k] : 0.0f;
      t[r * 4 + 3] = live ? gq[r] : 0.0f;
    }
#pragma unroll
    for (int s = 0; s < 12; ++s) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t[s] += __shfl_xor_sync(0xffffffffu, t[s], o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        if (grot) {
#pragma unroll
          for (int k = 0; k < 3; ++k) atomicAdd(grot + b * 9 + r * 3 + k, t[r * 4 + k]);
        }
        if (gtr) atomicAdd(gtr + b * 3 + r, t[r * 4 + 3]);
      }
    }
  }
}

static int fill_params(const dvf_desc* d, WarpParams& p) {
  if (!d) return DVF_EINVAL_NULL;
  if (d->B <= 0 || d->C <= 0 || d->H <= 0 || d->W <= 0 || (long long)d->H * d->W >= (1ll << 30)) return DVF_EINVAL_SHAPE;
  if (d->padding != DVF_PAD_ZEROS && d->padding != DVF_PAD_BORDER) return DVF_EINVAL_DTYPE;
  if (d->dtype != DVF_F32 && d->dtype != DVF_BF16) return DVF_EINVAL_DTYPE;
  if (d->layout != DVF_NCHW && d->layout != DVF_NHWC) return DVF_EINVAL_DTYPE;
  if (d->dtype != DVF_F32 || d->layout != DVF_NCHW) return DVF_EUNSUPPORTED;
  p.B = d->B;
  p.C = d->C;
  p.H = d->H;
  p.W = d->W;
  p.HW = d->H * d->W;
  p.divW = make_fastdiv((uint32_t)d->W);
  p.geo = make_geo(d->H, d->W, (d->flags & DVF_FLAG_ALIGN_CORNERS) != 0);
  p.allow_fast = d->W > 1 && d->H > 1 && d->W - 1 <= kMaxConstDiv && d->H - 1 <= kMaxConstDiv;
  p.zeros_padding = d->padding == DVF_PAD_ZEROS;
  p.ref_cuda = (d->flags & DVF_FLAG_REF_CUDA) != 0;
  p.blocks_per_image = (p.HW + kThreads * kWarpPPT - 1) / (kThreads * kWarpPPT);
  // backward: about 8 CTAs per SM over the whole batch, each walking several chunks, so that the CTA reduction and
  // its ticket are paid once per few thousand pixels
  int want = (num_sms() * 8 + p.B - 1) / p.B;
  if (want < 1) want = 1;
  if (want > p.blocks_per_image) want = p.blocks_per_image;
  p.bwd_iters = (p.blocks_per_image + want - 1) / want;
  p.bwd_blocks_per_image = (p.blocks_per_image + p.bwd_iters - 1) / p.bwd_iters;
  return DVF_OK;
}

}  // namespace dvf

using namespace dvf;

DVF_EXPORT int dvf_inverse_warp_fwd(const dvf_desc* d, const void* img, const float* depth, const float* P,
                                    const float* Kinv, void* warped, uint8_t* valid, void* stream) {
  WarpParams p = {};
  int st = fill_params(d, p);
  if (st != DVF_OK) return st;
  if (!img || !depth || !P || !Kinv || !warped) return DVF_EINVAL_NULL;
  if (!aligned(img, 4) || !aligned(depth, 4) || !aligned(P, 4) || !aligned(Kinv, 4) || !aligned(warped, 4)) return DVF_EINVAL_ALIGN;
  p.img = static_cast<const float*>(img);
  p.depth = depth;
  p.P = P;
  p.Kinv = Kinv;
  p.warped = static_cast<float*>(warped);
  p.valid = valid;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)(p.blocks_per_image * p.B);
  // images: packed pixel pairs (needs an even HW and 8-byte aligned planes); anything else: generic kernel
  const bool packed = !p.ref_cuda && p.C == 3 && p.HW % 2 == 0 && aligned(depth, 8) && aligned(warped, 8) && (!valid || aligned(valid, 2));
  if (packed) {
    if (p.zeros_padding) inverse_warp_fwd_c3x2_kernel<true><<<grid, kThreads, 0, cs>>>(p);
    else inverse_warp_fwd_c3x2_kernel<false><<<grid, kThreads, 0, cs>>>(p);
  } else {
    if (p.zeros_padding) inverse_warp_fwd_kernel<true><<<grid, kThreads, 0, cs>>>(p);
    else inverse_warp_fwd_kernel<false><<<grid, kThreads, 0, cs>>>(p);
  }
  return launch_status();
}

// Workspace = [partials + counters of the kernels of this file][workspace of the image-kernel route]: the two routes never
// share a byte, so one buffer serves whichever route a call takes (both leave their counters zero).
static size_t own_bwd_workspace(const WarpParams& p) {
  return align_up((size_t)p.blocks_per_image * p.B * kRedSlots * sizeof(float), 256) + align_up((size_t)p.B * sizeof(unsigned), 256);
}

DVF_EXPORT size_t dvf_inverse_warp_bwd_workspace_bytes(const dvf_desc* d) {
  WarpParams p = {};
  if (fill_params(d, p) != DVF_OK) return 0;
  return own_bwd_workspace(p) + (p.C == 3 ? warp_bwd_fused_workspace_bytes(d) : 0);
}

DVF_EXPORT int dvf_inverse_warp_bwd(const dvf_desc* d, const void* gout, const void* img, const float* depth,
                                    const float* P, const float* Kinv, float* gdepth, float* gP, void* gimg,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  WarpParams p = {};
  int st = fill_params(d, p);
  if (st != DVF_OK) return st;
  if (p.ref_cuda) return DVF_EUNSUPPORTED;   // gradients follow torch-CPU (include/dvf_b200.h)
  if (!gout || !img || !depth || !P || !Kinv || !gdepth || !gP) return DVF_EINVAL_NULL;
  if (!aligned(gout, 4) || !aligned(img, 4) || !aligned(depth, 4) || !aligned(gdepth, 4) || !aligned(gP, 4)) return DVF_EINVAL_ALIGN;
  const size_t need = dvf_inverse_warp_bwd_workspace_bytes(d);
  if (!workspace || workspace_bytes < need) return DVF_EWORKSPACE;
  if (!aligned(workspace, 256)) return DVF_EINVAL_ALIGN;
  // images without d img: the fused loss kernel fed with the upstream gradient (ring of bulk copies, balanced persistent
  // split: 59 -> see profiles/r2_summary.md); anything it does not cover stays on the kernels of this file
  if (!gimg && warp_bwd_fused_ok(d, gout, img, depth) && aligned(gdepth, 8)) {
    const size_t own = own_bwd_workspace(p);
    return warp_bwd_fused(d, gout, img, depth, P, Kinv, gdepth, gP, static_cast<char*>(workspace) + own, need - own, stream);
  }
  p.img = static_cast<const float*>(img);
  p.gout = static_cast<const float*>(gout);
  p.depth = depth;
  p.P = P;
  p.Kinv = Kinv;
  p.gdepth = gdepth;
  p.gP = gP;
  p.gimg = static_cast<float*>(gimg);
  p.partials = static_cast<float*>(workspace);
  p.img_counter = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) +
                                              align_up((size_t)p.blocks_per_image * p.B * kRedSlots * sizeof(float), 256));
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)(p.bwd_blocks_per_image * p.B);
  const bool packed = p.C == 3 && p.HW % 2 == 0 && aligned(depth, 8) && aligned(gdepth, 8) && aligned(gout, 8);
  if (packed && p.gimg) {
    if (p.zeros_padding) inverse_warp_bwd_c3x2_kernel<true, true><<<grid, kThreads, 0, cs>>>(p);
    else inverse_warp_bwd_c3x2_kernel<false, true><<<grid, kThreads, 0, cs>>>(p);
  } else if (packed) {
    if (p.zeros_padding) inverse_warp_bwd_c3x2_kernel<true, false><<<grid, kThreads, 0, cs>>>(p);
    else inverse_warp_bwd_c3x2_kernel<false, false><<<grid, kThreads, 0, cs>>>(p);
  } else {
    if (p.C == 3) {
      if (p.zeros_padding) inverse_warp_bwd_kernel<true, 3><<<grid, kThreads, 0, cs>>>(p);
      else inverse_warp_bwd_kernel<false, 3><<<grid, kThreads, 0, cs>>>(p);
    } else {
      if (p.zeros_padding) inverse_warp_bwd_kernel<true, 0><<<grid, kThreads, 0, cs>>>(p);
      else inverse_warp_bwd_kernel<false, 0><<<grid, kThreads, 0, cs>>>(p);
    }
  }
  return launch_status();
}

DVF_EXPORT int dvf_pixel2cam(const float* depth, const float* Kinv, int32_t B, int32_t H, int32_t W, float* cam,
                             void* stream) {
  if (!depth || !Kinv || !cam) return DVF_EINVAL_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || (long long)H * W >= (1ll << 30)) return DVF_EINVAL_SHAPE;
  const int HW = H * W;
  dim3 grid((HW + kThreads - 1) / kThreads, B);
  pixel2cam_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(depth, Kinv, HW, make_fastdiv((uint32_t)W), W, cam);
  return launch_status();
}

DVF_EXPORT int dvf_cam2pixel(const float* cam, const float* rot, const float* tr, int32_t B, int32_t H, int32_t W,
                             int32_t padding, float* grid_out, void* stream) {
  if (!cam || !grid_out) return DVF_EINVAL_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || (long long)H * W >= (1ll << 30)) return DVF_EINVAL_SHAPE;
  if (padding != DVF_PAD_ZEROS && padding != DVF_PAD_BORDER) return DVF_EINVAL_DTYPE;
  if (!aligned(grid_out, 8)) return DVF_EINVAL_ALIGN;
  const int HW = H * W;
  dim3 grid((HW + kThreads - 1) / kThreads, B);
  cam2pixel_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(cam, rot, tr, H, W, HW, padding == DVF_PAD_ZEROS, grid_out);
  return launch_status();
}

DVF_EXPORT int dvf_pixel2cam_bwd(const float* gcam, const float* Kinv, int32_t B, int32_t H, int32_t W, float* gdepth,
                                 void* stream) {
  if (!gcam || !Kinv || !gdepth) return DVF_EINVAL_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || (long long)H * W >= (1ll << 30)) return DVF_EINVAL_SHAPE;
  const int HW = H * W;
  dim3 grid((HW + kThreads - 1) / kThreads, B);
  pixel2cam_bwd_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(gcam, Kinv, HW, make_fastdiv((uint32_t)W), W, gdepth);
  return launch_status();
}

DVF_EXPORT int dvf_cam2pixel_bwd(const float* ggrid, const float* cam, const float* rot, const float* tr, int32_t B, int32_t H,
                                 int32_t W, int32_t padding, float* gcam, float* grot, float* gtr, void* stream) {
  if (!ggrid || !cam) return DVF_EINVAL_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || (long long)H * W >= (1ll << 30)) return DVF_EINVAL_SHAPE;
  if (padding != DVF_PAD_ZEROS && padding != DVF_PAD_BORDER) return DVF_EINVAL_DTYPE;
  if (!aligned(ggrid, 8)) return DVF_EINVAL_ALIGN;
  if (grot && !rot) return DVF_EINVAL_NULL;
  if (gtr && !tr) return DVF_EINVAL_NULL;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  if (grot && cudaMemsetAsync(grot, 0, sizeof(float) * 9 * B, cs) != cudaSuccess) return launch_status();
  if (gtr && cudaMemsetAsync(gtr, 0, sizeof(float) * 3 * B, cs) != cudaSuccess) return launch_status();
  const int HW = H * W;
  dim3 grid((HW + kThreads - 1) / kThreads, B);
  cam2pixel_bwd_kernel<<<grid, kThreads, 0, cs>>>(ggrid, cam, rot, tr, H, W, HW, padding == DVF_PAD_ZEROS, gcam, grot, gtr);
  return launch_status();
}
