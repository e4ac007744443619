// dvf_reg.cu -- the regularisers either side of the reconstruction loss:
//   * smooth_loss       loss_functions.py:23-41 / loss_functions_sfm.py:59-77: per scale, the sum of the means of
//                       |dxx|, |dxy|, |dyx|, |dyy| of the depth map, scales weighted by 1/scale_factor^k;
//   * explainability_loss  loss_functions_sfm.py:49-56: per scale, binary_cross_entropy(mask, 1) = mean(-log mask).
// The reference builds each from ~15 slicing / elementwise / reduction launches per scale and lets autograd chain
// through them.  Here ONE launch handles every scale, forward and backward together: a thread owns one pixel,
// evaluates the stencils it belongs to from a 5x5 neighbourhood (read-only, L1-resident), accumulates the loss
// terms, and writes its own gradient (gather form -- no atomics).  The differences are taken in the reference's
// order (difference of rounded first differences).  Block partials are folded by the last CTA in fp64 in a fixed
// order (same ticket scheme as the loss kernel).  HBM-bound: 4 B read + 4 B written per pixel.
#include "dvf_internal.h"
#include "dvf_math.cuh"

namespace dvf {

constexpr int kRegMaxLevels = DVF_MAX_LEVELS;

struct RegLevel {
  const float* x;     // [B, H, W] (smooth: depth map; expl: mask with all its channels folded into B)
  float* g;           // gradient, same shape, written (nullable)
  int B, H, W;
  float weight;       // upstream weight of this level (already includes 1/scale_factor^k)
  float ixx, iyy, ixy; // smooth: 1 / element count of the dxx, dyy and cross terms (IEEE division, as mean() does)
  uint32_t items;      // smooth: B * ceil(H/kStrip) * W work items
  FastDiv divSW, divW; // smooth: item -> (image, strip row, column)
  int block_begin;
};
struct RegParams {
  int n_levels, total_blocks;   // total_blocks = CTAs launched (= partials)
  int virtual_blocks;           // kThreads-element blocks of work, all levels
  RegLevel lv[kRegMaxLevels];
  double* partials;   // [total_blocks]
  unsigned* counter;  // [1], zero between launches
  float* out;         // [1] loss value
};

// sign(v) in {-1, 0, +1}, 0 for NaN.  The smoothness kernel evaluates 44 of these per 4-pixel strip and is bound by
// instruction issue (78 % of the issue slots busy, profiles/r2_summary.md).  Two instructions: FSET.BF.NE gives 1.0f / 0.0f
// for an ORDERED v != 0 (false for NaN) and LOP3 copies the sign bit of v onto it.  (Round 1 spent compare + compare +
// select + select on the ALU pipe; then two saturating multiplies by 2^100 per half on the FMA pipe, five instructions.)
__device__ __forceinline__ float sgnf(float v) {
  float nz, r;
  asm("set.ne.f32.f32 %0, %1, 0f00000000;" : "=f"(nz) : "f"(v));
  asm("copysign.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "f"(nz));
  return r;
}

template <typename F>
__device__ __forceinline__ void finish_scalar(double local, const RegParams& p, F) {
  __shared__ double s_w[kThreads / 32];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if (lane == 0) s_w[warp] = local;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += s_w[w];
    __stcg(p.partials + blockIdx.x, t);
    __threadfence();
    s_last = (atomicAdd(p.counter, 1u) == (unsigned)(p.total_blocks - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double s = 0.0;
  for (int k = tid; k < p.total_blocks; k += kThreads) s += __ldcg(p.partials + k);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) s_w[warp] = s;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += s_w[w];
    *p.out = (float)t;
    *p.counter = 0u;
  }
}

// Smoothness term, one STRIP of kStrip vertically adjacent pixels per thread (neighbouring threads = neighbouring
// columns, so every load of a warp is a contiguous row segment).  The 28 map values the strip's stencils touch
// are loaded once into registers (clamped addresses; validity is handled by the conditions below) and every
// second difference is evaluated once and shared by the pixels it belongs to -- 28 stencil evaluations per strip
// instead of 14 per pixel.  Each difference is formed exactly as the reference forms it: a difference of ROUNDED
// first differences (loss_functions.py:28-33).
#ifndef DVF_SMOOTH_STRIP   // experiment builds
#define DVF_SMOOTH_STRIP 4
#endif
constexpr int kStrip = DVF_SMOOTH_STRIP;

__device__ __forceinline__ float dd(float a, float b, float c) { return sub(sub(c, b), sub(b, a)); }   // (c-b) - (b-a)

// signs / magnitudes of the two cross differences anchored at m00: dxy = (m11-m10)-(m01-m00), dyx = (m11-m01)-(m10-m00)
__device__ __forceinline__ void cross(float m00, float m01, float m10, float m11, float& sgn_sum, float& abs_sum) {
  const float dxy = sub(sub(m11, m10), sub(m01, m00));
  const float dyx = sub(sub(m11, m01), sub(m10, m00));
  sgn_sum = sgnf(dxy) + sgnf(dyx);
  abs_sum = fabsf(dxy) + fabsf(dyx);
}

// kInterior: the strip and its whole 5 x 8 neighbourhood lie inside the map (the vast majority of strips): no index
// clamping, every stencil exists -- the border conditions fold away at compile time and the 28 loads are one row
// pointer each plus constant column offsets.
template <bool kInterior>
__device__ __forceinline__ float smooth_strip(const RegLevel& lv, const float* __restrict__ m, float* __restrict__ g, int y, int x) {
  const int H = lv.H, W = lv.W;
  // C[r]: column x, rows y-2..y+5; L1/R1: columns x-1 / x+1, rows y-1..y+4; L2/R2: columns x-2 / x+2, rows y..y+3
  float C[kStrip + 4], L1[kStrip + 2], R1[kStrip + 2], L2[kStrip], R2[kStrip];
  if (kInterior) {
    const float* row = m + (y - 2) * W + x;
#pragma unroll
    for (int r = 0; r < kStrip + 4; ++r) {
      C[r] = __ldg(row);
      if (r >= 1 && r < kStrip + 3) {
        L1[r - 1] = __ldg(row - 1);
        R1[r - 1] = __ldg(row + 1);
      }
      if (r >= 2 && r < kStrip + 2) {
        L2[r - 2] = __ldg(row - 2);
        R2[r - 2] = __ldg(row + 2);
      }
      row += W;
    }
  } else {
    const int xl1 = max(x - 1, 0), xl2 = max(x - 2, 0), xr1 = min(x + 1, W - 1), xr2 = min(x + 2, W - 1);
#pragma unroll
    for (int r = 0; r < kStrip + 4; ++r) {
      const float* row = m + min(max(y + r - 2, 0), H - 1) * W;
      C[r] = __ldg(row + x);
      if (r >= 1 && r < kStrip + 3) {
        L1[r - 1] = __ldg(row + xl1);
        R1[r - 1] = __ldg(row + xr1);
      }
      if (r >= 2 && r < kStrip + 2) {
        L2[r - 2] = __ldg(row + xl2);
        R2[r - 2] = __ldg(row + xr2);
      }
    }
  }
  const bool xm2 = kInterior || x >= 2, xm1 = kInterior || x >= 1, xp1 = kInterior || x + 1 < W, xp2 = kInterior || x + 2 < W;
  // dyy anchored at rows y-2 .. y+3 (column x)
  float Dy[kStrip + 2];
#pragma unroll
  for (int a = 0; a < kStrip + 2; ++a) {
    const int ay = y + a - 2;
    Dy[a] = (kInterior || (ay >= 0 && ay + 2 < H)) ? dd(C[a], C[a + 1], C[a + 2]) : 0.0f;   // sgn(0) = |0| = 0: an absent stencil adds nothing
  }
  // cross terms anchored at rows y-1 .. y+3, columns x-1 (XL) and x (XR)
  float sL[kStrip + 1], sR[kStrip + 1], aR[kStrip + 1];
#pragma unroll
  for (int a = 0; a < kStrip + 1; ++a) {
    const int ay = y + a - 1;
    const bool rows = kInterior || (ay >= 0 && ay + 1 < H);
    float ab;
    sL[a] = sR[a] = aR[a] = 0.0f;
    if (rows && xm1) cross(L1[a], C[a + 1], L1[a + 1], C[a + 2], sL[a], ab);
    if (rows && xp1) cross(C[a + 1], R1[a], C[a + 2], R1[a + 1], sR[a], aR[a]);
  }
  float lsum = 0.0f;
#pragma unroll
  for (int j = 0; j < kStrip; ++j) {
    if (!kInterior && y + j >= H) break;
    // dxx of row y+j anchored at x-2, x-1, x
    const float c = C[j + 2];
    const float dA = xm2 ? dd(L2[j], L1[j + 1], c) : 0.0f;
    const float dB = (xm1 && xp1) ? dd(L1[j + 1], c, R1[j + 1]) : 0.0f;
    const float dC = xp2 ? dd(c, R1[j + 1], R2[j]) : 0.0f;
    // forward: this pixel owns the stencils anchored at it
    lsum += mul(fabsf(dC), lv.ixx);
    lsum += mul(fabsf(Dy[j + 2]), lv.iyy);
    lsum += mul(aR[j + 1], lv.ixy);
    if (g) {
      // backward (gather form): every stencil that contains the pixel contributes sign * coefficient / count:
      // +1, -2, +1 along a second difference; +1 at a cross stencil's anchor and its diagonal, -1 at the other corners
      float gv = (sgnf(dA) - 2.0f * sgnf(dB) + sgnf(dC)) * lv.ixx;
      gv += (sgnf(Dy[j]) - 2.0f * sgnf(Dy[j + 1]) + sgnf(Dy[j + 2])) * lv.iyy;
      gv += ((sR[j + 1] - sL[j + 1]) - (sR[j] - sL[j])) * lv.ixy;
      g[(y + j) * W + x] = gv * lv.weight;
    }
  }
  return lsum;
}

// Both kernels walk "virtual blocks" of kThreads work items with a grid-stride loop from a grid of a few CTAs per
// SM: one partial and one ticket per CTA instead of one per 256 elements (a single ticket address serialises).
__global__ void __launch_bounds__(kThreads) smooth_loss_kernel(const __grid_constant__ RegParams p) {
  double local = 0.0;
  float lsum = 0.0f;     // this thread's share of the current level (a few dozen terms): folded into `local` per level
  int l = 0;
  for (int vb = blockIdx.x; vb < p.virtual_blocks; vb += gridDim.x) {
    if (l + 1 < p.n_levels && vb >= p.lv[l + 1].block_begin) {
      local += (double)lsum * (double)p.lv[l].weight;
      lsum = 0.0f;
      while (l + 1 < p.n_levels && vb >= p.lv[l + 1].block_begin) ++l;
    }
    const RegLevel& lv = p.lv[l];
    const uint32_t item = (uint32_t)(vb - lv.block_begin) * kThreads + threadIdx.x;   // (image, strip row, column)
    if (item < lv.items) {
      const uint32_t b = fastdiv(item, lv.divSW), rem = item - b * lv.divSW.d_;
      const uint32_t sr = fastdiv(rem, lv.divW), x = rem - sr * lv.divW.d_;
      const size_t off = (size_t)b * lv.H * lv.W;
      const int y0 = (int)sr * kStrip, x0 = (int)x;
      float* const gb = lv.g ? lv.g + off : nullptr;
      if (x0 >= 2 && x0 + 2 < lv.W && y0 >= 2 && y0 + kStrip + 1 < lv.H) lsum += smooth_strip<true>(lv, lv.x + off, gb, y0, x0);
      else lsum += smooth_strip<false>(lv, lv.x + off, gb, y0, x0);
    }
  }
  local += (double)lsum * (double)p.lv[l].weight;
  finish_scalar(local, p, 0);
}

// explainability term: a virtual block is kExplPer * kThreads consecutive elements, every thread keeps kExplPer
// loads in flight (the kernel is a pure stream: 4 B in, 4 B out per element)
constexpr int kExplPer = 4;

__global__ void __launch_bounds__(kThreads) explainability_loss_kernel(const __grid_constant__ RegParams p) {
  double local = 0.0;
  float lsum = 0.0f;     // sum of max(log x, -100) over this thread's elements of the current level
  int l = 0;
  for (int vb = blockIdx.x; vb < p.virtual_blocks; vb += gridDim.x) {
    if (l + 1 < p.n_levels && vb >= p.lv[l + 1].block_begin) {
      local -= (double)lsum * (double)p.lv[l].weight / (double)((long long)p.lv[l].B * p.lv[l].H * p.lv[l].W);
      lsum = 0.0f;
      while (l + 1 < p.n_levels && vb >= p.lv[l + 1].block_begin) ++l;
    }
    const RegLevel& lv = p.lv[l];
    const long long n = (long long)lv.B * lv.H * lv.W;
    const long long base = (long long)(vb - lv.block_begin) * (kThreads * kExplPer) + threadIdx.x;
    const float fn = (float)n, rn = lv.ixx;    // ixx doubles as the correctly rounded 1/n for this kernel
    float xv[kExplPer];
#pragma unroll
    for (int j = 0; j < kExplPer; ++j) {
      const long long idx = base + (long long)j * kThreads;
      xv[j] = idx < n ? ld_stream(lv.x + idx) : 1.0f;
    }
#pragma unroll
    for (int j = 0; j < kExplPer; ++j) {
      const long long idx = base + (long long)j * kThreads;
      if (idx >= n) break;
      const float x = xv[j];
      // F.binary_cross_entropy(x, 1): -max(log x, -100), mean over all elements
      const float lg = logf(x);
      lsum += (lg < -100.0f) ? -100.0f : lg;          // clamp_min(-100) lets NaN through (fmaxf would drop it)
      if (lv.g) {   // torch: (x - 1) / max((1 - x) * x, 1e-12) / n -- two IEEE divisions
        const float num = sub(x, 1.0f);
        const float den = fmaxf(mul(sub(1.0f, x), x), 1e-12f);        // in [1e-12, 0.25] whatever x is
        float q;
        if (fabsf(num) <= 1e20f) {      // every intermediate of the reciprocal-based sequence stays normal: exact
          q = div_by(div_by(num, den, rcp_refined(den)), fn, rn);   // rn = 1.0f / fn: the correctly rounded reciprocal (dvf_math.cuh)
        } else {                        // absurd inputs, NaN: the plain operator
          q = div(div(num, den), fn);
        }
        st_stream(lv.g + idx, q * lv.weight);
      }
    }
  }
  local -= (double)lsum * (double)p.lv[l].weight / (double)((long long)p.lv[l].B * p.lv[l].H * p.lv[l].W);
  finish_scalar(local, p, 0);
}

static int fill(const dvf_reg_level* levels, int n_levels, bool smooth, RegParams& p) {
  if (!levels) return DVF_EINVAL_NULL;
  if (n_levels <= 0 || n_levels > kRegMaxLevels) return DVF_EINVAL_SHAPE;
  int begin = 0;
  for (int l = 0; l < n_levels; ++l) {
    const dvf_reg_level& s = levels[l];
    if (!s.x) return DVF_EINVAL_NULL;
    if (s.B <= 0 || s.H <= 0 || s.W <= 0) return DVF_EINVAL_SHAPE;
    const long long n = (long long)s.B * s.H * s.W;
    if (n >= (1ll << 40)) return DVF_EINVAL_SHAPE;
    // element counts of the four terms (mean denominators), loss_functions.py:35-39; unused when a count is 0
    const float nxx = (float)((double)s.B * s.H * (s.W > 2 ? s.W - 2 : 0));
    const float nyy = (float)((double)s.B * (s.H > 2 ? s.H - 2 : 0) * s.W);
    const float nxy = (float)((double)s.B * (s.H > 1 ? s.H - 1 : 0) * (s.W > 1 ? s.W - 1 : 0));
    const int strips = (s.H + kStrip - 1) / kStrip;
    const long long items = (long long)s.B * strips * s.W;
    if (smooth && items >= (1ll << 31)) return DVF_EINVAL_SHAPE;
    RegLevel& d = p.lv[l];
    d.x = s.x; d.g = s.g; d.B = s.B; d.H = s.H; d.W = s.W; d.weight = s.weight;
    d.ixx = nxx > 0 ? 1.0f / nxx : 0.0f;
    d.iyy = nyy > 0 ? 1.0f / nyy : 0.0f;
    d.ixy = nxy > 0 ? 1.0f / nxy : 0.0f;
    if (!smooth) d.ixx = 1.0f / (float)n;     // explainability: 1/n for the gradient's second division
    d.items = (uint32_t)(smooth ? items : 0);
    d.divSW = make_fastdiv((uint32_t)((long long)strips * s.W));
    d.divW = make_fastdiv((uint32_t)s.W);
    d.block_begin = begin;
    // smooth: one strip of kStrip pixels per thread; explainability: kExplPer elements per thread
    begin += smooth ? (int)((items + kThreads - 1) / kThreads) : (int)((n + kThreads * kExplPer - 1) / (kThreads * kExplPer));
  }
  p.n_levels = n_levels;
  p.virtual_blocks = begin;
  const int cap = num_sms() * 8;     // 8 CTAs of 256 threads per SM: one resident wave
  p.total_blocks = begin < cap ? begin : cap;
  return DVF_OK;
}

}  // namespace dvf

using namespace dvf;

DVF_EXPORT size_t dvf_reg_workspace_bytes(const dvf_reg_level* levels, int32_t n_levels) {
  RegParams p;
  if (fill(levels, n_levels, false, p) != DVF_OK) return 0;
  return 256 + (size_t)num_sms() * 8 * sizeof(double);   // ticket + one partial per CTA; never more than 8 CTAs per SM
}

static int run_reg(bool smooth, const dvf_reg_level* levels, int32_t n_levels, float* out, void* workspace,
                   size_t workspace_bytes, void* stream) {
  RegParams p;
  int st = fill(levels, n_levels, smooth, p);
  if (st != DVF_OK) return st;
  if (!out) return DVF_EINVAL_NULL;
  if (!workspace || workspace_bytes < dvf_reg_workspace_bytes(levels, n_levels)) return DVF_EWORKSPACE;
  if (!aligned(workspace, 256)) return DVF_EINVAL_ALIGN;
  p.counter = static_cast<unsigned*>(workspace);
  p.partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
  p.out = out;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  if (smooth) smooth_loss_kernel<<<p.total_blocks, kThreads, 0, cs>>>(p);
  else explainability_loss_kernel<<<p.total_blocks, kThreads, 0, cs>>>(p);
  return launch_status();
}

DVF_EXPORT int dvf_smooth_loss(const dvf_reg_level* levels, int32_t n_levels, float* out, void* workspace,
                               size_t workspace_bytes, void* stream) {
  return run_reg(true, levels, n_levels, out, workspace, workspace_bytes, stream);
}
DVF_EXPORT int dvf_explainability_loss(const dvf_reg_level* levels, int32_t n_levels, float* out, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  return run_reg(false, levels, n_levels, out, workspace, workspace_bytes, stream);
}
