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
  int block_begin;
};
struct RegParams {
  int n_levels, total_blocks;
  RegLevel lv[kRegMaxLevels];
  double* partials;   // [total_blocks]
  unsigned* counter;  // [1], zero between launches
  float* out;         // [1] loss value
};

__device__ __forceinline__ float sgnf(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }

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

// second differences exactly as the reference forms them: differences of rounded first differences
__device__ __forceinline__ float dxx_at(const float* m, int W, int y, int x) {   // needs x+2 < W
  const float a = m[y * W + x], b = m[y * W + x + 1], c = m[y * W + x + 2];
  return sub(sub(c, b), sub(b, a));
}
__device__ __forceinline__ float dyy_at(const float* m, int W, int y, int x) {   // needs y+2 < H
  const float a = m[y * W + x], b = m[(y + 1) * W + x], c = m[(y + 2) * W + x];
  return sub(sub(c, b), sub(b, a));
}
__device__ __forceinline__ float dxy_at(const float* m, int W, int y, int x) {   // d/dy of dx; needs x+1 < W, y+1 < H
  return sub(sub(m[(y + 1) * W + x + 1], m[(y + 1) * W + x]), sub(m[y * W + x + 1], m[y * W + x]));
}
__device__ __forceinline__ float dyx_at(const float* m, int W, int y, int x) {   // d/dx of dy
  return sub(sub(m[(y + 1) * W + x + 1], m[y * W + x + 1]), sub(m[(y + 1) * W + x], m[y * W + x]));
}

__global__ void __launch_bounds__(kThreads) smooth_loss_kernel(const __grid_constant__ RegParams p) {
  int l = 0;
  while (l + 1 < p.n_levels && (int)blockIdx.x >= p.lv[l + 1].block_begin) ++l;
  const RegLevel& lv = p.lv[l];
  const int H = lv.H, W = lv.W, HW = H * W;
  const long long idx = (long long)(blockIdx.x - lv.block_begin) * kThreads + threadIdx.x;
  double local = 0.0;
  if (idx < (long long)lv.B * HW) {
    const int b = (int)(idx / HW), r = (int)(idx - (long long)b * HW), y = r / W, x = r - y * W;
    const float* m = lv.x + (size_t)b * HW;
    // element counts of the four terms (mean denominators), loss_functions.py:35-39
    const float nxx = (float)((double)lv.B * H * (W > 2 ? W - 2 : 0));
    const float nyy = (float)((double)lv.B * (H > 2 ? H - 2 : 0) * W);
    const float nxy = (float)((double)lv.B * (H > 1 ? H - 1 : 0) * (W > 1 ? W - 1 : 0));
    // forward: this pixel owns the stencils anchored at it
    float lsum = 0.0f;
    if (x + 2 < W) lsum += fabsf(dxx_at(m, W, y, x)) / nxx;
    if (y + 2 < H) lsum += fabsf(dyy_at(m, W, y, x)) / nyy;
    if (x + 1 < W && y + 1 < H) lsum += (fabsf(dxy_at(m, W, y, x)) + fabsf(dyx_at(m, W, y, x))) / nxy;
    local = (double)lsum * (double)lv.weight;
    // backward (gather form): every stencil that contains (y, x) contributes sign * coefficient / count
    if (lv.g) {
      float g = 0.0f;
      // dxx anchored at x-2, x-1, x with coefficients +1, -2, +1
      if (x >= 2) g += sgnf(dxx_at(m, W, y, x - 2)) / nxx;
      if (x >= 1 && x + 1 < W) g -= 2.0f * sgnf(dxx_at(m, W, y, x - 1)) / nxx;
      if (x + 2 < W) g += sgnf(dxx_at(m, W, y, x)) / nxx;
      if (y >= 2) g += sgnf(dyy_at(m, W, y - 2, x)) / nyy;
      if (y >= 1 && y + 1 < H) g -= 2.0f * sgnf(dyy_at(m, W, y - 1, x)) / nyy;
      if (y + 2 < H) g += sgnf(dyy_at(m, W, y, x)) / nyy;
      // cross terms anchored at (y-1|y, x-1|x): +1 at the anchor and its diagonal, -1 at the other two corners
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int ay = y - dy, ax = x - dx;
          if (ay >= 0 && ax >= 0 && ay + 1 < H && ax + 1 < W) {
            const float s = (sgnf(dxy_at(m, W, ay, ax)) + sgnf(dyx_at(m, W, ay, ax))) / nxy;
            g += (dy == dx) ? s : -s;
          }
        }
      lv.g[(size_t)b * HW + r] = g * lv.weight;
    }
  }
  finish_scalar(local, p, 0);
}

__global__ void __launch_bounds__(kThreads) explainability_loss_kernel(const __grid_constant__ RegParams p) {
  int l = 0;
  while (l + 1 < p.n_levels && (int)blockIdx.x >= p.lv[l + 1].block_begin) ++l;
  const RegLevel& lv = p.lv[l];
  const long long n = (long long)lv.B * lv.H * lv.W;
  const long long idx = (long long)(blockIdx.x - lv.block_begin) * kThreads + threadIdx.x;
  double local = 0.0;
  if (idx < n) {
    const float x = lv.x[idx];
    // F.binary_cross_entropy(x, 1): -max(log x, -100), mean over all elements
    const float lg = fmaxf(logf(x), -100.0f);
    local = -(double)lg / (double)n * (double)lv.weight;
    if (lv.g) {   // torch: (x - 1) / max((1 - x) * x, 1e-12) / n
      const float den = fmaxf(mul(sub(1.0f, x), x), 1e-12f);
      lv.g[idx] = div(div(sub(x, 1.0f), den), (float)n) * lv.weight;
    }
  }
  finish_scalar(local, p, 0);
}

static int fill(const dvf_reg_level* levels, int n_levels, RegParams& p) {
  if (!levels) return DVF_EINVAL_NULL;
  if (n_levels <= 0 || n_levels > kRegMaxLevels) return DVF_EINVAL_SHAPE;
  int begin = 0;
  for (int l = 0; l < n_levels; ++l) {
    const dvf_reg_level& s = levels[l];
    if (!s.x) return DVF_EINVAL_NULL;
    if (s.B <= 0 || s.H <= 0 || s.W <= 0) return DVF_EINVAL_SHAPE;
    const long long n = (long long)s.B * s.H * s.W;
    if (n >= (1ll << 40)) return DVF_EINVAL_SHAPE;
    p.lv[l] = RegLevel{s.x, s.g, s.B, s.H, s.W, s.weight, begin};
    begin += (int)((n + kThreads - 1) / kThreads);
  }
  p.n_levels = n_levels;
  p.total_blocks = begin;
  return DVF_OK;
}

}  // namespace dvf

using namespace dvf;

DVF_EXPORT size_t dvf_reg_workspace_bytes(const dvf_reg_level* levels, int32_t n_levels) {
  RegParams p;
  if (fill(levels, n_levels, p) != DVF_OK) return 0;
  return 256 + (size_t)p.total_blocks * sizeof(double);
}

static int run_reg(bool smooth, const dvf_reg_level* levels, int32_t n_levels, float* out, void* workspace,
                   size_t workspace_bytes, void* stream) {
  RegParams p;
  int st = fill(levels, n_levels, p);
  if (st != DVF_OK) return st;
  if (!out) return DVF_EINVAL_NULL;
  if (!workspace || workspace_bytes < 256 + (size_t)p.total_blocks * sizeof(double)) return DVF_EWORKSPACE;
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
