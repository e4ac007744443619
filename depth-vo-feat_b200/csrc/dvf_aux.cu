// dvf_aux.cu -- producers either side of the fused loss.
//
// dvf_area_pyramid replaces the per-scale F.interpolate(img, (h,w), mode='area') calls of
// pytorch_version/loss_functions_sfm.py:18-19 (and loss_function_sfm_old.py:13-15).  For the
// integer factors the reference uses, 'area' is adaptive average pooling: the window is summed
// in fp32 in row-major order and divided by its size -- reproduced here so the pyramid is
// bit-identical to torch-CPU's.  One thread owns an 8x8 input tile (two 16 B loads per row,
// sectors shared with its neighbours) and emits the /2, /4 and /8 outputs from registers: the
// full-resolution image is read from HBM once for all levels instead of once per level.
#include "dvf_internal.h"

namespace dvf {

struct PyrParams {
  const float* img;
  float* out[3];
  int n_out, BC, H, W;
};

__global__ void __launch_bounds__(kThreads) area_pyramid8_kernel(const __grid_constant__ PyrParams p) {
  const int tw = p.W >> 3, th = p.H >> 3;
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= (long long)p.BC * th * tw) return;
  const int tx = (int)(t % tw);
  const int ty = (int)((t / tw) % th);
  const int n = (int)(t / ((long long)tw * th));
  const float* src = p.img + ((size_t)n * p.H + (size_t)ty * 8) * p.W + (size_t)tx * 8;
  float v[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * p.W));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * p.W) + 1);
    v[r][0] = a.x; v[r][1] = a.y; v[r][2] = a.z; v[r][3] = a.w;
    v[r][4] = b.x; v[r][5] = b.y; v[r][6] = b.z; v[r][7] = b.w;
  }
#pragma unroll
  for (int lvl = 0; lvl < 3; ++lvl) {
    if (lvl >= p.n_out) break;
    const int f = 2 << lvl;          // 2, 4, 8
    const int per = 8 / f;           // outputs per tile side
    const int oh = p.H / f, ow = p.W / f;
    float* dst = p.out[lvl] + ((size_t)n * oh + (size_t)ty * per) * ow + (size_t)tx * per;
#pragma unroll
    for (int oy = 0; oy < per; ++oy)
#pragma unroll
      for (int ox = 0; ox < per; ++ox) {
        float s = 0.0f;
#pragma unroll
        for (int y = 0; y < f; ++y)
#pragma unroll
          for (int x = 0; x < f; ++x) s = __fadd_rn(s, v[oy * f + y][ox * f + x]);
        dst[(size_t)oy * ow + ox] = __fdiv_rn(s, (float)(f * f));
      }
  }
}

// any integer factor / any size: one thread per output pixel of one level
__global__ void __launch_bounds__(kThreads) area_level_kernel(const float* __restrict__ img, int BC, int H, int W,
                                                              int oh, int ow, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= (long long)BC * oh * ow) return;
  const int ox = (int)(t % ow), oy = (int)((t / ow) % oh), n = (int)(t / ((long long)ow * oh));
  // adaptive pooling window: [floor(o*I/O), ceil((o+1)*I/O))
  const int ys = (int)(((long long)oy * H) / oh), ye = (int)((((long long)oy + 1) * H + oh - 1) / oh);
  const int xs = (int)(((long long)ox * W) / ow), xe = (int)((((long long)ox + 1) * W + ow - 1) / ow);
  float s = 0.0f;
  for (int y = ys; y < ye; ++y)
    for (int x = xs; x < xe; ++x) s = __fadd_rn(s, __ldg(img + ((size_t)n * H + y) * W + x));
  out[t] = __fdiv_rn(s, (float)((ye - ys) * (xe - xs)));
}

// ---- layout change of feature maps: dense NCHW -> channels-last (and back) ---------------------------------------
// The reference hands feature maps over in NCHW (unsupervise.py:104-109); the feature-loss kernel wants a texel's
// channels contiguous.  Per image this is a [R, S] -> [S, R] transpose (R = C, S = H*W forwards): 32x32 tiles through
// shared memory, reads coalesced along S, writes coalesced along R.  4 B read + 4 B written per element (HBM bound).
template <typename T>
__global__ void __launch_bounds__(256) transpose_tiles_kernel(const T* __restrict__ src, T* __restrict__ dst, int R, int S) {
  __shared__ T tile[32][33];
  const size_t img = (size_t)blockIdx.z * R * S;
  const int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int r = r0 + ty + k, sidx = s0 + tx;
    if (r < R && sidx < S) tile[ty + k][tx] = src[img + (size_t)r * S + sidx];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int sidx = s0 + ty + k, r = r0 + tx;
    if (r < R && sidx < S) dst[img + (size_t)sidx * R + r] = tile[tx][ty + k];
  }
}

}  // namespace dvf

using namespace dvf;

// src [B, R, S] -> dst [B, S, R]; elem_bytes 4 (fp32) or 2 (bf16 / fp16)
// fp32 transpose that also multiplies by a device scalar: the gradient maps of the channels-last feature loss go back to
// the reference's dense NCHW layout and take the upstream gradient in the same pass
__global__ void __launch_bounds__(256) transpose_scale_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int S,
                                                              const float* __restrict__ scale) {
  __shared__ float tile[32][33];
  const float sc = scale ? __ldg(scale) : 1.0f;
  const size_t img = (size_t)blockIdx.z * R * S;
  const int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int r = r0 + ty + k, sidx = s0 + tx;
    if (r < R && sidx < S) tile[ty + k][tx] = src[img + (size_t)r * S + sidx];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int sidx = s0 + ty + k, r = r0 + tx;
    if (r < R && sidx < S) dst[img + (size_t)sidx * R + r] = __fmul_rn(tile[tx][ty + k], sc);
  }
}

DVF_EXPORT int dvf_transpose_planes_scaled(const float* src, float* dst, int32_t B, int32_t R, int32_t S, const float* scale,
                                           void* stream) {
  if (!src || !dst) return DVF_EINVAL_NULL;
  if (B <= 0 || R <= 0 || S <= 0 || B > 65535) return DVF_EINVAL_SHAPE;
  const dim3 grid((S + 31) / 32, (R + 31) / 32, B);
  if (grid.y > 65535) return DVF_EINVAL_SHAPE;
  transpose_scale_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, R, S, scale);
  return launch_status();
}

DVF_EXPORT int dvf_transpose_planes(const void* src, void* dst, int32_t B, int32_t R, int32_t S, int32_t elem_bytes, void* stream) {
  if (!src || !dst) return DVF_EINVAL_NULL;
  if (B <= 0 || R <= 0 || S <= 0 || B > 65535 || (R + 31) / 32 > 65535) return DVF_EINVAL_SHAPE;
  if (elem_bytes != 4 && elem_bytes != 2) return DVF_EINVAL_DTYPE;
  if (!aligned(src, (size_t)elem_bytes) || !aligned(dst, (size_t)elem_bytes)) return DVF_EINVAL_ALIGN;
  const dim3 grid((unsigned)((S + 31) / 32), (unsigned)((R + 31) / 32), (unsigned)B);
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  if (elem_bytes == 4) transpose_tiles_kernel<uint32_t><<<grid, 256, 0, cs>>>(static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst), R, S);
  else transpose_tiles_kernel<uint16_t><<<grid, 256, 0, cs>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), R, S);
  return launch_status();
}

DVF_EXPORT int dvf_area_pyramid(const float* img, int32_t BC, int32_t H, int32_t W, int32_t n_out, float* const* outs,
                                void* stream) {
  if (!img || !outs) return DVF_EINVAL_NULL;
  if (BC <= 0 || H <= 0 || W <= 0 || n_out <= 0 || n_out > 3) return DVF_EINVAL_SHAPE;
  for (int l = 0; l < n_out; ++l)
    if (!outs[l]) return DVF_EINVAL_NULL;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  const bool tile8 = (H % 8 == 0) && (W % 8 == 0) && aligned(img, 16);
  if (tile8) {
    PyrParams p;
    p.img = img;
    p.n_out = n_out;
    p.BC = BC;
    p.H = H;
    p.W = W;
    for (int l = 0; l < 3; ++l) p.out[l] = l < n_out ? outs[l] : nullptr;
    const long long tiles = (long long)BC * (H / 8) * (W / 8);
    area_pyramid8_kernel<<<(unsigned)((tiles + kThreads - 1) / kThreads), kThreads, 0, cs>>>(p);
    return launch_status();
  }
  for (int l = 0; l < n_out; ++l) {
    const int f = 2 << l;
    const int oh = H / f, ow = W / f;
    if (oh <= 0 || ow <= 0) return DVF_EINVAL_SHAPE;
    const long long n = (long long)BC * oh * ow;
    area_level_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, cs>>>(img, BC, H, W, oh, ow, outs[l]);
    int st = launch_status();
    if (st != DVF_OK) return st;
  }
  return DVF_OK;
}

DVF_EXPORT int dvf_area_downsample(const float* img, int32_t BC, int32_t H, int32_t W, int32_t h, int32_t w, float* out,
                                   void* stream) {
  if (!img || !out) return DVF_EINVAL_NULL;
  if (BC <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0) return DVF_EINVAL_SHAPE;
  const long long n = (long long)BC * h * w;
  area_level_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      img, BC, H, W, h, w, out);
  return launch_status();
}
