// dvf_caffe.cu -- the Caffe-era formulation of the same hot path: GeoTransform, PinHole,
// InverseWarping and AbsLoss of caffe/src/caffe/layers/{geometry_transformation,pin_hole_layer,
// inverse_warping_layer,abs_loss_layer}.cu, which pytorch_version/geo_transform.py transliterates (non-functional
// there) and unsupervise_dvo.py:95-122 calls.  Conventions differ from the PyTorch path: PIXEL-space sample
// positions (no normalisation, no validity mask), K = (fx,fy,cx,cy), a 4x4 SE(3) matrix, Z + 1e-12 instead of a
// clamp, and sign(0) = -1 in the L1 gradient.  Per-element expressions follow the reference kernels (including the
// promotion of `Z + 1e-12` to double).  What changes is the reduction strategy: the reference issues 12 + 4 + 4
// global atomicAdds PER PIXEL onto 20 addresses per image; here the sums are folded per warp (shuffles) and per
// CTA (shared memory) and one atomic per CTA and entry reaches memory.
#include "dvf_internal.h"
#include "dvf_math.cuh"
#include "dvf_reduce.cuh"

namespace dvf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fold kN per-thread values over the CTA and add them to dst[0..kN) with one atomic per entry
template <int kN>
__device__ __forceinline__ void cta_atomic_add(const float (&v)[kN], float* dst) {
  __shared__ float s[kThreads / 32][kN];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kN; ++k) {
    const float r = warp_sum(v[k]);
    if (lane == 0) s[warp][k] = r;
  }
  __syncthreads();
  if (threadIdx.x < kN) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += s[w][threadIdx.x];
    atomicAdd(dst + threadIdx.x, t);
  }
}

// grid = (ceil(HW / kThreads), N): a CTA never straddles images
__global__ void __launch_bounds__(kThreads) caffe_geo_fwd_kernel(const float* __restrict__ depth, const float* __restrict__ T,
                                                                const float* __restrict__ K, int H, int W,
                                                                float* __restrict__ pts) {
  const int n = blockIdx.y, HW = H * W, idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= HW) return;
  const int y = idx / W, x = idx - y * W;
  const float* t = T + n * 16;
  const float fx = K[n * 4], fy = K[n * 4 + 1], cx = K[n * 4 + 2], cy = K[n * 4 + 3];
  const float d = depth[(size_t)n * HW + idx];
  const float X = mul(div(sub((float)x, cx), fx), d), Y = mul(div(sub((float)y, cy), fy), d);   // (x-cx)/fx*d
#pragma unroll
  for (int r = 0; r < 3; ++r)   // t0*X + t1*Y + t2*d + t3, left to right (nvcc would contract; the oracle does not)
    pts[((size_t)n * 3 + r) * HW + idx] = add(add(add(mul(t[r * 4], X), mul(t[r * 4 + 1], Y)), mul(t[r * 4 + 2], d)), t[r * 4 + 3]);
}

constexpr int kGeoBwdPx = 8;   // pixels per thread of the GeoTransform backward: its 16 sums are folded once per CTA

__global__ void __launch_bounds__(kThreads) caffe_geo_bwd_kernel(const float* __restrict__ top, const float* __restrict__ depth,
                                                                const float* __restrict__ T, const float* __restrict__ K, int H,
                                                                int W, float* __restrict__ depth_diff, float* __restrict__ T_diff,
                                                                float* __restrict__ K_diff) {
  const int n = blockIdx.y, HW = H * W;
  const float* t = T + n * 16;
  const float fx = K[n * 4], fy = K[n * 4 + 1], cx = K[n * 4 + 2], cy = K[n * 4 + 3];
  float acc[kRedSlots];   // 0..11: dT rows 0-2, 12..15: d(fx, fy, cx, cy)
#pragma unroll
  for (int k = 0; k < kRedSlots; ++k) acc[k] = 0.0f;
#pragma unroll 2
  for (int j = 0; j < kGeoBwdPx; ++j) {
    const int idx = (blockIdx.x * kGeoBwdPx + j) * kThreads + threadIdx.x;
    if (idx >= HW) break;
    const int y = idx / W, x = idx - y * W;
    const float g[3] = {ld_stream(top + ((size_t)n * 3 + 0) * HW + idx), ld_stream(top + ((size_t)n * 3 + 1) * HW + idx),
                        ld_stream(top + ((size_t)n * 3 + 2) * HW + idx)};
    const float bX = div(sub((float)x, cx), fx), bY = div(sub((float)y, cy), fy), d = ld_stream(depth + (size_t)n * HW + idx);
    float dd = 0.0f;
#pragma unroll
    for (int r = 0; r < 3; ++r) dd = add(dd, mul(g[r], add(add(mul(t[r * 4], bX), mul(t[r * 4 + 1], bY)), t[r * 4 + 2])));
    if (depth_diff) st_stream(depth_diff + (size_t)n * HW + idx, dd);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      acc[r * 4 + 0] += mul(mul(g[r], bX), d);
      acc[r * 4 + 1] += mul(mul(g[r], bY), d);
      acc[r * 4 + 2] += mul(g[r], d);
      acc[r * 4 + 3] += g[r];
    }
    const float sx = add(add(mul(g[0], t[0]), mul(g[1], t[4])), mul(g[2], t[8]));
    const float sy = add(add(mul(g[0], t[1]), mul(g[1], t[5])), mul(g[2], t[9]));
    acc[12] += mul(sx, mul(div(-bX, fx), d));    // d/dfx
    acc[13] += mul(sy, mul(div(-bY, fy), d));    // d/dfy
    acc[14] += mul(sx, div(-d, fx));             // d/dcx
    acc[15] += mul(sy, div(-d, fy));             // d/dcy
  }
  __shared__ float s[kThreads / 32][kRedSlots];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float r = butterfly16(acc, lane);          // 15 shuffles fold all 16 sums over the warp
  if ((lane & 1) == 0) s[warp][butterfly_slot(lane)] = r;
  __syncthreads();
  if (threadIdx.x < 16) {
    float v = 0.0f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) v += s[w][threadIdx.x];
    if (threadIdx.x < 12) { if (T_diff) atomicAdd(T_diff + n * 16 + threadIdx.x, v); }
    else if (K_diff) atomicAdd(K_diff + n * 4 + (threadIdx.x - 12), v);
  }
}

__global__ void __launch_bounds__(kThreads) caffe_pinhole_fwd_kernel(const float* __restrict__ pts, const float* __restrict__ K,
                                                                    int HW, float* __restrict__ coords) {
  const int n = blockIdx.y, idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= HW) return;
  const float fx = K[n * 4], fy = K[n * 4 + 1], cx = K[n * 4 + 2], cy = K[n * 4 + 3];
  const float X = pts[((size_t)n * 3 + 0) * HW + idx], Y = pts[((size_t)n * 3 + 1) * HW + idx], Z = pts[((size_t)n * 3 + 2) * HW + idx];
  const double zz = (double)Z + 1e-12;   // the reference's `Z+1e-12` is a double expression
  coords[((size_t)n * 2 + 0) * HW + idx] = (float)((double)mul(fx, X) / zz + (double)cx);
  coords[((size_t)n * 2 + 1) * HW + idx] = (float)((double)mul(fy, Y) / zz + (double)cy);
}

__global__ void __launch_bounds__(kThreads) caffe_pinhole_bwd_kernel(const float* __restrict__ cdiff, const float* __restrict__ pts,
                                                                    const float* __restrict__ K, int HW,
                                                                    float* __restrict__ pts_diff, float* __restrict__ K_diff) {
  const int n = blockIdx.y, idx = blockIdx.x * kThreads + threadIdx.x;
  const float fx = K[n * 4], fy = K[n * 4 + 1];
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (idx < HW) {
    const size_t oX = ((size_t)n * 3 + 0) * HW + idx, oY = ((size_t)n * 3 + 1) * HW + idx, oZ = ((size_t)n * 3 + 2) * HW + idx;
    const float gx = cdiff[((size_t)n * 2 + 0) * HW + idx], gy = cdiff[((size_t)n * 2 + 1) * HW + idx];
    const float X = pts[oX], Y = pts[oY], Z = pts[oZ];
    const double zz = (double)Z + 1e-12, z2 = (double)mul(Z, Z) + 1e-12;
    if (pts_diff) {
      pts_diff[oX] = (float)((double)mul(gx, fx) / zz);
      pts_diff[oY] = (float)((double)mul(gy, fy) / zz);
      pts_diff[oZ] = add((float)((double)mul(mul(-gx, fx), X) / z2), (float)((double)mul(mul(-gy, fy), Y) / z2));
    }
    acc[0] = (float)((double)mul(gx, X) / zz);
    acc[1] = (float)((double)mul(gy, Y) / zz);
    acc[2] = gx;
    acc[3] = gy;
  }
  if (K_diff) cta_atomic_add<4>(acc, K_diff + n * 4);
}

// kC > 0: channel count known at compile time (images: 3) -> the channel loop unrolls and all gathers of a pixel
// are in flight together; kC == 0: any channel count
template <int kC>
__global__ void __launch_bounds__(kThreads) caffe_warp_fwd_kernel(const float* __restrict__ U, const float* __restrict__ xy, int Crt, int H,
                                                                 int W, float* __restrict__ out) {
  const int C = kC > 0 ? kC : Crt;
  const int n = blockIdx.y, HW = H * W, idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= HW) return;
  const float xx = ld_stream(xy + ((size_t)n * 2 + 0) * HW + idx), yy = ld_stream(xy + ((size_t)n * 2 + 1) * HW + idx);
  const float fx1 = floorf(xx), fy1 = floorf(yy);
  const int x1 = __float2int_rz(fx1), x2 = x1 + 1, y1 = __float2int_rz(fy1), y2 = y1 + 1;
  const float wx2 = sub(xx, (float)x1), wx1 = sub((float)x2, xx), wy2 = sub(yy, (float)y1), wy1 = sub((float)y2, yy);
  const bool bx1 = (unsigned)x1 < (unsigned)W, bx2 = (unsigned)x2 < (unsigned)W, by1 = (unsigned)y1 < (unsigned)H, by2 = (unsigned)y2 < (unsigned)H;
  const bool p11 = bx1 && by1, p12 = bx1 && by2, p21 = bx2 && by1, p22 = bx2 && by2;
  const float w11 = mul(wx1, wy1), w12 = mul(wx1, wy2), w21 = mul(wx2, wy1), w22 = mul(wx2, wy2);
  const float* pl = U + (size_t)n * C * HW + x1 + y1 * W;     // only dereferenced under the predicates
  float* o = out + (size_t)n * C * HW + idx;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float u11 = p11 ? __ldg(pl) : 0.0f, u12 = p12 ? __ldg(pl + W) : 0.0f;
    const float u21 = p21 ? __ldg(pl + 1) : 0.0f, u22 = p22 ? __ldg(pl + W + 1) : 0.0f;
    float v = 0.0f;   // += w_x * w_y * U, in the reference's tap order
    v = p11 ? add(v, mul(w11, u11)) : v;
    v = p12 ? add(v, mul(w12, u12)) : v;
    v = p21 ? add(v, mul(w21, u21)) : v;
    v = p22 ? add(v, mul(w22, u22)) : v;
    st_stream(o, v);
    pl += HW;
    o += HW;
  }
}

template <int kC, bool kDiffU>
__global__ void __launch_bounds__(kThreads) caffe_warp_bwd_kernel(const float* __restrict__ top, const float* __restrict__ U,
                                                                 const float* __restrict__ xy, int Crt, int H, int W,
                                                                 float* __restrict__ U_diff, float* __restrict__ xy_diff) {
  const int C = kC > 0 ? kC : Crt;
  const int n = blockIdx.y, HW = H * W, idx = blockIdx.x * kThreads + threadIdx.x;
  if (idx >= HW) return;
  const size_t ox = ((size_t)n * 2 + 0) * HW + idx, oy = ((size_t)n * 2 + 1) * HW + idx;
  const float xx = ld_stream(xy + ox), yy = ld_stream(xy + oy);
  const int x1 = __float2int_rz(floorf(xx)), x2 = x1 + 1, y1 = __float2int_rz(floorf(yy)), y2 = y1 + 1;
  const float wx2 = sub(xx, (float)x1), wx1 = sub((float)x2, xx), wy2 = sub(yy, (float)y1), wy1 = sub((float)y2, yy);
  const bool bx1 = (unsigned)x1 < (unsigned)W, bx2 = (unsigned)x2 < (unsigned)W, by1 = (unsigned)y1 < (unsigned)H, by2 = (unsigned)y2 < (unsigned)H;
  const bool p11 = bx1 && by1, p12 = bx1 && by2, p21 = bx2 && by1, p22 = bx2 && by2;
  float tl = 0.0f, tr = 0.0f, bl = 0.0f, br = 0.0f;
  const size_t tap = (size_t)n * C * HW + x1 + y1 * W;
  const float* pl = U + tap;
  const float* gt = top + (size_t)n * C * HW + idx;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float g = ld_stream(gt);
    const float u11 = p11 ? __ldg(pl) : 0.0f, u12 = p12 ? __ldg(pl + W) : 0.0f;
    const float u21 = p21 ? __ldg(pl + 1) : 0.0f, u22 = p22 ? __ldg(pl + W + 1) : 0.0f;
    tl = p11 ? add(tl, mul(g, u11)) : tl;
    bl = p12 ? add(bl, mul(g, u12)) : bl;
    tr = p21 ? add(tr, mul(g, u21)) : tr;
    br = p22 ? add(br, mul(g, u22)) : br;
    if (kDiffU) {
      float* gu = U_diff + tap + (size_t)c * HW;
      if (p11) atomicAdd(gu, mul(mul(g, wx1), wy1));
      if (p12) atomicAdd(gu + W, mul(mul(g, wx1), wy2));
      if (p21) atomicAdd(gu + 1, mul(mul(g, wx2), wy1));
      if (p22) atomicAdd(gu + W + 1, mul(mul(g, wx2), wy2));
    }
    pl += HW;
    gt += HW;
  }
  if (xy_diff) {
    st_stream(xy_diff + ox, add(mul(sub(tr, tl), wy1), mul(sub(br, bl), wy2)));
    st_stream(xy_diff + oy, add(mul(sub(bl, tl), wx1), mul(sub(br, tr), wx2)));
  }
}

__global__ void __launch_bounds__(kThreads) caffe_abs_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t count,
                                                                 float alpha, float* __restrict__ ga, float* __restrict__ gb,
                                                                 double* __restrict__ acc) {
  double local = 0.0;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < count; i += (size_t)gridDim.x * kThreads) {
    const float d = sub(a[i], b[i]);
    local += fabs((double)d);
    const float sg = (d > 0.0f) ? 1.0f : -1.0f;   // (d > 0) - (d <= 0): -1 at d == 0 (and NaN maps to -1... as in the reference: 0 - 0 = 0 for NaN)
    const float s2 = (d != d) ? 0.0f : sg;
    if (ga) ga[i] = mul(alpha, s2);
    if (gb) gb[i] = mul(-alpha, s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double s[kThreads / 32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += s[w];
    atomicAdd(acc, t);
  }
}
__global__ void caffe_abs_loss_finish(const double* acc, int num, float* out) { *out = (float)(*acc / (double)num); }


// ---- edge-aware smoothness of the Caffe graphs (experiments/depth/train.prototxt:4022-4234) ------------------------------
// The prototxt spells it with ten layers: 3x3 "EdgeX" / "EdgeY" convolutions without padding (caffe/include/caffe/filler.hpp:
// 266-315: EdgeX = half the difference between the row below and the row above, EdgeY = half the difference between the
// right and the left column), AbsVal, a 1x1 convolution with weights -0.33 over the three colour channels, Exp, the same
// two edge convolutions on the inverse depth, Eltwise PROD, and AbsLoss against zeros with loss_weight 10:
//   gx(i,j) = exp(-0.33 * sum_c |EdgeX(I_c)(i,j)|),   dx = gx * EdgeX(D),   loss_x = sum |dx| / N      (same with y)
// over the (H-2) x (W-2) window origins.  One launch computes both loss terms and d(w*loss_x + w*loss_y)/dD; the image is
// data (lr_mult 0 convolutions on an input blob).  Gather form of the gradient: D(r,c) is the lower operand of window
// (r-2,c-1), the upper one of (r,c-1), the right one of (r-1,c-2) and the left one of (r-1,c).
struct EdgeWin {
  float gx, gy, dx, dy;
};
__device__ __forceinline__ EdgeWin edge_window(const float* __restrict__ img, const float* __restrict__ D, int HW, int W, int i, int j) {
  // window origin (i, j): rows i..i+2, columns j..j+2
  const int up = i * W + j + 1, dn = (i + 2) * W + j + 1, lf = (i + 1) * W + j, rt = (i + 1) * W + j + 2;
  float sx = 0.0f, sy = 0.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* p = img + (size_t)c * HW;
    sx = add(sx, fabsf(mul(0.5f, sub(__ldg(p + dn), __ldg(p + up)))));
    sy = add(sy, fabsf(mul(0.5f, sub(__ldg(p + rt), __ldg(p + lf)))));
  }
  EdgeWin w;
  w.gx = expf(mul(-0.33f, sx));
  w.gy = expf(mul(-0.33f, sy));
  w.dx = mul(w.gx, mul(0.5f, sub(__ldg(D + dn), __ldg(D + up))));
  w.dy = mul(w.gy, mul(0.5f, sub(__ldg(D + rt), __ldg(D + lf))));
  return w;
}
// d|v|/dv of AbsLoss's second bottom: +1 for v >= 0, -1 for v < 0 (abs_loss_layer.cu:27-29 with diff = 0 - v), 0 for NaN
__device__ __forceinline__ float abs_loss_sign(float v) { return v != v ? 0.0f : (v >= 0.0f ? 1.0f : -1.0f); }

__global__ void __launch_bounds__(kThreads) caffe_edge_smooth_kernel(const float* __restrict__ img, const float* __restrict__ invd, int H,
                                                                    int W, float alpha /* loss_weight / N */, float* __restrict__ ginv,
                                                                    double* __restrict__ acc /*[2]*/) {
  const int n = blockIdx.y, HW = H * W;
  const int idx = blockIdx.x * kThreads + threadIdx.x;
  const float* I = img + (size_t)n * 3 * HW;
  const float* D = invd + (size_t)n * HW;
  double lx = 0.0, ly = 0.0;
  if (idx < HW) {
    const int r = idx / W, c = idx - r * W;
    auto valid = [&](int i, int j) { return i >= 0 && j >= 0 && i < H - 2 && j < W - 2; };
    if (valid(r, c)) {
      const EdgeWin w = edge_window(I, D, HW, W, r, c);
      lx = fabs((double)w.dx);
      ly = fabs((double)w.dy);
    }
    if (ginv) {
      float g = 0.0f;
      if (valid(r - 2, c - 1)) { const EdgeWin w = edge_window(I, D, HW, W, r - 2, c - 1); g = add(g, mul(0.5f, mul(mul(alpha, abs_loss_sign(w.dx)), w.gx))); }
      if (valid(r, c - 1)) { const EdgeWin w = edge_window(I, D, HW, W, r, c - 1); g = sub(g, mul(0.5f, mul(mul(alpha, abs_loss_sign(w.dx)), w.gx))); }
      if (valid(r - 1, c - 2)) { const EdgeWin w = edge_window(I, D, HW, W, r - 1, c - 2); g = add(g, mul(0.5f, mul(mul(alpha, abs_loss_sign(w.dy)), w.gy))); }
      if (valid(r - 1, c)) { const EdgeWin w = edge_window(I, D, HW, W, r - 1, c); g = sub(g, mul(0.5f, mul(mul(alpha, abs_loss_sign(w.dy)), w.gy))); }
      ginv[(size_t)n * HW + idx] = g;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lx += __shfl_xor_sync(0xffffffffu, lx, o);
    ly += __shfl_xor_sync(0xffffffffu, ly, o);
  }
  __shared__ double s[2][kThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    s[0][threadIdx.x >> 5] = lx;
    s[1][threadIdx.x >> 5] = ly;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += s[threadIdx.x][w];
    atomicAdd(acc + threadIdx.x, t);
  }
}
__global__ void caffe_edge_smooth_finish(const double* acc, int num, float* out) {
  out[0] = (float)(acc[0] / (double)num);
  out[1] = (float)(acc[1] / (double)num);
}

static bool bad_nhw(int N, int H, int W) { return N <= 0 || H <= 0 || W <= 0 || N > 65535 || (long long)H * W >= (1ll << 30); }
static dim3 grid_of(int N, int HW) { return dim3((HW + kThreads - 1) / kThreads, N); }

}  // namespace dvf

using namespace dvf;

DVF_EXPORT int dvf_caffe_geo_fwd(const float* depth, const float* T, const float* K, int32_t N, int32_t H, int32_t W, float* pts, void* stream) {
  if (!depth || !T || !K || !pts) return DVF_EINVAL_NULL;
  if (bad_nhw(N, H, W)) return DVF_EINVAL_SHAPE;
  caffe_geo_fwd_kernel<<<grid_of(N, H * W), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(depth, T, K, H, W, pts);
  return launch_status();
}

DVF_EXPORT int dvf_caffe_geo_bwd(const float* top, const float* depth, const float* T, const float* K, int32_t N, int32_t H, int32_t W,
                                 float* depth_diff, float* T_diff, float* K_diff, void* stream) {
  if (!top || !depth || !T || !K) return DVF_EINVAL_NULL;
  if (bad_nhw(N, H, W)) return DVF_EINVAL_SHAPE;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  if (T_diff) cudaMemsetAsync(T_diff, 0, sizeof(float) * 16 * N, cs);
  if (K_diff) cudaMemsetAsync(K_diff, 0, sizeof(float) * 4 * N, cs);
  const dim3 grid((H * W + kThreads * kGeoBwdPx - 1) / (kThreads * kGeoBwdPx), N);
  caffe_geo_bwd_kernel<<<grid, kThreads, 0, cs>>>(top, depth, T, K, H, W, depth_diff, T_diff, K_diff);
  return launch_status();
}

DVF_EXPORT int dvf_caffe_pinhole_fwd(const float* pts, const float* K, int32_t N, int32_t H, int32_t W, float* coords, void* stream) {
  if (!pts || !K || !coords) return DVF_EINVAL_NULL;
  if (bad_nhw(N, H, W)) return DVF_EINVAL_SHAPE;
  caffe_pinhole_fwd_kernel<<<grid_of(N, H * W), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(pts, K, H * W, coords);
  return launch_status();
}

DVF_EXPORT int dvf_caffe_pinhole_bwd(const float* coords_diff, const float* pts, const float* K, int32_t N, int32_t H, int32_t W,
                                     float* pts_diff, float* K_diff, void* stream) {
  if (!coords_diff || !pts || !K) return DVF_EINVAL_NULL;
  if (bad_nhw(N, H, W)) return DVF_EINVAL_SHAPE;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  if (K_diff) cudaMemsetAsync(K_diff, 0, sizeof(float) * 4 * N, cs);
  caffe_pinhole_bwd_kernel<<<grid_of(N, H * W), kThreads, 0, cs>>>(coords_diff, pts, K, H * W, pts_diff, K_diff);
  return launch_status();
}

DVF_EXPORT int dvf_caffe_warp_fwd(const float* img, const float* coords, int32_t N, int32_t C, int32_t H, int32_t W, float* out, void* stream) {
  if (!img || !coords || !out) return DVF_EINVAL_NULL;
  if (bad_nhw(N, H, W) || C <= 0) return DVF_EINVAL_SHAPE;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  if (C == 3) caffe_warp_fwd_kernel<3><<<grid_of(N, H * W), kThreads, 0, cs>>>(img, coords, C, H, W, out);
  else caffe_warp_fwd_kernel<0><<<grid_of(N, H * W), kThreads, 0, cs>>>(img, coords, C, H, W, out);
  return launch_status();
}

DVF_EXPORT int dvf_caffe_warp_bwd(const float* top, const float* img, const float* coords, int32_t N, int32_t C, int32_t H, int32_t W,
                                  float* img_diff, float* coords_diff, void* stream) {
  if (!top || !img || !coords) return DVF_EINVAL_NULL;
  if (bad_nhw(N, H, W) || C <= 0) return DVF_EINVAL_SHAPE;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  if (img_diff) cudaMemsetAsync(img_diff, 0, sizeof(float) * (size_t)N * C * H * W, cs);
  const dim3 grid = grid_of(N, H * W);
  if (C == 3 && img_diff) caffe_warp_bwd_kernel<3, true><<<grid, kThreads, 0, cs>>>(top, img, coords, C, H, W, img_diff, coords_diff);
  else if (C == 3) caffe_warp_bwd_kernel<3, false><<<grid, kThreads, 0, cs>>>(top, img, coords, C, H, W, img_diff, coords_diff);
  else if (img_diff) caffe_warp_bwd_kernel<0, true><<<grid, kThreads, 0, cs>>>(top, img, coords, C, H, W, img_diff, coords_diff);
  else caffe_warp_bwd_kernel<0, false><<<grid, kThreads, 0, cs>>>(top, img, coords, C, H, W, img_diff, coords_diff);
  return launch_status();
}

DVF_EXPORT int dvf_caffe_abs_loss(const float* a, const float* b, uint64_t count, int32_t num, float weight, float* loss, float* ga,
                                  float* gb, void* workspace, void* stream) {
  if (!a || !b || !loss || !workspace) return DVF_EINVAL_NULL;
  if (count == 0 || num <= 0) return DVF_EINVAL_SHAPE;
  if (!aligned(workspace, 8)) return DVF_EINVAL_ALIGN;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  double* acc = static_cast<double*>(workspace);
  cudaMemsetAsync(acc, 0, sizeof(double), cs);
  uint64_t blocks = (count + kThreads - 1) / kThreads;
  const uint64_t cap = (uint64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  caffe_abs_loss_kernel<<<(unsigned)blocks, kThreads, 0, cs>>>(a, b, (size_t)count, weight / (float)num, ga, gb, acc);
  caffe_abs_loss_finish<<<1, 1, 0, cs>>>(acc, num, loss);
  return launch_status();
}

DVF_EXPORT int dvf_caffe_edge_smooth_loss(const float* img, const float* inv_depth, int32_t N, int32_t H, int32_t W, float weight,
                                          float* loss, float* ginv, void* workspace, void* stream) {
  if (!img || !inv_depth || !loss || !workspace) return DVF_EINVAL_NULL;
  if (bad_nhw(N, H, W) || H < 3 || W < 3) return DVF_EINVAL_SHAPE;
  if (!aligned(workspace, 8)) return DVF_EINVAL_ALIGN;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  double* acc = static_cast<double*>(workspace);
  cudaMemsetAsync(acc, 0, 2 * sizeof(double), cs);
  caffe_edge_smooth_kernel<<<grid_of(N, H * W), kThreads, 0, cs>>>(img, inv_depth, H, W, weight / (float)N, ginv, acc);
  caffe_edge_smooth_finish<<<1, 1, 0, cs>>>(acc, N, loss);
  return launch_status();
}
