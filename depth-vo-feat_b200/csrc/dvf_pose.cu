// dvf_pose.cu -- 6-DoF vector -> [R|t] -> per-level projection matrices, and the backward.
//
// Replaces pytorch_version/inverse_warp.py: euler2mat :77-114, quat2mat :117-138, pose_vec2mat
// :141-157, `intrinsics @ pose_mat` :188, and the per-scale intrinsics of
// loss_functions_sfm.py:20-21.  The reference spends ~25 tiny launches (sin, cos, stack, bmm, cat)
// per call and twice that in autograd; here one thread per (image, view) does the whole chain in
// registers, forward in fp32 with the reference's operation order (3x3 products are plain
// multiply-adds, left to right, as torch evaluates tiny bmm's), backward analytically in fp64.
// (The fused loss kernel can also evaluate the same routines itself: dvf_photo_loss_fused_pose.)
#include "dvf_pose.cuh"

namespace dvf {

// ---- REF_CUDA arithmetic profile (DVF_ROT_REF_CUDA or-ed into `rotation`) --------------------
// The reference run with torch-CUDA eager rounds the pose chain differently from torch-CPU (this library's default target):
// torch.sin / torch.cos are libdevice's sinf / cosf, and the tiny batched matmuls (xmat @ ymat @ zmat, intrinsics @ pose_mat:
// cuBLAS) accumulate as an FMA chain, k ascending, instead of multiply / add / add.  Established on the B200 against torch
// 2.11+cu128 (profiles/ref_cuda_probe2.py: 100 % of the entries for batches up to 256 -- a 200 000-matrix [3,3]@[3,4] batch
// takes another cuBLAS kernel -- and 100 % of 800 000 angles).  Everything per pixel already rounds alike on both devices
// (profiles/ref_cuda_probe.py), so the profile lives in this file only: callers that want torch-CUDA's bits compute P with
// it and hand P to the loss / warp entries (the fused-pose entry evaluates the torch-CPU profile).
template <int N>
__device__ __forceinline__ void mm3_fma(const float* a, const float* b, float* out) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c)
      out[r * N + c] = fma_(a[r * 3 + 2], b[2 * N + c], fma_(a[r * 3 + 1], b[1 * N + c], mul(a[r * 3 + 0], b[0 * N + c])));
}
static __device__ __noinline__ void posemat_fwd_ref_cuda(const float* vec, int rotation, float* pm /*3x4*/) {
  float R[9];
  if (rotation == DVF_ROT_EULER) {
    const float x = vec[3], y = vec[4], z = vec[5];
    const float cz = cosf(z), sz = sinf(z), cy = cosf(y), sy = sinf(y), cx = cosf(x), sx = sinf(x);
    const float zero = mul(z, 0.0f), one = add(zero, 1.0f);
    const float zm[9] = {cz, -sz, zero, sz, cz, zero, zero, zero, one};
    const float ym[9] = {cy, zero, sy, zero, one, zero, -sy, zero, cy};
    const float xm[9] = {one, zero, zero, zero, cx, -sx, zero, sx, cx};
    float xy[9];
    mm3_fma<3>(xm, ym, xy);
    mm3_fma<3>(xy, zm, R);
  } else {
    // quat2mat has no matmul and no trigonometry, but torch-CUDA sums the norm in another order: 91 % of the entries agree
    // (profiles/ref_cuda_quat_probe.py) -- not part of the verified profile
    rotation_fwd(vec + 3, rotation, R);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) pm[r * 4 + c] = R[r * 3 + c];
    pm[r * 4 + 3] = vec[r];
  }
}

struct Scales {
  float ds[DVF_MAX_LEVELS];
  int n;
};

__global__ void pose_proj_fwd_kernel(const float* __restrict__ vec, const float* __restrict__ K,
                                     const float* __restrict__ Kinv, int B, int V, int rotation, Scales sc,
                                     float* __restrict__ posemat, float* __restrict__ P, float* __restrict__ Kinv_s) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= B * V) return;
  const int b = n / V;
  const bool ref_cuda = (rotation & DVF_ROT_REF_CUDA) != 0;
  rotation &= 0xff;
  float pm[12];
  if (ref_cuda) posemat_fwd_ref_cuda(vec + (size_t)n * 6, rotation, pm);
  else posemat_fwd(vec + (size_t)n * 6, rotation, pm);
  if (posemat) {
#pragma unroll
    for (int k = 0; k < 12; ++k) posemat[(size_t)n * 12 + k] = pm[k];
  }
  if (!K) return;
  float Kb[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Kb[k] = K[b * 9 + k];
  for (int l = 0; l < sc.n; ++l) {
    float Ks[9], Pl[12];
    scaled_K(Kb, sc.ds[l], Ks);
    if (ref_cuda) mm3_fma<4>(Ks, pm, Pl);
    else mm3<4>(Ks, pm, Pl);
    if (P) {
#pragma unroll
      for (int k = 0; k < 12; ++k) P[((size_t)l * B * V + n) * 12 + k] = Pl[k];
    }
    if (Kinv_s && Kinv && (n % V) == 0) {
      float Ms[9];
      scaled_Kinv(Kinv + b * 9, sc.ds[l], Ms);
#pragma unroll
      for (int k = 0; k < 9; ++k) Kinv_s[((size_t)l * B + b) * 9 + k] = Ms[k];
    }
  }
}

__global__ void pose_proj_bwd_kernel(const float* __restrict__ gP, const float* __restrict__ gposemat,
                                     const float* __restrict__ vec, const float* __restrict__ K, int B, int V,
                                     int rotation, Scales sc, float* __restrict__ gvec) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= B * V) return;
  const int b = n / V;
  double gM[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) gM[k] = gposemat ? (double)gposemat[(size_t)n * 12 + k] : 0.0;
  if (gP && K) {
    for (int l = 0; l < sc.n; ++l) {
      float Ks[9];
      scaled_K(K + b * 9, sc.ds[l], Ks);
      accumulate_gM(Ks, gP + ((size_t)l * B * V + n) * 12, gM);
    }
  }
  float g6[6];
  posemat_bwd(gM, vec + (size_t)n * 6, rotation, g6);
#pragma unroll
  for (int k = 0; k < 6; ++k) gvec[(size_t)n * 6 + k] = g6[k];
}

static int make_scales(const float* downscale, int n_levels, Scales& sc) {
  if (n_levels < 0 || n_levels > DVF_MAX_LEVELS) return DVF_EINVAL_SHAPE;
  if (n_levels > 0 && !downscale) return DVF_EINVAL_NULL;
  sc.n = n_levels;
  for (int l = 0; l < DVF_MAX_LEVELS; ++l) sc.ds[l] = l < n_levels ? downscale[l] : 1.0f;
  return DVF_OK;
}

}  // namespace dvf

using namespace dvf;

DVF_EXPORT int dvf_pose_proj_fwd(const float* vec, const float* K, const float* Kinv, int32_t B, int32_t V,
                                 int32_t rotation, const float* downscale, int32_t n_levels, float* posemat, float* P,
                                 float* Kinv_s, void* stream) {
  if (!vec) return DVF_EINVAL_NULL;
  if (B <= 0 || V <= 0) return DVF_EINVAL_SHAPE;
  if ((rotation & ~DVF_ROT_REF_CUDA) != DVF_ROT_EULER && (rotation & ~DVF_ROT_REF_CUDA) != DVF_ROT_QUAT) return DVF_EINVAL_DTYPE;
  if (!posemat && !P && !Kinv_s) return DVF_EINVAL_NULL;
  if ((P || Kinv_s) && !K) return DVF_EINVAL_NULL;
  if (Kinv_s && !Kinv) return DVF_EINVAL_NULL;
  Scales sc;
  int st = make_scales(downscale, n_levels, sc);
  if (st != DVF_OK) return st;
  const int n = B * V, threads = 128;
  pose_proj_fwd_kernel<<<(n + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      vec, K, Kinv, B, V, rotation, sc, posemat, P, Kinv_s);
  return launch_status();
}

DVF_EXPORT int dvf_pose_proj_bwd(const float* gP, const float* gposemat, const float* vec, const float* K, int32_t B,
                                 int32_t V, int32_t rotation, const float* downscale, int32_t n_levels, float* gvec,
                                 void* stream) {
  if (!vec || !gvec) return DVF_EINVAL_NULL;
  if (!gP && !gposemat) return DVF_EINVAL_NULL;
  if (gP && !K) return DVF_EINVAL_NULL;
  if (B <= 0 || V <= 0) return DVF_EINVAL_SHAPE;
  rotation &= ~DVF_ROT_REF_CUDA;   // the backward is analytic in fp64: one form for both profiles
  if (rotation != DVF_ROT_EULER && rotation != DVF_ROT_QUAT) return DVF_EINVAL_DTYPE;
  Scales sc;
  int st = make_scales(downscale, n_levels, sc);
  if (st != DVF_OK) return st;
  const int n = B * V, threads = 128;
  pose_proj_bwd_kernel<<<(n + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      gP, gposemat, vec, K, B, V, rotation, sc, gvec);
  return launch_status();
}

// ---- torch-CPU's fp32 sin / cos as a stand-alone operator (dvf_pose.cuh: torch_sinf / torch_cosf) -------------------
namespace dvf {
__global__ void torch_trig_kernel(const float* __restrict__ x, long long n, float* __restrict__ s, float* __restrict__ c) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (s) s[i] = torch_sinf(v);
    if (c) c[i] = torch_cosf(v);
  }
}
}  // namespace dvf

DVF_EXPORT int dvf_torch_sincos(const float* x, int64_t n, float* sin_out, float* cos_out, void* stream) {
  if (n < 0) return DVF_EINVAL_SHAPE;
  if (n == 0) return DVF_OK;
  if (!x || (!sin_out && !cos_out)) return DVF_EINVAL_NULL;
  long long blocks = (n + 255) / 256;
  if (blocks > 8ll * dvf::num_sms()) blocks = 8ll * dvf::num_sms();
  dvf::torch_trig_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, sin_out, cos_out);
  return dvf::launch_status();
}
