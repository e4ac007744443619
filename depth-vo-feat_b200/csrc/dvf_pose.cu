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
  float pm[12];
  posemat_fwd(vec + (size_t)n * 6, rotation, pm);
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
    mm3<4>(Ks, pm, Pl);
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
  if (rotation != DVF_ROT_EULER && rotation != DVF_ROT_QUAT) return DVF_EINVAL_DTYPE;
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
