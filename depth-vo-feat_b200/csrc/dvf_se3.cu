// dvf_se3.cu -- se(3) -> SE(3) exponential map on the device.
//
// Replaces pytorch_version/se3_generate.py:7-103 (SE3_Generator_KITTI; verbatim copies in model.py:33-131,
// fixmodel.py:10-108 and caffe/python/pygeometry.py:6-115): the reference leaves the GPU every iteration, loops over
// the batch in numpy and copies the result back.  Convention (se3_generate.py:13-46): input (w(3), u(3)),
// output [[R, R u], [0, 1]] with R = I + c1 [w]x + c2 [w]x^2 (Rodrigues), R = I + [w]x when |w|^2 < 1e-12.
// Precision follows the reference: theta, c1, c2 in fp32, the matrix algebra in fp64, fp64 output.
// One thread per pose; launch-latency bound.
#include "dvf_internal.h"

namespace dvf {

__device__ __forceinline__ void skew(const double* v, double* m) {
  m[0] = 0; m[1] = -v[2]; m[2] = v[1];
  m[3] = v[2]; m[4] = 0; m[5] = -v[0];
  m[6] = -v[1]; m[7] = v[0]; m[8] = 0;
}

// R and theta of one rotation vector
__device__ void rodrigues(const float* w, double* R, double* wx, float* theta) {
  const double wd[3] = {(double)w[0], (double)w[1], (double)w[2]};
  skew(wd, wx);
  const float th = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(w[0], w[0]), __fmul_rn(w[1], w[1])), __fmul_rn(w[2], w[2])));
  *theta = th;
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
  if (__fmul_rn(th, th) < 1e-12f) {                       // se3_generate.py:33
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] += wx[k];
    return;
  }
  const float c1 = __fdiv_rn(sinf(th), th);               // :37
  const float sh = sinf(__fdiv_rn(th, 2.0f));
  const float c2 = __fdiv_rn(__fmul_rn(2.0f, __fmul_rn(sh, sh)), __fmul_rn(th, th));   // :38
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += wx[r * 3 + k] * wx[k * 3 + c];
      R[r * 3 + c] += (double)c1 * wx[r * 3 + c] + (double)c2 * s;   // :42
    }
}

__global__ void se3_exp_fwd_kernel(const float* __restrict__ in, int B, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* w = in + (size_t)b * 6;
  const float* u = w + 3;
  double R[9], wx[9];
  float th;
  rodrigues(w, R, wx, &th);
  double* o = out + (size_t)b * 16;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) o[r * 4 + c] = R[r * 3 + c];
    o[r * 4 + 3] = R[r * 3] * (double)u[0] + R[r * 3 + 1] * (double)u[1] + R[r * 3 + 2] * (double)u[2];   // :46
  }
  o[12] = o[13] = o[14] = 0.0;
  o[15] = 1.0;
}

__global__ void se3_exp_bwd_kernel(const float* __restrict__ in, const double* __restrict__ gout, int B,
                                   float* __restrict__ gin) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* w = in + (size_t)b * 6;
  const float* u = w + 3;
  const double* g = gout + (size_t)b * 16;
  double R[9], wx[9];
  float th;
  rodrigues(w, R, wx, &th);
  const double dT[3] = {g[3], g[7], g[11]};
#pragma unroll
  for (int c = 0; c < 3; ++c) gin[(size_t)b * 6 + 3 + c] = (float)(dT[0] * R[c] + dT[1] * R[3 + c] + dT[2] * R[6 + c]);   // :68
  double dR[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) dR[r * 3 + c] = g[r * 4 + c] + dT[r] * (double)u[c];   // :71-75
  const bool small = __fmul_rn(th, th) < 1e-12f;
  const double th2 = (double)__fmul_rn(th, th);
  for (int idx = 0; idx < 3; ++idx) {
    double dRdw[9];
    if (small) {
      // the reference's generator matrices (se3_generate.py:79-81), signs as written there
      double e[3] = {0.0, 0.0, 0.0};
      e[idx] = -1.0;
      skew(e, dRdw);
    } else {
      double col[3], ct[3], cr[9];
#pragma unroll
      for (int r = 0; r < 3; ++r) col[r] = ((r == idx) ? 1.0 : 0.0) - R[r * 3 + idx];
#pragma unroll
      for (int r = 0; r < 3; ++r) ct[r] = wx[r * 3] * col[0] + wx[r * 3 + 1] * col[1] + wx[r * 3 + 2] * col[2];   // :87
      skew(ct, cr);                                                                                         // :88-94
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < 3; ++k) s += (((double)w[idx] * wx[r * 3 + k] + cr[r * 3 + k]) / th2) * R[k * 3 + c];   // :99
          dRdw[r * 3 + c] = s;
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) s += dR[k] * dRdw[k];   // :100
    gin[(size_t)b * 6 + idx] = (float)s;
  }
}

}  // namespace dvf

using namespace dvf;

DVF_EXPORT int dvf_se3_exp_fwd(const float* in, int32_t B, double* out, void* stream) {
  if (!in || !out) return DVF_EINVAL_NULL;
  if (B <= 0) return DVF_EINVAL_SHAPE;
  if (!aligned(out, 8)) return DVF_EINVAL_ALIGN;
  se3_exp_fwd_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(in, B, out);
  return launch_status();
}

DVF_EXPORT int dvf_se3_exp_bwd(const float* in, const double* gout, int32_t B, float* gin, void* stream) {
  if (!in || !gout || !gin) return DVF_EINVAL_NULL;
  if (B <= 0) return DVF_EINVAL_SHAPE;
  if (!aligned(gout, 8)) return DVF_EINVAL_ALIGN;
  se3_exp_bwd_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(in, gout, B, gin);
  return launch_status();
}
