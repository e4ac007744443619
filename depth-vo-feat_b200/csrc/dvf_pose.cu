// dvf_pose.cu -- 6-DoF vector -> [R|t] -> per-level projection matrices, and the backward.
//
// Replaces pytorch_version/inverse_warp.py: euler2mat :77-114, quat2mat :117-138, pose_vec2mat
// :141-157, `intrinsics @ pose_mat` :188, and the per-scale intrinsics of
// loss_functions_sfm.py:20-21.  The reference spends ~25 tiny launches (sin, cos, stack, bmm, cat)
// per call and twice that in autograd; here one thread per (image, view) does the whole chain in
// registers, forward in fp32 with the reference's operation order (3x3 products are plain
// multiply-adds, left to right, as torch evaluates tiny bmm's), backward analytically in fp64.
#include "dvf_internal.h"
#include "dvf_math.cuh"

namespace dvf {

// out = a(3x3) @ b(3xn), no FMA, (p0 + p1) + p2
template <int N>
__device__ __forceinline__ void mm3(const float* a, const float* b, float* out) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c)
      out[r * N + c] = add(add(mul(a[r * 3 + 0], b[0 * N + c]), mul(a[r * 3 + 1], b[1 * N + c])), mul(a[r * 3 + 2], b[2 * N + c]));
}

__device__ void rotation_fwd(const float* ang, int rotation, float* R) {
  if (rotation == DVF_ROT_EULER) {
    const float x = ang[0], y = ang[1], z = ang[2];
    const float cz = cosf(z), sz = sinf(z), cy = cosf(y), sy = sinf(y), cx = cosf(x), sx = sinf(x);
    const float zero = mul(z, 0.0f);      // inverse_warp.py:93  zeros = z*0
    const float one = add(zero, 1.0f);    // :94
    const float zm[9] = {cz, -sz, zero, sz, cz, zero, zero, zero, one};
    const float ym[9] = {cy, zero, sy, zero, one, zero, -sy, zero, cy};
    const float xm[9] = {one, zero, zero, zero, cx, -sx, zero, sx, cx};
    float xy[9];
    mm3<3>(xm, ym, xy);
    mm3<3>(xy, zm, R);                    // :113  xmat @ ymat @ zmat
  } else {
    float q[4] = {add(mul(ang[0], 0.0f), 1.0f), ang[0], ang[1], ang[2]};  // :125
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) ss = add(ss, mul(q[k], q[k]));
    const float nrm = sqrtf(ss);                                          // :126
    const float w = div(q[0], nrm), x = div(q[1], nrm), y = div(q[2], nrm), z = div(q[3], nrm);
    const float w2 = mul(w, w), x2 = mul(x, x), y2 = mul(y, y), z2 = mul(z, z);
    const float wx = mul(w, x), wy = mul(w, y), wz = mul(w, z), xy = mul(x, y), xz = mul(x, z), yz = mul(y, z);
    R[0] = sub(sub(add(w2, x2), y2), z2);
    R[1] = sub(mul(2.0f, xy), mul(2.0f, wz));
    R[2] = add(mul(2.0f, wy), mul(2.0f, xz));
    R[3] = add(mul(2.0f, wz), mul(2.0f, xy));
    R[4] = sub(add(sub(w2, x2), y2), z2);
    R[5] = sub(mul(2.0f, yz), mul(2.0f, wx));
    R[6] = sub(mul(2.0f, xz), mul(2.0f, wy));
    R[7] = add(mul(2.0f, wx), mul(2.0f, yz));
    R[8] = add(sub(sub(w2, x2), y2), z2);
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
  float R[9], pm[12];
  rotation_fwd(vec + (size_t)n * 6 + 3, rotation, R);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) pm[r * 4 + c] = R[r * 3 + c];
    pm[r * 4 + 3] = vec[(size_t)n * 6 + r];
  }
  if (posemat) {
#pragma unroll
    for (int k = 0; k < 12; ++k) posemat[(size_t)n * 12 + k] = pm[k];
  }
  if (!K) return;
  float Kb[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Kb[k] = K[b * 9 + k];
  for (int l = 0; l < sc.n; ++l) {
    const float ds = sc.ds[l];
    float Ks[9], Pl[12];
#pragma unroll
    for (int k = 0; k < 9; ++k) Ks[k] = (k < 6 && ds != 1.0f) ? div(Kb[k], ds) : Kb[k];  // rows 0-1 / downscale
    mm3<4>(Ks, pm, Pl);
    if (P) {
#pragma unroll
      for (int k = 0; k < 12; ++k) P[((size_t)l * B * V + n) * 12 + k] = Pl[k];
    }
    if (Kinv_s && Kinv && (n % V) == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float m = Kinv[b * 9 + k];
        Kinv_s[((size_t)l * B + b) * 9 + k] = ((k % 3) < 2 && ds != 1.0f) ? mul(m, ds) : m;  // cols 0-1 * downscale
      }
    }
  }
}

__device__ __forceinline__ void dmm3(const double* a, const double* b, double* o) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) o[r * 3 + c] = a[r * 3] * b[c] + a[r * 3 + 1] * b[3 + c] + a[r * 3 + 2] * b[6 + c];
}
__device__ __forceinline__ double ddot9(const double* a, const double* b) {
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 9; ++k) s += a[k] * b[k];
  return s;
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
      const float ds = sc.ds[l];
      const float* g = gP + ((size_t)l * B * V + n) * 12;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float kk = K[b * 9 + k * 3 + r];
            const float ks = (k < 2 && ds != 1.0f) ? div(kk, ds) : kk;
            s += (double)ks * (double)g[k * 4 + c];
          }
          gM[r * 4 + c] += s;
        }
    }
  }
  double gR[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) gR[r * 3 + c] = gM[r * 4 + c];
    gvec[(size_t)n * 6 + r] = (float)gM[r * 4 + 3];
  }
  const float* a = vec + (size_t)n * 6 + 3;
  if (rotation == DVF_ROT_EULER) {
    double sx, cx, sy, cy, sz, cz;
    sincos((double)a[0], &sx, &cx);
    sincos((double)a[1], &sy, &cy);
    sincos((double)a[2], &sz, &cz);
    const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
    const double dRx[9] = {0, 0, 0, 0, -sx, -cx, 0, cx, -sx}, dRy[9] = {-sy, 0, cy, 0, 0, 0, -cy, 0, -sy}, dRz[9] = {-sz, -cz, 0, cz, -sz, 0, 0, 0, 0};
    double t1[9], t2[9];
    dmm3(dRx, Ry, t1); dmm3(t1, Rz, t2);
    gvec[(size_t)n * 6 + 3] = (float)ddot9(gR, t2);
    dmm3(Rx, dRy, t1); dmm3(t1, Rz, t2);
    gvec[(size_t)n * 6 + 4] = (float)ddot9(gR, t2);
    dmm3(Rx, Ry, t1); dmm3(t1, dRz, t2);
    gvec[(size_t)n * 6 + 5] = (float)ddot9(gR, t2);
  } else {
    const double q0[4] = {1.0, (double)a[0], (double)a[1], (double)a[2]};
    const double nn = sqrt(q0[0] * q0[0] + q0[1] * q0[1] + q0[2] * q0[2] + q0[3] * q0[3]);
    const double w = q0[0] / nn, x = q0[1] / nn, y = q0[2] / nn, z = q0[3] / nn;
    // d(sum gR*R)/d(unit quaternion), R as in quat2mat (inverse_warp.py:135-137)
    const double gw = 2 * (w * (gR[0] + gR[4] + gR[8]) + x * (gR[7] - gR[5]) + y * (gR[2] - gR[6]) + z * (gR[3] - gR[1]));
    const double gx = 2 * (x * (gR[0] - gR[4] - gR[8]) + w * (gR[7] - gR[5]) + y * (gR[1] + gR[3]) + z * (gR[2] + gR[6]));
    const double gy = 2 * (y * (gR[4] - gR[0] - gR[8]) + w * (gR[2] - gR[6]) + x * (gR[1] + gR[3]) + z * (gR[5] + gR[7]));
    const double gz = 2 * (z * (gR[8] - gR[0] - gR[4]) + w * (gR[3] - gR[1]) + x * (gR[2] + gR[6]) + y * (gR[5] + gR[7]));
    const double dot = gw * w + gx * x + gy * y + gz * z;
    gvec[(size_t)n * 6 + 3] = (float)((gx - x * dot) / nn);
    gvec[(size_t)n * 6 + 4] = (float)((gy - y * dot) / nn);
    gvec[(size_t)n * 6 + 5] = (float)((gz - z * dot) / nn);
  }
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
