// dvf_pose.cuh -- device routines shared by the stand-alone pose kernels (dvf_pose.cu) and the fused loss
// kernel, which evaluates them in its prologue / epilogue: 6-DoF vector -> [R|t] (inverse_warp.py:77-157),
// P = K_s @ [R|t] (:188, loss_functions_sfm.py:20) and the analytic fp64 backward.
#pragma once
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

// ---- torch-CPU's fp32 sin / cos, bit for bit -------------------------------------------------------------------------
// inverse_warp.py:89-106 takes torch.cos / torch.sin of the Euler angles.  On the CPU torch evaluates them with MKL's
// VML (vmsSin / vmsCos, high-accuracy mode), whose FMA code path is: n = round(|x| / pi) by the 1.5 * 2^23 trick in
// fp32, r = |x| - n * pi in fp64 (pi in two parts), an odd degree-9 polynomial in fp64, ONE rounding to fp32, sign from
// the parity of n.  fp64 FMAs are IEEE-exact on the GPU as well, so the same sequence gives the same bits; CUDA's sinf /
// cosf agree with it on only ~93 % of inputs, and a projection matrix that is 1 ulp off moves bilinear cells and validity
// masks.  Valid for |x| <= 10000 (above that MKL switches to a table-driven reduction and this falls back to sinf /
// cosf).  Runs once per (image, view) in a kernel prologue: its cost is nil.  Pinned by tests/golden/trig_f32.npz.
__device__ __forceinline__ float torch_trig_poly(double r) {
  const double r2 = __dmul_rn(r, r);
  double p = 0x1.5dbdf0e4c7deep-19;
  p = __fma_rn(p, r2, -0x1.9f6ffeea73463p-13);
  p = __fma_rn(p, r2, 0x1.110ed3804ca96p-7);
  p = __fma_rn(p, r2, -0x1.55554bc836587p-3);
  p = __dmul_rn(r2, p);
  return __double2float_rn(__fma_rn(p, r, r));
}
__device__ __forceinline__ double torch_trig_reduce(float ax, float n) {
  double r = (double)ax;
  r = __fma_rn(-(double)n, 0x1.921fb5444p+1, r);            // pi, high 34 bits
  return __fma_rn(-(double)n, 0x1.68c234c4c6629p-38, r);    // pi, rest
}
static __device__ __noinline__ float torch_sinf(float x) {
  const float ax = fabsf(x);
  if (!(ax <= 10000.0f)) return sinf(x);
  const float y = __fmaf_rn(ax, 0x1.45f306p-2f, 12582912.0f);
  const float n = __fsub_rn(y, 12582912.0f);
  const uint32_t sign = (__float_as_uint(x) & 0x80000000u) ^ (__float_as_uint(y) << 31);
  return __uint_as_float(__float_as_uint(torch_trig_poly(torch_trig_reduce(ax, n))) ^ sign);
}
static __device__ __noinline__ float torch_cosf(float x) {
  const float ax = fabsf(x);
  if (!(ax <= 10000.0f)) return cosf(x);
  const float y = __fmaf_rn(__fadd_rn(ax, 0x1.921fb6p+0f), 0x1.45f306p-2f, 12582912.0f);
  const float n = __fsub_rn(__fsub_rn(y, 12582912.0f), 0.5f);
  return __uint_as_float(__float_as_uint(torch_trig_poly(torch_trig_reduce(ax, n))) ^ (__float_as_uint(y) << 31));
}

// euler2mat given the six trigonometric values (cz, sz, cy, sy, cx, sx): lets callers evaluate sinf/cosf in
// parallel threads and compose afterwards; identical operation order to rotation_fwd
__device__ __forceinline__ void euler_compose(float z, float cz, float sz, float cy, float sy, float cx, float sx, float* R) {
  const float zero = mul(z, 0.0f);      // inverse_warp.py:93  zeros = z*0
  const float one = add(zero, 1.0f);    // :94
  const float zm[9] = {cz, -sz, zero, sz, cz, zero, zero, zero, one};
  const float ym[9] = {cy, zero, sy, zero, one, zero, -sy, zero, cy};
  const float xm[9] = {one, zero, zero, zero, cx, -sx, zero, sx, cx};
  float xy[9];
  mm3<3>(xm, ym, xy);
  mm3<3>(xy, zm, R);                    // :113  xmat @ ymat @ zmat
}

__device__ __forceinline__ void rotation_fwd(const float* ang, int rotation, float* R) {
  if (rotation == DVF_ROT_EULER) {
    const float x = ang[0], y = ang[1], z = ang[2];
    euler_compose(z, torch_cosf(z), torch_sinf(z), torch_cosf(y), torch_sinf(y), torch_cosf(x), torch_sinf(x), R);
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


// [R|t] of one pose vector (tx,ty,tz,rx,ry,rz)
__device__ __forceinline__ void posemat_fwd(const float* vec, int rotation, float* pm /*3x4*/) {
  float R[9];
  rotation_fwd(vec + 3, rotation, R);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) pm[r * 4 + c] = R[r * 3 + c];
    pm[r * 4 + 3] = vec[r];
  }
}

// rows 0-1 of K divided by the level's downscale (true division, loss_functions_sfm.py:20)
__device__ __forceinline__ void scaled_K(const float* K, float ds, float* Ks) {
#pragma unroll
  for (int k = 0; k < 9; ++k) Ks[k] = (k < 6 && ds != 1.0f) ? div(K[k], ds) : K[k];
}
// columns 0-1 of K^-1 multiplied by the downscale (:21)
__device__ __forceinline__ void scaled_Kinv(const float* Kinv, float ds, float* Ms) {
#pragma unroll
  for (int k = 0; k < 9; ++k) Ms[k] = ((k % 3) < 2 && ds != 1.0f) ? mul(Kinv[k], ds) : Kinv[k];
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


// gM += K_s^T @ gP   (dL/d pose_mat from dL/dP of one level), fp64
__device__ __forceinline__ void accumulate_gM(const float* Ks, const float* g /*3x4*/, double* gM /*3x4*/) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += (double)Ks[k * 3 + r] * (double)g[k * 4 + c];
      gM[r * 4 + c] += s;
    }
}

// d(sum gM * [R|t]) / d vec, fp64: gM is dL/d pose_mat (3x4)
__device__ __forceinline__ void posemat_bwd(const double* gM, const float* vec, int rotation, float* gvec /*6*/) {
  double gR[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) gR[r * 3 + c] = gM[r * 4 + c];
    gvec[r] = (float)gM[r * 4 + 3];
  }
  const float* a = vec + 3;
  if (rotation == DVF_ROT_EULER) {
    double sx, cx, sy, cy, sz, cz;
    sincos((double)a[0], &sx, &cx);
    sincos((double)a[1], &sy, &cy);
    sincos((double)a[2], &sz, &cz);
    const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
    const double dRx[9] = {0, 0, 0, 0, -sx, -cx, 0, cx, -sx}, dRy[9] = {-sy, 0, cy, 0, 0, 0, -cy, 0, -sy}, dRz[9] = {-sz, -cz, 0, cz, -sz, 0, 0, 0, 0};
    double t1[9], t2[9];
    dmm3(dRx, Ry, t1); dmm3(t1, Rz, t2);
    gvec[3] = (float)ddot9(gR, t2);
    dmm3(Rx, dRy, t1); dmm3(t1, Rz, t2);
    gvec[4] = (float)ddot9(gR, t2);
    dmm3(Rx, Ry, t1); dmm3(t1, dRz, t2);
    gvec[5] = (float)ddot9(gR, t2);
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
    gvec[3] = (float)((gx - x * dot) / nn);
    gvec[4] = (float)((gy - y * dot) / nn);
    gvec[5] = (float)((gz - z * dot) / nn);
  }
}

}  // namespace dvf
