/*
 * dvf_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic that the reference's hot path
 * (Depth-VO-Feat, pytorch_version/) executes when it runs on torch-CPU fp32.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 * The product (depth-vo-feat_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit (forward
 * coordinate chain, warped image, masks) or to <=1e-6 (gradients, reductions)
 * against the reference's own Python executed in the build container, and
 * against the golden vectors in tests/golden/ that oracle/gen_golden.py wrote
 * from that same execution (tests/test_oracle_*.py).
 *
 * Every routine cites the reference lines it restates (paths relative to the
 * reference checkout, pytorch_version/...).  Arithmetic order notes ("FMA
 * chain", "true division") were established by experiment against torch 2.11
 * CPU and are what makes the bilinear cell / validity mask reproducible.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -mfma).
 * All tensors are dense row-major fp32, images NCHW, unless stated.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "torch_trig.h"   /* torch-CPU's fp32 sin / cos (MKL VML, HA) restated */

#define DVFO_API __attribute__((visibility("default")))

enum { DVFO_PAD_ZEROS = 0, DVFO_PAD_BORDER = 1 };
enum { DVFO_ROT_EULER = 0, DVFO_ROT_QUAT = 1 };

DVFO_API int dvfo_version(void) { return 2; }

/* torch.sin / torch.cos of n fp32 values (torch_trig.h) */
DVFO_API void dvfo_torch_trig(const float *x, long n, float *s, float *c) {
  for (long i = 0; i < n; ++i) {
    s[i] = dvfo_torch_sinf(x[i]);
    c[i] = dvfo_torch_cosf(x[i]);
  }
}

/* ------------------------------------------------------------------------ */
/* small dense helpers                                                       */
/* ------------------------------------------------------------------------ */

/* torch-CPU bmm for tiny operands ([B,3,3]@[B,3,3], [B,3,3]@[B,3,4]):
 * plain multiply-add, left to right, no FMA (verified bit-exact). */
static void mm3_small(const float *a /*3x3*/, const float *b /*3xn*/, int n,
                      float *out /*3xn*/) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < n; ++c) {
      float p0 = a[r * 3 + 0] * b[0 * n + c];
      float p1 = a[r * 3 + 1] * b[1 * n + c];
      float p2 = a[r * 3 + 2] * b[2 * n + c];
      out[r * n + c] = (p0 + p1) + p2;
    }
}

/* torch-CPU bmm [B,3,3]@[B,3,HW] (inverse_warp.py:39 and :55): one rounding of
 * the first product, then an FMA chain with k ascending (verified bit-exact). */
static inline float dot3_fma(float m0, float m1, float m2, float c0, float c1,
                             float c2) {
  float acc = m0 * c0;
  acc = fmaf(m1, c1, acc);
  acc = fmaf(m2, c2, acc);
  return acc;
}

/* ------------------------------------------------------------------------ */
/* pose_vec2mat / euler2mat / quat2mat  (inverse_warp.py:77-157)             */
/* ------------------------------------------------------------------------ */

/* euler2mat, inverse_warp.py:77-114:  R = (Rx @ Ry) @ Rz. */
static void euler2mat_one(const float *ang, float *R) {
  float x = ang[0], y = ang[1], z = ang[2];
  /* torch.cos / torch.sin, :89-90, :98-99, :105-106 -- see torch_trig.h */
  float cz = dvfo_torch_cosf(z), sz = dvfo_torch_sinf(z);
  float cy = dvfo_torch_cosf(y), sy = dvfo_torch_sinf(y);
  float cx = dvfo_torch_cosf(x), sx = dvfo_torch_sinf(x);
  float zero = z * 0.0f;       /* :93 zeros = z.detach()*0  */
  float one = zero + 1.0f;     /* :94                         */
  float zm[9] = {cz, -sz, zero, sz, cz, zero, zero, zero, one};
  float ym[9] = {cy, zero, sy, zero, one, zero, -sy, zero, cy};
  float xm[9] = {one, zero, zero, zero, cx, -sx, zero, sx, cx};
  float xy[9];
  mm3_small(xm, ym, 3, xy);
  mm3_small(xy, zm, 3, R); /* :113 */
}

/* quat2mat, inverse_warp.py:117-138. */
static void quat2mat_one(const float *q3, float *R) {
  float q[4] = {q3[0] * 0.0f + 1.0f, q3[0], q3[1], q3[2]}; /* :125 */
  /* :126 norm(p=2, dim=1): torch-CPU accumulates the 4 squares in fp32 */
  float ss = 0.0f;
  for (int k = 0; k < 4; ++k) ss += q[k] * q[k];
  float nrm = sqrtf(ss);
  float w = q[0] / nrm, x = q[1] / nrm, y = q[2] / nrm, z = q[3] / nrm;
  float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z; /* :131 */
  float wx = w * x, wy = w * y, wz = w * z;             /* :132 */
  float xy = x * y, xz = x * z, yz = y * z;             /* :133 */
  R[0] = ((w2 + x2) - y2) - z2;
  R[1] = 2.0f * xy - 2.0f * wz;
  R[2] = 2.0f * wy + 2.0f * xz;
  R[3] = 2.0f * wz + 2.0f * xy;
  R[4] = ((w2 - x2) + y2) - z2;
  R[5] = 2.0f * yz - 2.0f * wx;
  R[6] = 2.0f * xz - 2.0f * wy;
  R[7] = 2.0f * wx + 2.0f * yz;
  R[8] = ((w2 - x2) - y2) + z2;
}

/* pose_vec2mat, inverse_warp.py:141-157: vec=(tx,ty,tz,rx,ry,rz) -> [R|t]. */
DVFO_API void dvfo_pose_vec2mat(const float *vec, int n, int rotation_mode,
                                float *out /*[n,3,4]*/) {
  for (int b = 0; b < n; ++b) {
    float R[9];
    if (rotation_mode == DVFO_ROT_QUAT)
      quat2mat_one(vec + b * 6 + 3, R);
    else
      euler2mat_one(vec + b * 6 + 3, R);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) out[b * 12 + r * 4 + c] = R[r * 3 + c];
      out[b * 12 + r * 4 + 3] = vec[b * 6 + r];
    }
  }
}

/* proj_cam_to_src_pixel = intrinsics @ pose_mat, inverse_warp.py:188. */
DVFO_API void dvfo_project(const float *K /*[n,3,3]*/,
                           const float *posemat /*[n,3,4]*/, int n,
                           float *P /*[n,3,4]*/) {
  for (int b = 0; b < n; ++b) mm3_small(K + b * 9, posemat + b * 12, 4, P + b * 12);
}

/* per-scale intrinsics, loss_functions_sfm.py:20-21:
 *   K_s    = cat(K[:,0:2]/downscale, K[:,2:])       (true division)
 *   Kinv_s = cat(Kinv[:,:,0:2]*downscale, Kinv[:,:,2:]) */
DVFO_API void dvfo_scale_intrinsics(const float *K, const float *Kinv, int n,
                                    float downscale, float *Ks, float *Kinvs) {
  for (int b = 0; b < n; ++b)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        int i = b * 9 + r * 3 + c;
        Ks[i] = (r < 2) ? K[i] / downscale : K[i];
        Kinvs[i] = (c < 2) ? Kinv[i] * downscale : Kinv[i];
      }
}

/* ------------------------------------------------------------------------ */
/* per-pixel forward chain                                                   */
/* ------------------------------------------------------------------------ */

typedef struct {
  float ray[3];  /* Kinv @ (j,i,1)              inverse_warp.py:38-39 */
  float cam[3];  /* ray * depth                 :40                   */
  float q[3];    /* P_rot @ cam + P_tr          :55-60                */
  float Z;       /* clamp(q_z, min=1e-3)        :63                   */
  float u, v;    /* X/Z, Y/Z                                          */
  float xn, yn;  /* normalised, after the zeros-padding overwrite :65-71 */
  int mx, my;    /* coordinate was overwritten by 2 (gradient killed) */
} px_chain;

static inline float clamp_min(float v, float lo) {
  /* torch.clamp propagates NaN */
  if (v != v) return v;
  return v < lo ? lo : v;
}

static inline void chain_fwd(const float *P /*3x4*/, const float *M /*3x3*/,
                             float d, int i, int j, int H, int W, int padding,
                             px_chain *o) {
  float fj = (float)j, fi = (float)i;
  for (int k = 0; k < 3; ++k) {
    o->ray[k] = dot3_fma(M[k * 3 + 0], M[k * 3 + 1], M[k * 3 + 2], fj, fi, 1.0f);
    o->cam[k] = o->ray[k] * d;
  }
  for (int k = 0; k < 3; ++k)
    o->q[k] = dot3_fma(P[k * 4 + 0], P[k * 4 + 1], P[k * 4 + 2], o->cam[0],
                       o->cam[1], o->cam[2]) +
              P[k * 4 + 3];
  o->Z = clamp_min(o->q[2], 1e-3f);
  o->u = o->q[0] / o->Z;
  o->v = o->q[1] / o->Z;
  /* 2*(X/Z)/(w-1) - 1 : exact doubling, TRUE division by float(w-1), subtract */
  o->xn = (2.0f * o->u) / (float)(W - 1) - 1.0f;
  o->yn = (2.0f * o->v) / (float)(H - 1) - 1.0f;
  if (padding & 4 /* DVFO_REF_CUDA, defined below */) {
    /* torch-CUDA eager: division by a python scalar = multiplication by its fp32 reciprocal
     * (ATen/native/cuda/BinaryDivTrueKernel.cu: a * (1 / b)) */
    const float rw = 1.0f / (float)(W - 1), rh = 1.0f / (float)(H - 1);
    o->xn = (2.0f * o->u) * rw - 1.0f;
    o->yn = (2.0f * o->v) * rh - 1.0f;
  }
  o->mx = o->my = 0;
  if ((padding & 1) == DVFO_PAD_ZEROS) {
    if (o->xn > 1.0f || o->xn < -1.0f) { o->xn = 2.0f; o->mx = 1; }
    if (o->yn > 1.0f || o->yn < -1.0f) { o->yn = 2.0f; o->my = 1; }
  }
}

/* F.grid_sample(..., mode='bilinear', align_corners=False) as run by torch-CPU:
 * the vectorised kernel un-normalises as fma(x+1, size/2, -0.5)  (verified
 * bit-exact; the generic header's ((x+1)*size-1)/2 is NOT what executes). */
typedef struct {
  float ix, iy;       /* un-normalised (and, for border padding, clipped)  */
  float gmx, gmy;     /* d ix / d x_n  (size/2, or 0 where border-clipped)  */
  int x0, y0;         /* floor                                             */
  float w, e, n, s;   /* w = ix-x0, e = 1-w, n = iy-y0, s = 1-n            */
  int cuda_w;         /* DVFO_REF_CUDA: tap weights formed as ATen's CUDA kernel forms them */
  float cw[4];        /* nw, ne, sw, se                                     */
} samp_loc;

/* padding arguments carry the grid_sample convention in bit 1: DVFO_ALIGN_CORNERS = align_corners=True (the torch <= 1.2
 * behaviour the reference was written for; ATen ComputeLocation: (x+1) * ((size-1)/2), one rounded product) */
#define DVFO_ALIGN_CORNERS 2
/* bit 2: DVFO_REF_CUDA = the per-pixel rounding of the reference run with torch-CUDA eager (forward only; PARITY PINNED by
 * tests/golden/ref_cuda_warp.npz, produced on a B200 by oracle/gen_golden_ref_cuda.py from torch-CUDA itself): scalar
 * division as a reciprocal multiply (chain_fwd) and the bilinear weights of ATen/native/cuda/GridSampler.cu,
 * nw = (x1 - ix) * (y1 - iy), ne = (ix - x0) * (y1 - iy), sw = (x1 - ix) * (iy - y0), se = (ix - x0) * (iy - y0). */
#define DVFO_REF_CUDA 4
static inline float unnormalize(float c, int size, int align_corners) {
  if (align_corners) return (c + 1.0f) * ((float)(size - 1) / 2.0f);
  return fmaf(c + 1.0f, (float)size / 2.0f, -0.5f);
}

static inline void locate(float xn, float yn, int H, int W, int padding,
                          samp_loc *L) {
  const int ac = (padding & DVFO_ALIGN_CORNERS) != 0;
  L->ix = unnormalize(xn, W, ac);
  L->iy = unnormalize(yn, H, ac);
  L->gmx = ac ? (float)(W - 1) / 2.0f : (float)W / 2.0f;
  L->gmy = ac ? (float)(H - 1) / 2.0f : (float)H / 2.0f;
  if ((padding & 1) == DVFO_PAD_BORDER) {
    /* clip_coordinates(_set_grad), ATen/native/GridSampler.h */
    float mxv = (float)(W - 1), myv = (float)(H - 1);
    if (L->ix <= 0.0f) { L->ix = 0.0f; L->gmx = 0.0f; }
    else if (L->ix >= mxv) { L->ix = mxv; L->gmx = 0.0f; }
    if (L->iy <= 0.0f) { L->iy = 0.0f; L->gmy = 0.0f; }
    else if (L->iy >= myv) { L->iy = myv; L->gmy = 0.0f; }
  }
  float fx = floorf(L->ix), fy = floorf(L->iy);
  L->w = L->ix - fx;
  L->e = 1.0f - L->w;
  L->n = L->iy - fy;
  L->s = 1.0f - L->n;
  L->cuda_w = (padding & DVFO_REF_CUDA) != 0;
  if (L->cuda_w) {
    const float x1 = fx + 1.0f, y1 = fy + 1.0f;
    L->cw[0] = (x1 - L->ix) * (y1 - L->iy);
    L->cw[1] = (L->ix - fx) * (y1 - L->iy);
    L->cw[2] = (x1 - L->ix) * (L->iy - fy);
    L->cw[3] = (L->ix - fx) * (L->iy - fy);
  }
  /* NaN / huge coordinates: every tap is out of bounds */
  if (!(fx >= -2.0f && fx <= (float)W + 1.0f)) fx = -2.0f;
  if (!(fy >= -2.0f && fy <= (float)H + 1.0f)) fy = -2.0f;
  L->x0 = (int)fx;
  L->y0 = (int)fy;
}

static inline int inb(int x, int y, int H, int W) {
  return x >= 0 && x < W && y >= 0 && y < H;
}

/* ------------------------------------------------------------------------ */
/* pixel2cam / cam2pixel / grid_sample / inverse_warp forward                */
/* ------------------------------------------------------------------------ */

/* pixel2cam, inverse_warp.py:26-40 -> [B,3,H,W]. */
DVFO_API void dvfo_pixel2cam(const float *depth, const float *Kinv, int B, int H,
                             int W, float *cam) {
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) {
        const float *M = Kinv + b * 9;
        float d = depth[((size_t)b * H + i) * W + j];
        for (int k = 0; k < 3; ++k) {
          float ray = dot3_fma(M[k * 3], M[k * 3 + 1], M[k * 3 + 2], (float)j,
                               (float)i, 1.0f);
          cam[(((size_t)b * 3 + k) * H + i) * W + j] = ray * d;
        }
      }
}

/* cam2pixel, inverse_warp.py:43-74 composed with pixel2cam -> grid [B,H,W,2]. */
DVFO_API void dvfo_grid(const float *depth, const float *P, const float *Kinv,
                        int B, int H, int W, int padding, float *grid) {
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) {
        px_chain c;
        size_t p = ((size_t)b * H + i) * W + j;
        chain_fwd(P + b * 12, Kinv + b * 9, depth[p], i, j, H, W, padding, &c);
        grid[p * 2 + 0] = c.xn;
        grid[p * 2 + 1] = c.yn;
      }
}

/* one output pixel of grid_sample, all channels; returns 1 if any channel != 0 */
static inline int sample_px(const float *img_b /*[C,H,W]*/, int C, int H, int W,
                            const samp_loc *L, float *out /*stride HW*/,
                            size_t out_stride) {
  int x0 = L->x0, y0 = L->y0, x1 = x0 + 1, y1 = y0 + 1;
  int bnw = inb(x0, y0, H, W), bne = inb(x1, y0, H, W);
  int bsw = inb(x0, y1, H, W), bse = inb(x1, y1, H, W);
  float wnw = L->s * L->e, wne = L->s * L->w, wsw = L->n * L->e, wse = L->n * L->w;
  if (L->cuda_w) { wnw = L->cw[0]; wne = L->cw[1]; wsw = L->cw[2]; wse = L->cw[3]; }
  int any = 0;
  for (int c = 0; c < C; ++c) {
    const float *pl = img_b + (size_t)c * H * W;
    float vnw = bnw ? pl[(size_t)y0 * W + x0] : 0.0f;
    float vne = bne ? pl[(size_t)y0 * W + x1] : 0.0f;
    float vsw = bsw ? pl[(size_t)y1 * W + x0] : 0.0f;
    float vse = bse ? pl[(size_t)y1 * W + x1] : 0.0f;
    /* nw*w + ne*w + sw*w + se*w, contracted to an FMA chain (verified) */
    float acc = vnw * wnw;
    acc = fmaf(vne, wne, acc);
    acc = fmaf(vsw, wsw, acc);
    acc = fmaf(vse, wse, acc);
    out[c * out_stride] = acc;
    any |= (acc != 0.0f);
  }
  return any;
}

/* F.grid_sample forward on an explicit grid (inverse_warp.py:191). */
DVFO_API void dvfo_grid_sample(const float *img, const float *grid, int B, int C,
                               int H, int W, int padding, float *out) {
  size_t HW = (size_t)H * W;
  for (int b = 0; b < B; ++b)
    for (size_t p = 0; p < HW; ++p) {
      samp_loc L;
      locate(grid[(b * HW + p) * 2], grid[(b * HW + p) * 2 + 1], H, W, padding, &L);
      sample_px(img + (size_t)b * C * HW, C, H, W, &L, out + (size_t)b * C * HW + p, HW);
    }
}

/* inverse_warp forward given P = K @ pose_vec2mat(pose), inverse_warp.py:160-193.
 * valid (optional, uint8 [B,H,W]) = 1 - prod_c(warped_c == 0), loss_functions.py:11. */
DVFO_API void dvfo_inverse_warp_fwd(const float *img, const float *depth,
                                    const float *P, const float *Kinv, int B,
                                    int C, int H, int W, int padding,
                                    float *warped, uint8_t *valid) {
  size_t HW = (size_t)H * W;
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) {
        size_t p = (size_t)i * W + j;
        px_chain c;
        samp_loc L;
        chain_fwd(P + b * 12, Kinv + b * 9, depth[b * HW + p], i, j, H, W, padding, &c);
        locate(c.xn, c.yn, H, W, padding, &L);
        int any = sample_px(img + (size_t)b * C * HW, C, H, W, &L,
                            warped + (size_t)b * C * HW + p, HW);
        if (valid) valid[b * HW + p] = (uint8_t)any;
      }
}

/* ------------------------------------------------------------------------ */
/* backward of inverse_warp (torch autograd's fp32 sequence, restated)       */
/* ------------------------------------------------------------------------ */

/* Given dL/dwarped (gout) produce dL/dimg (scatter, optional), dL/ddepth and
 * dL/dP [B,3,4].  Sequence restated from autograd of inverse_warp.py:
 *   grid_sampler_2d_backward -> index_put(mask) -> sub -> div(w-1) -> mul 2 ->
 *   div(X,Z) -> clamp -> add(P_tr) -> bmm(P_rot,cam) -> mul(depth).
 * Per-pixel arithmetic is fp32 like the reference; the reductions over pixels
 * (dP) are carried in double (the reference's own fp32 GEMM order is not
 * reproducible; double is the more accurate side of the 1e-5 tolerance). */
DVFO_API void dvfo_inverse_warp_bwd(const float *gout, const float *img,
                                    const float *depth, const float *P,
                                    const float *Kinv, int B, int C, int H, int W,
                                    int padding, float *gimg /*nullable, zeroed by callee*/,
                                    float *gdepth, float *gP /*[B,3,4]*/) {
  size_t HW = (size_t)H * W;
  if (gimg) memset(gimg, 0, sizeof(float) * (size_t)B * C * HW);
  for (int b = 0; b < B; ++b) {
    double accP[12];
    for (int k = 0; k < 12; ++k) accP[k] = 0.0;
    const float *Pb = P + b * 12;
    const float *img_b = img + (size_t)b * C * HW;
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) {
        size_t p = (size_t)i * W + j;
        px_chain c;
        samp_loc L;
        chain_fwd(Pb, Kinv + b * 9, depth[b * HW + p], i, j, H, W, padding, &c);
        locate(c.xn, c.yn, H, W, padding, &L);
        int x0 = L.x0, y0 = L.y0, x1 = x0 + 1, y1 = y0 + 1;
        int bnw = inb(x0, y0, H, W), bne = inb(x1, y0, H, W);
        int bsw = inb(x0, y1, H, W), bse = inb(x1, y1, H, W);
        float wnw = L.s * L.e, wne = L.s * L.w, wsw = L.n * L.e, wse = L.n * L.w;
        float gx = 0.0f, gy = 0.0f;
        for (int ch = 0; ch < C; ++ch) {
          const float *pl = img_b + (size_t)ch * HW;
          float g = gout[((size_t)b * C + ch) * HW + p];
          float vnw = bnw ? pl[(size_t)y0 * W + x0] : 0.0f;
          float vne = bne ? pl[(size_t)y0 * W + x1] : 0.0f;
          float vsw = bsw ? pl[(size_t)y1 * W + x0] : 0.0f;
          float vse = bse ? pl[(size_t)y1 * W + x1] : 0.0f;
          if (gimg) {
            float *gp = gimg + ((size_t)b * C + ch) * HW;
            if (bnw) gp[(size_t)y0 * W + x0] += wnw * g;
            if (bne) gp[(size_t)y0 * W + x1] += wne * g;
            if (bsw) gp[(size_t)y1 * W + x0] += wsw * g;
            if (bse) gp[(size_t)y1 * W + x1] += wse * g;
          }
          /* gx += ((ne-nw)*s + (se-sw)*n) * g, contracted exactly like the
           * torch-CPU kernel: t = fma(se-sw, n, (ne-nw)*s); gx = fma(t, g, gx)
           * (verified bit-exact against grid_sampler_2d_backward) */
          gx = fmaf(fmaf(vse - vsw, L.n, (vne - vnw) * L.s), g, gx);
          gy = fmaf(fmaf(vse - vne, L.w, (vsw - vnw) * L.e), g, gy);
        }
        float gxn = c.mx ? 0.0f : gx * L.gmx; /* index_put kills masked coords */
        float gyn = c.my ? 0.0f : gy * L.gmy;
        /* sub(1): identity; div by float(w-1); mul by 2 */
        float gu = (gxn / (float)(W - 1)) * 2.0f;
        float gv = (gyn / (float)(H - 1)) * 2.0f;
        /* div(X,Z): dX = g/Z ; dZ = -g*((X/Z)/Z) */
        float gq0 = gu / c.Z;
        float gq1 = gv / c.Z;
        float gZ = (-gu) * (c.u / c.Z) + (-gv) * (c.v / c.Z);
        /* clamp(min): gradient passes where q_z >= min */
        float gq2 = (c.q[2] >= 1e-3f) ? gZ : 0.0f;
        float gq[3] = {gq0, gq1, gq2};
        /* dP_tr = sum gq ; dP_rot = gq (x) cam */
        for (int r = 0; r < 3; ++r) {
          for (int k = 0; k < 3; ++k)
            accP[r * 4 + k] += (double)gq[r] * (double)c.cam[k];
          accP[r * 4 + 3] += (double)gq[r];
        }
        /* dcam = P_rot^T @ gq ; ddepth = sum_k dcam_k * ray_k */
        float gd = 0.0f;
        for (int k = 0; k < 3; ++k) {
          float gc = dot3_fma(Pb[0 * 4 + k], Pb[1 * 4 + k], Pb[2 * 4 + k], gq[0],
                              gq[1], gq[2]);
          gd += gc * c.ray[k];
        }
        gdepth[b * HW + p] = gd;
      }
    for (int k = 0; k < 12; ++k) gP[b * 12 + k] = (float)accP[k];
  }
}

/* ------------------------------------------------------------------------ */
/* backward of P = K @ [R(r)|t]  w.r.t. the 6-vector (and K)                 */
/* ------------------------------------------------------------------------ */

/* dL/dP [n,3,4] -> dL/dvec [n,6] (+ optional dL/dK [n,3,3]).  Analytic chain
 * of inverse_warp.py:188 (K @ pose_mat), :156 (cat) and :77-138 (euler / quat),
 * evaluated in double from the fp32 inputs. */
DVFO_API void dvfo_pose_bwd(const float *gP, const float *K, const float *vec,
                            int n, int rotation_mode, float *gvec, float *gK) {
  for (int b = 0; b < n; ++b) {
    const float *g = gP + b * 12, *Kb = K + b * 9, *v = vec + b * 6;
    double gM[12]; /* dL/dpose_mat = K^T @ gP */
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += (double)Kb[k * 3 + r] * (double)g[k * 4 + c];
        gM[r * 4 + c] = s;
      }
    double gR[9];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) gR[r * 3 + c] = gM[r * 4 + c];
      gvec[b * 6 + r] = (float)gM[r * 4 + 3];
    }
    if (gK) {
      float pm[12];
      dvfo_pose_vec2mat(v, 1, rotation_mode, pm);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          double s = 0.0;
          for (int k = 0; k < 4; ++k) s += (double)g[r * 4 + k] * (double)pm[c * 4 + k];
          gK[b * 9 + r * 3 + c] = (float)s;
        }
    }
    if (rotation_mode == DVFO_ROT_EULER) {
      double x = v[3], y = v[4], z = v[5];
      double cx = cos(x), sx = sin(x), cy = cos(y), sy = sin(y), cz = cos(z), sz = sin(z);
      double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
      double Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
      double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
      double dRx[9] = {0, 0, 0, 0, -sx, -cx, 0, cx, -sx};
      double dRy[9] = {-sy, 0, cy, 0, 0, 0, -cy, 0, -sy};
      double dRz[9] = {-sz, -cz, 0, cz, -sz, 0, 0, 0, 0};
      const double *A[3][3] = {{dRx, Ry, Rz}, {Rx, dRy, Rz}, {Rx, Ry, dRz}};
      for (int a = 0; a < 3; ++a) {
        double t1[9], t2[9];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += A[a][0][r * 3 + k] * A[a][1][k * 3 + c];
            t1[r * 3 + c] = s;
          }
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += t1[r * 3 + k] * A[a][2][k * 3 + c];
            t2[r * 3 + c] = s;
          }
        double s = 0;
        for (int k = 0; k < 9; ++k) s += gR[k] * t2[k];
        gvec[b * 6 + 3 + a] = (float)s;
      }
    } else {
      /* quat: q = (1,a,b,c)/n ; R(q) quadratic in unit q */
      double q0[4] = {1.0, v[3], v[4], v[5]};
      double n2 = 0;
      for (int k = 0; k < 4; ++k) n2 += q0[k] * q0[k];
      double nn = sqrt(n2);
      double w = q0[0] / nn, x = q0[1] / nn, y = q0[2] / nn, z = q0[3] / nn;
      /* dL/d(unit q) */
      double gw = 0, gx = 0, gy = 0, gz = 0;
      /* R entries, row-major, as in quat2mat_one */
      /* R0 = w2+x2-y2-z2 */ gw += gR[0] * 2 * w; gx += gR[0] * 2 * x; gy -= gR[0] * 2 * y; gz -= gR[0] * 2 * z;
      /* R1 = 2xy-2wz     */ gx += gR[1] * 2 * y; gy += gR[1] * 2 * x; gw -= gR[1] * 2 * z; gz -= gR[1] * 2 * w;
      /* R2 = 2wy+2xz     */ gw += gR[2] * 2 * y; gy += gR[2] * 2 * w; gx += gR[2] * 2 * z; gz += gR[2] * 2 * x;
      /* R3 = 2wz+2xy     */ gw += gR[3] * 2 * z; gz += gR[3] * 2 * w; gx += gR[3] * 2 * y; gy += gR[3] * 2 * x;
      /* R4 = w2-x2+y2-z2 */ gw += gR[4] * 2 * w; gx -= gR[4] * 2 * x; gy += gR[4] * 2 * y; gz -= gR[4] * 2 * z;
      /* R5 = 2yz-2wx     */ gy += gR[5] * 2 * z; gz += gR[5] * 2 * y; gw -= gR[5] * 2 * x; gx -= gR[5] * 2 * w;
      /* R6 = 2xz-2wy     */ gx += gR[6] * 2 * z; gz += gR[6] * 2 * x; gw -= gR[6] * 2 * y; gy -= gR[6] * 2 * w;
      /* R7 = 2wx+2yz     */ gw += gR[7] * 2 * x; gx += gR[7] * 2 * w; gy += gR[7] * 2 * z; gz += gR[7] * 2 * y;
      /* R8 = w2-x2-y2+z2 */ gw += gR[8] * 2 * w; gx -= gR[8] * 2 * x; gy -= gR[8] * 2 * y; gz += gR[8] * 2 * z;
      double gu[4] = {gw, gx, gy, gz}, u[4] = {w, x, y, z};
      double dot = 0;
      for (int k = 0; k < 4; ++k) dot += gu[k] * u[k];
      /* d(q/|q|) : (g - u (g.u)) / |q| ; component 0 is the constant 1 */
      for (int k = 1; k < 4; ++k) gvec[b * 6 + 3 + (k - 1)] = (float)((gu[k] - u[k] * dot) / nn);
    }
  }
}

/* ------------------------------------------------------------------------ */
/* masked photometric / feature reconstruction loss (one scale, V views)     */
/* ------------------------------------------------------------------------ */

/* loss_functions.py:7-20 (V=2, no mask), loss_functions_sfm.py:10-36 (V refs,
 * optional explainability mask), loss_function_sfm_old.py:8-36:
 *   warped_v = inverse_warp(src_v, depth, P_v)
 *   valid_v  = 1 - prod_c(warped_v == 0)
 *   diff_v   = (tgt - warped_v) * valid_v [* expl[:,v]]
 *   term_v   = mean(|diff_v|)  over B*C*H*W
 * and, for upstream dL/dterm_v = 1, the gradients w.r.t. depth (summed over
 * views), P_v, expl, src_v and tgt.  Any output pointer may be NULL.
 *   srcs  : V pointers to [B,C,H,W];  P: [B,V,3,4];  expl: [B,V,H,W] or NULL
 *   terms : V doubles;  gP: [B,V,3,4];  gsrc: V pointers (each nullable)
 *   valid : uint8 [V,B,H,W]                                                  */
DVFO_API void dvfo_photo_loss(const float *tgt, const float *const *srcs,
                              const float *depth, const float *P,
                              const float *Kinv, const float *expl, int B, int C,
                              int H, int W, int V, int padding, double *terms,
                              float *gdepth, float *gP, float *gexpl,
                              float *const *gsrc, float *gtgt, uint8_t *valid) {
  size_t HW = (size_t)H * W, N = (size_t)B * C * HW;
  float inv_n = 1.0f / (float)N; /* autograd: grad of mean = g.expand / N */
  float *warped = (float *)malloc(sizeof(float) * N);
  float *gw = (float *)malloc(sizeof(float) * N);
  float *gd_v = (float *)malloc(sizeof(float) * (size_t)B * HW);
  float *Pv = (float *)malloc(sizeof(float) * (size_t)B * 12);
  float *gPv = (float *)malloc(sizeof(float) * (size_t)B * 12);
  uint8_t *val = (uint8_t *)malloc((size_t)B * HW);
  if (gdepth) memset(gdepth, 0, sizeof(float) * (size_t)B * HW);
  if (gtgt) memset(gtgt, 0, sizeof(float) * N);
  for (int v = 0; v < V; ++v) {
    for (int b = 0; b < B; ++b) memcpy(Pv + b * 12, P + ((size_t)b * V + v) * 12, 48);
    dvfo_inverse_warp_fwd(srcs[v], depth, Pv, Kinv, B, C, H, W, padding, warped, val);
    if (valid) memcpy(valid + (size_t)v * B * HW, val, (size_t)B * HW);
    double acc = 0.0;
    for (int b = 0; b < B; ++b)
      for (size_t p = 0; p < HW; ++p) {
        float m = val[b * HW + p] ? 1.0f : 0.0f;
        float ex = expl ? expl[((size_t)b * V + v) * HW + p] : 1.0f;
        float ge = 0.0f;
        for (int c = 0; c < C; ++c) {
          size_t o = ((size_t)b * C + c) * HW + p;
          float d0 = (tgt[o] - warped[o]) * m; /* loss_functions.py:12 */
          float d1 = expl ? d0 * ex : d0;      /* loss_functions_sfm.py:31 */
          acc += fabs((double)d1);
          float sg = (d1 > 0.0f) ? 1.0f : ((d1 < 0.0f) ? -1.0f : 0.0f); /* sign(0)=0 */
          float gd1 = sg * inv_n;
          float gd0 = expl ? gd1 * ex : gd1;
          if (expl) ge += gd1 * d0;
          float gdiff = gd0 * m; /* d(tgt - warped) */
          gw[o] = -gdiff;
          if (gtgt) gtgt[o] += gdiff;
        }
        if (gexpl) gexpl[((size_t)b * V + v) * HW + p] = ge;
      }
    terms[v] = acc / (double)N;
    if (gdepth || gP || (gsrc && gsrc[v])) {
      dvfo_inverse_warp_bwd(gw, srcs[v], depth, Pv, Kinv, B, C, H, W, padding,
                            (gsrc && gsrc[v]) ? gsrc[v] : NULL, gd_v, gPv);
      if (gdepth)
        for (size_t k = 0; k < (size_t)B * HW; ++k) gdepth[k] += gd_v[k];
      if (gP)
        for (int b = 0; b < B; ++b) memcpy(gP + ((size_t)b * V + v) * 12, gPv + b * 12, 48);
    }
  }
  free(warped); free(gw); free(gd_v); free(Pv); free(gPv); free(val);
}

/* ------------------------------------------------------------------------ */
/* neighbours of the path: area down-sampling, smoothness, explainability    */
/* ------------------------------------------------------------------------ */

/* F.interpolate(img, (h,w), mode='area') for integer factors
 * (loss_functions_sfm.py:18-19) == adaptive_avg_pool2d: fp32 row-major window
 * sum, then division by the window size. */
DVFO_API void dvfo_area_downsample(const float *img, int BC, int H, int W, int h,
                                   int w, float *out) {
  for (int n = 0; n < BC; ++n)
    for (int oy = 0; oy < h; ++oy)
      for (int ox = 0; ox < w; ++ox) {
        int ys = (int)floorf((float)(oy * H) / h), ye = (int)ceilf((float)((oy + 1) * H) / h);
        int xs = (int)floorf((float)(ox * W) / w), xe = (int)ceilf((float)((ox + 1) * W) / w);
        float sum = 0.0f;
        for (int y = ys; y < ye; ++y)
          for (int x = xs; x < xe; ++x) sum += img[((size_t)n * H + y) * W + x];
        out[((size_t)n * h + oy) * w + ox] = sum / (float)((ye - ys) * (xe - xs));
      }
}

/* smooth_loss for one map [B,1,H,W] (loss_functions.py:23-41): sum of the means
 * of |dxx|,|dxy|,|dyx|,|dyy|; optional gradient for upstream 1 (weight applied
 * by the caller). */
DVFO_API double dvfo_smooth_loss(const float *d, int B, int H, int W, float *gd) {
  size_t HW = (size_t)H * W;
  double total = 0.0;
  if (gd) memset(gd, 0, sizeof(float) * B * HW);
  double sxx = 0, sxy = 0, syx = 0, syy = 0;
  size_t nxx = (size_t)B * H * (W > 2 ? W - 2 : 0), nxy = (size_t)B * (H > 1 ? H - 1 : 0) * (W > 1 ? W - 1 : 0);
  size_t nyy = (size_t)B * (H > 2 ? H - 2 : 0) * W;
#define SGN(v) ((v) > 0.0f ? 1.0f : ((v) < 0.0f ? -1.0f : 0.0f))
  for (int b = 0; b < B; ++b) {
    const float *m = d + b * HW;
    float *g = gd ? gd + b * HW : NULL;
    for (int y = 0; y < H; ++y)
      for (int x = 0; x + 2 < W; ++x) { /* dx2 = dx[x+1]-dx[x] */
        float dx0 = m[y * W + x + 1] - m[y * W + x], dx1 = m[y * W + x + 2] - m[y * W + x + 1];
        float v = dx1 - dx0;
        sxx += fabs((double)v);
        if (g) { float s = SGN(v) / (float)nxx; g[y * W + x + 2] += s; g[y * W + x + 1] -= 2 * s; g[y * W + x] += s; }
      }
    for (int y = 0; y + 1 < H; ++y)
      for (int x = 0; x + 1 < W; ++x) { /* dxdy (of dx along y) and dydx (of dy along x): same stencil */
        float dxa = m[y * W + x + 1] - m[y * W + x], dxb = m[(y + 1) * W + x + 1] - m[(y + 1) * W + x];
        float v1 = dxb - dxa;
        float dya = m[(y + 1) * W + x] - m[y * W + x], dyb = m[(y + 1) * W + x + 1] - m[y * W + x + 1];
        float v2 = dyb - dya;
        sxy += fabs((double)v1);
        syx += fabs((double)v2);
        if (g) {
          float s1 = SGN(v1) / (float)nxy, s2 = SGN(v2) / (float)nxy, s = s1 + s2;
          g[(y + 1) * W + x + 1] += s; g[(y + 1) * W + x] -= s; g[y * W + x + 1] -= s; g[y * W + x] += s;
        }
      }
    for (int y = 0; y + 2 < H; ++y)
      for (int x = 0; x < W; ++x) {
        float dy0 = m[(y + 1) * W + x] - m[y * W + x], dy1 = m[(y + 2) * W + x] - m[(y + 1) * W + x];
        float v = dy1 - dy0;
        syy += fabs((double)v);
        if (g) { float s = SGN(v) / (float)nyy; g[(y + 2) * W + x] += s; g[(y + 1) * W + x] -= 2 * s; g[y * W + x] += s; }
      }
  }
#undef SGN
  if (nxx) total += sxx / (double)nxx;
  if (nxy) total += sxy / (double)nxy + syx / (double)nxy;
  if (nyy) total += syy / (double)nyy;
  return total;
}

/* explainability_loss for one mask tensor (loss_functions_sfm.py:49-56):
 * binary_cross_entropy(mask, ones) = mean(-max(log(mask), -100)). */
DVFO_API double dvfo_explainability_loss(const float *mask, size_t n, float *gmask) {
  double acc = 0.0;
  for (size_t k = 0; k < n; ++k) {
    float l = logf(mask[k]);
    if (l < -100.0f) l = -100.0f;
    acc -= (double)l;
    /* torch: grad = (x - y) / max((1-x)*x, 1e-12) / n, with y = 1 */
    if (gmask) {
      float x = mask[k];
      float den = (1.0f - x) * x;
      if (den < 1e-12f) den = 1e-12f;
      gmask[k] = ((x - 1.0f) / den) / (float)n;
    }
  }
  return acc / (double)n;
}

/* ------------------------------------------------------------------------ */
/* se(3) -> SE(3) exponential map  (pytorch_version/se3_generate.py:7-103)   */
/* ------------------------------------------------------------------------ */

/* forward, se3_generate.py:9-54.  in: [B,6] fp32 = (w(3), u(3));  out: [B,4,4] fp64 = [[R, R u],[0,1]].
 * Mixed precision as in the reference: theta, c1, c2 in fp32 (numpy float32 scalars), the matrix algebra in fp64. */
static void se3_R(const float *w, double *R, float *theta_out) {
  double wx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}; /* :15-21 */
  float theta = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);   /* np.linalg.norm on float32 */
  for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
  if (theta * theta < 1e-12f) {                                     /* :33 */
    for (int k = 0; k < 9; ++k) R[k] += wx[k];
  } else {
    float c1 = sinf(theta) / theta;                                 /* :37 */
    float sh = sinf(theta / 2);
    float c2 = 2 * (sh * sh) / (theta * theta);                    /* :38 */
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += wx[r * 3 + k] * wx[k * 3 + c];
        R[r * 3 + c] += (double)c1 * wx[r * 3 + c] + (double)c2 * s; /* :42 */
      }
  }
  *theta_out = theta;
}

DVFO_API void dvfo_se3_exp_fwd(const float *in, int B, double *out) {
  for (int b = 0; b < B; ++b) {
    const float *w = in + b * 6, *u = in + b * 6 + 3;
    double R[9];
    float th;
    se3_R(w, R, &th);
    double *o = out + b * 16;
    for (int k = 0; k < 16; ++k) o[k] = 0.0;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) o[r * 4 + c] = R[r * 3 + c];
      o[r * 4 + 3] = R[r * 3] * (double)u[0] + R[r * 3 + 1] * (double)u[1] + R[r * 3 + 2] * (double)u[2]; /* :46 */
    }
    o[15] = 1.0;
  }
}

/* backward, se3_generate.py:56-103.  gout: [B,4,4] fp64;  gin: [B,6] fp32. */
DVFO_API void dvfo_se3_exp_bwd(const float *in, const double *gout, int B, float *gin) {
  static const double gen[3][9] = {{0, 0, 0, 0, 0, 1, 0, -1, 0}, {0, 0, -1, 0, 0, 0, 1, 0, 0}, {0, 1, 0, -1, 0, 0, 0, 0, 0}}; /* :79-81 */
  for (int b = 0; b < B; ++b) {
    const float *w = in + b * 6, *u = in + b * 6 + 3;
    const double *g = gout + b * 16;
    double R[9];
    float th;
    se3_R(w, R, &th);
    double wx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double dT[3] = {g[3], g[7], g[11]};
    for (int c = 0; c < 3; ++c) /* dLdut = dLdT @ R  :68 */
      gin[b * 6 + 3 + c] = (float)(dT[0] * R[c] + dT[1] * R[3 + c] + dT[2] * R[6 + c]);
    double dR[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) dR[r * 3 + c] = g[r * 4 + c] + dT[r] * (double)u[c]; /* :71-75 */
    for (int idx = 0; idx < 3; ++idx) {
      double dRdw[9];
      if (th * th < 1e-12f) {
        for (int k = 0; k < 9; ++k) dRdw[k] = gen[idx][k]; /* :97 */
      } else {
        /* cross_term = wx @ ((I - R) e_idx) :87 ; cross = skew(cross_term) :88-94 */
        double col[3], ct[3];
        for (int r = 0; r < 3; ++r) col[r] = ((r == idx) ? 1.0 : 0.0) - R[r * 3 + idx];
        for (int r = 0; r < 3; ++r) ct[r] = wx[r * 3] * col[0] + wx[r * 3 + 1] * col[1] + wx[r * 3 + 2] * col[2];
        double cr[9] = {0, -ct[2], ct[1], ct[2], 0, -ct[0], -ct[1], ct[0], 0};
        double th2 = (double)(th * th); /* theta[j]**2 is a float32 square */
        double A[9];
        for (int k = 0; k < 9; ++k) A[k] = ((double)w[idx] * wx[k] + cr[k]) / th2; /* :99 */
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) dRdw[r * 3 + c] = A[r * 3] * R[c] + A[r * 3 + 1] * R[3 + c] + A[r * 3 + 2] * R[6 + c];
      }
      double s = 0;
      for (int k = 0; k < 9; ++k) s += dR[k] * dRdw[k]; /* :100 */
      gin[b * 6 + idx] = (float)s;
    }
  }
}

/* ------------------------------------------------------------------------ */
/* Caffe-convention layers: GeoTransform, PinHole,            */
/* InverseWarping, AbsLoss -- restated from the .cu files in caffe/src/caffe/layers. */
/* PARITY UNPINNED for this block: BVLC Caffe is not vendored, the layers    */
/* cannot be built or run here; the restatement follows the kernel sources   */
/* expression by expression (same C promotion rules: `Z+1e-12` is double).   */
/* Layouts: depth [N,H,W], T [N,16] (row-major 4x4), K [N,4]=(fx,fy,cx,cy),  */
/* pts [N,3,H,W], coords [N,2,H,W] in PIXEL units, images [N,C,H,W].         */
/* ------------------------------------------------------------------------ */

/* GeoTransform forward, geometry_transformation.cu:10-47 */
DVFO_API void dvfo_caffe_geo_fwd(const float *depth, const float *T, const float *K, int N, int H, int W, float *pts) {
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const float *t = T + n * 16, *k = K + n * 4;
        const float fx = k[0], fy = k[1], cx = k[2], cy = k[3];
        const float d = depth[((size_t)n * H + y) * W + x];
        const float X = (x - cx) / fx * d, Y = (y - cy) / fy * d;
        pts[(((size_t)n * 3 + 0) * H + y) * W + x] = t[0] * X + t[1] * Y + t[2] * d + t[3];
        pts[(((size_t)n * 3 + 1) * H + y) * W + x] = t[4] * X + t[5] * Y + t[6] * d + t[7];
        pts[(((size_t)n * 3 + 2) * H + y) * W + x] = t[8] * X + t[9] * Y + t[10] * d + t[11];
      }
}

/* GeoTransform backward, geometry_transformation.cu:72-176 (atomics replaced by ordered fp64 sums) */
DVFO_API void dvfo_caffe_geo_bwd(const float *top, const float *depth, const float *T, const float *K, int N, int H, int W,
                                 float *depth_diff, float *T_diff /*[N,16]*/, float *K_diff /*[N,4]*/) {
  for (int n = 0; n < N; ++n) {
    double accT[16] = {0}, accK[4] = {0};
    const float *t = T + n * 16, *k = K + n * 4;
    const float fx = k[0], fy = k[1], cx = k[2], cy = k[3];
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const size_t o = ((size_t)n * H + y) * W + x;
        const float gx = top[(((size_t)n * 3 + 0) * H + y) * W + x], gy = top[(((size_t)n * 3 + 1) * H + y) * W + x],
                    gz = top[(((size_t)n * 3 + 2) * H + y) * W + x];
        const float bX = (x - cx) / fx, bY = (y - cy) / fy, d = depth[o];
        float dd = 0;
        dd += gx * (t[0] * bX + t[1] * bY + t[2]);
        dd += gy * (t[4] * bX + t[5] * bY + t[6]);
        dd += gz * (t[8] * bX + t[9] * bY + t[10]);
        depth_diff[o] = dd;
        const float g3[3] = {gx, gy, gz};
        for (int r = 0; r < 3; ++r) {
          accT[r * 4 + 0] += g3[r] * bX * d;
          accT[r * 4 + 1] += g3[r] * bY * d;
          accT[r * 4 + 2] += g3[r] * d;
          accT[r * 4 + 3] += g3[r] * 1.0f;
        }
        float sx = 0, sy = 0;
        sx += gx * t[0]; sx += gy * t[4]; sx += gz * t[8];
        sy += gx * t[1]; sy += gy * t[5]; sy += gz * t[9];
        accK[2] += sx * (-d / fx);
        accK[3] += sy * (-d / fy);
        accK[0] += sx * (-bX / fx * d);
        accK[1] += sy * (-bY / fy * d);
      }
    for (int q = 0; q < 16; ++q) T_diff[n * 16 + q] = (float)accT[q];
    for (int q = 0; q < 4; ++q) K_diff[n * 4 + q] = (float)accK[q];
  }
}

/* PinHoleProjection forward, pin_hole_layer.cu:10-50 (proj_coords output; the flow output is always silenced) */
DVFO_API void dvfo_caffe_pinhole_fwd(const float *pts, const float *K, int N, int H, int W, float *coords) {
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const float *k = K + n * 4;
        const float fx = k[0], fy = k[1], cx = k[2], cy = k[3];
        const float X = pts[(((size_t)n * 3 + 0) * H + y) * W + x], Y = pts[(((size_t)n * 3 + 1) * H + y) * W + x],
                    Z = pts[(((size_t)n * 3 + 2) * H + y) * W + x];
        coords[(((size_t)n * 2 + 0) * H + y) * W + x] = (float)(fx * X / (Z + 1e-12) + cx);
        coords[(((size_t)n * 2 + 1) * H + y) * W + x] = (float)(fy * Y / (Z + 1e-12) + cy);
      }
}

/* PinHole backward, pin_hole_layer.cu:76-146 with top_flows_diff = 0 */
DVFO_API void dvfo_caffe_pinhole_bwd(const float *cdiff, const float *pts, const float *K, int N, int H, int W, float *pts_diff,
                                     float *K_diff) {
  for (int n = 0; n < N; ++n) {
    double accK[4] = {0};
    const float fx = K[n * 4], fy = K[n * 4 + 1];
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const size_t oX = (((size_t)n * 3 + 0) * H + y) * W + x, oY = (((size_t)n * 3 + 1) * H + y) * W + x,
                     oZ = (((size_t)n * 3 + 2) * H + y) * W + x;
        const float gx = cdiff[(((size_t)n * 2 + 0) * H + y) * W + x], gy = cdiff[(((size_t)n * 2 + 1) * H + y) * W + x];
        const float X = pts[oX], Y = pts[oY], Z = pts[oZ];
        float v;
        v = 0; v += (float)(gx * fx / (Z + 1e-12)); pts_diff[oX] = v;
        v = 0; v += (float)(gy * fy / (Z + 1e-12)); pts_diff[oY] = v;
        v = 0; v += (float)(-gx * fx * X / (Z * Z + 1e-12)); v += (float)(-gy * fy * Y / (Z * Z + 1e-12)); pts_diff[oZ] = v;
        accK[2] += gx;
        accK[3] += gy;
        accK[0] += (float)(gx * X / (Z + 1e-12));
        accK[1] += (float)(gy * Y / (Z + 1e-12));
      }
    for (int q = 0; q < 4; ++q) K_diff[n * 4 + q] = (float)accK[q];
  }
}

/* InverseWarping forward, inverse_warping_layer.cu:10-52 (pixel-space bilinear, per-tap zero outside the image) */
DVFO_API void dvfo_caffe_warp_fwd(const float *U, const float *xy, int N, int C, int H, int W, float *out) {
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const float xx = xy[(((size_t)n * 2 + 0) * H + y) * W + x], yy = xy[(((size_t)n * 2 + 1) * H + y) * W + x];
        const int x1 = (int)floorf(xx), x2 = x1 + 1, y1 = (int)floorf(yy), y2 = y1 + 1;
        const float wx2 = xx - (float)x1, wx1 = (float)x2 - xx, wy2 = yy - (float)y1, wy1 = (float)y2 - yy;
        for (int c = 0; c < C; ++c) {
          const float *pl = U + ((size_t)n * C + c) * H * W;
          float v = 0;
          if (x1 >= 0 && x1 <= W - 1 && y1 >= 0 && y1 <= H - 1) v += wx1 * wy1 * pl[x1 + y1 * W];
          if (x1 >= 0 && x1 <= W - 1 && y2 >= 0 && y2 <= H - 1) v += wx1 * wy2 * pl[x1 + y2 * W];
          if (x2 >= 0 && x2 <= W - 1 && y1 >= 0 && y1 <= H - 1) v += wx2 * wy1 * pl[x2 + y1 * W];
          if (x2 >= 0 && x2 <= W - 1 && y2 >= 0 && y2 <= H - 1) v += wx2 * wy2 * pl[x2 + y2 * W];
          out[((size_t)n * C + c) * H * W + x + y * W] = v;
        }
      }
}

/* InverseWarping backward, inverse_warping_layer.cu:84-169 */
DVFO_API void dvfo_caffe_warp_bwd(const float *top, const float *U, const float *xy, int N, int C, int H, int W,
                                  float *U_diff /*nullable, zeroed here*/, float *xy_diff) {
  if (U_diff) memset(U_diff, 0, sizeof(float) * (size_t)N * C * H * W);
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const size_t ox = (((size_t)n * 2 + 0) * H + y) * W + x, oy = (((size_t)n * 2 + 1) * H + y) * W + x;
        const float xx = xy[ox], yy = xy[oy];
        const int x1 = (int)floorf(xx), x2 = x1 + 1, y1 = (int)floorf(yy), y2 = y1 + 1;
        const float wx2 = xx - (float)x1, wx1 = (float)x2 - xx, wy2 = yy - (float)y1, wy1 = (float)y2 - yy;
        float tl = 0, tr = 0, bl = 0, br = 0;
        for (int c = 0; c < C; ++c) {
          const size_t off = ((size_t)n * C + c) * H * W;
          const float g = top[off + (size_t)W * y + x];
          if (x1 >= 0 && x1 <= W - 1 && y1 >= 0 && y1 <= H - 1) { if (U_diff) U_diff[off + x1 + y1 * W] += g * wx1 * wy1; tl += g * U[off + W * y1 + x1]; }
          if (x1 >= 0 && x1 <= W - 1 && y2 >= 0 && y2 <= H - 1) { if (U_diff) U_diff[off + x1 + y2 * W] += g * wx1 * wy2; bl += g * U[off + W * y2 + x1]; }
          if (x2 >= 0 && x2 <= W - 1 && y1 >= 0 && y1 <= H - 1) { if (U_diff) U_diff[off + x2 + y1 * W] += g * wx2 * wy1; tr += g * U[off + W * y1 + x2]; }
          if (x2 >= 0 && x2 <= W - 1 && y2 >= 0 && y2 <= H - 1) { if (U_diff) U_diff[off + x2 + y2 * W] += g * wx2 * wy2; br += g * U[off + W * y2 + x2]; }
        }
        xy_diff[ox] = (tr - tl) * wy1 + (br - bl) * wy2;
        xy_diff[oy] = (bl - tl) * wx1 + (br - tr) * wx2;
      }
}

/* AbsLoss, abs_loss_layer.cu:10-50: loss = sum|a-b| / num ; d/da = alpha * ((d>0) - (d<=0)), alpha = w/num */
DVFO_API double dvfo_caffe_abs_loss(const float *a, const float *b, size_t count, int num, float weight, float *ga, float *gb) {
  double s = 0;
  const float alpha = weight / (float)num;
  for (size_t i = 0; i < count; ++i) {
    const float d = a[i] - b[i];
    s += fabs((double)d);
    const float sg = (float)((d > 0) - (d <= 0));
    if (ga) ga[i] = alpha * sg;
    if (gb) gb[i] = -alpha * sg;
  }
  return s / (double)num;
}

/* Edge-aware smoothness of the Caffe graphs (experiments/depth/train.prototxt:4022-4234; EdgeX / EdgeY fillers
 * caffe/include/caffe/filler.hpp:266-315; AbsLoss caffe/src/caffe/layers/abs_loss_layer.cu:10-50).  PARITY UNPINNED like the
 * other Caffe-convention blocks (BVLC Caffe is not buildable here): the layer chain is restated operator by operator on
 * materialised blobs -- convolution without padding, AbsVal, 1x1 convolution with -0.33, Exp, Eltwise PROD, AbsLoss against
 * zeros -- and back-propagated layer by layer (scatter form), so that the CUDA kernel's fused gather form is checked against
 * an independently structured computation.  loss[0..1] un-weighted, ginv = d(weight*(loss0+loss1))/d inv_depth. */
DVFO_API void dvfo_caffe_edge_smooth(const float *img, const float *invd, int N, int H, int W, float weight, double *loss,
                                     float *ginv) {
  const int h = H - 2, w = W - 2, HW = H * W;
  float *gx = (float *)malloc(sizeof(float) * h * w), *gy = (float *)malloc(sizeof(float) * h * w);
  float *ex = (float *)malloc(sizeof(float) * h * w), *ey = (float *)malloc(sizeof(float) * h * w);
  double lx = 0.0, ly = 0.0;
  if (ginv) memset(ginv, 0, sizeof(float) * (size_t)N * HW);
  for (int n = 0; n < N; ++n) {
    const float *I = img + (size_t)n * 3 * HW, *D = invd + (size_t)n * HW;
    for (int i = 0; i < h; ++i)
      for (int j = 0; j < w; ++j) {
        float sx = 0.0f, sy = 0.0f;
        for (int c = 0; c < 3; ++c) {
          const float *p = I + (size_t)c * HW;
          sx = sx + fabsf(0.5f * (p[(i + 2) * W + j + 1] - p[i * W + j + 1]));   /* EdgeX: rows below - above */
          sy = sy + fabsf(0.5f * (p[(i + 1) * W + j + 2] - p[(i + 1) * W + j])); /* EdgeY: columns right - left */
        }
        gx[i * w + j] = expf(-0.33f * sx);
        gy[i * w + j] = expf(-0.33f * sy);
        ex[i * w + j] = gx[i * w + j] * (0.5f * (D[(i + 2) * W + j + 1] - D[i * W + j + 1]));
        ey[i * w + j] = gy[i * w + j] * (0.5f * (D[(i + 1) * W + j + 2] - D[(i + 1) * W + j]));
        lx += fabs((double)ex[i * w + j]);
        ly += fabs((double)ey[i * w + j]);
      }
    if (ginv) {
      float *G = ginv + (size_t)n * HW;
      const float alpha = weight / (float)N;
      for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
          const float vx = ex[i * w + j], vy = ey[i * w + j];
          const float sx = vx != vx ? 0.0f : (vx >= 0.0f ? 1.0f : -1.0f), sy = vy != vy ? 0.0f : (vy >= 0.0f ? 1.0f : -1.0f);
          const float tx = 0.5f * ((alpha * sx) * gx[i * w + j]), ty = 0.5f * ((alpha * sy) * gy[i * w + j]);
          G[(i + 2) * W + j + 1] += tx;
          G[i * W + j + 1] -= tx;
          G[(i + 1) * W + j + 2] += ty;
          G[(i + 1) * W + j] -= ty;
        }
    }
  }
  loss[0] = lx / (double)N;
  loss[1] = ly / (double)N;
  free(gx); free(gy); free(ex); free(ey);
}

/* SSIM reconstruction term -- NOT in the reference (the checkout contains no SSIM): PARITY UNPINNED by construction.  It
 * restates the definition written in depth-vo-feat_b200/csrc/dvf_ssim.cu (3x3 average-pool SSIM without padding,
 * C1 = 0.01^2, C2 = 0.03^2, l = clamp((1 - SSIM)/2, 0, 1), mean over B*C*(H-2)*(W-2) windows, windows touching an invalid
 * pixel contribute zero) in fp64, window by window, and back-propagates each window to its nine pixels (scatter form;
 * the kernel gathers).  tests/ additionally compare it with a torch restatement (avg_pool2d + autograd). */
DVFO_API double dvfo_ssim_loss(const float *x, const float *y, const unsigned char *valid, int B, int C, int H, int W,
                               float *gy) {
  const double C1 = 0.01 * 0.01, C2 = 0.03 * 0.03, N = (double)B * C * (H - 2) * (W - 2);
  double total = 0.0;
  double *g = gy ? (double *)calloc((size_t)B * C * H * W, sizeof(double)) : NULL;
  for (int p = 0; p < B * C; ++p) {
    const float *xp = x + (size_t)p * H * W, *yp = y + (size_t)p * H * W;
    const unsigned char *vp = valid ? valid + (size_t)(p / C) * H * W : NULL;
    for (int i = 0; i < H - 2; ++i)
      for (int j = 0; j < W - 2; ++j) {
        double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
        int m = 1;
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) {
            const double u = xp[(i + a) * W + j + b], v = yp[(i + a) * W + j + b];
            sx += u; sy += v; sxx += u * u; syy += v * v; sxy += u * v;
            if (vp && !vp[(i + a) * W + j + b]) m = 0;
          }
        if (!m) continue;
        const double mx = sx / 9, my = sy / 9, vx = sxx / 9 - mx * mx, vy = syy / 9 - my * my, cxy = sxy / 9 - mx * my;
        const double A1 = 2 * mx * my + C1, A2 = 2 * cxy + C2, B1 = mx * mx + my * my + C1, B2 = vx + vy + C2;
        const double S = (A1 * A2) / (B1 * B2), l = 0.5 * (1 - S);
        total += l < 0 ? 0 : (l > 1 ? 1 : l);
        if (g && l > 0 && l < 1) {
          const double d = B1 * B2, k = -0.5 / N;
          const double dmy = ((2 * mx * A2 - 2 * mx * A1) * d - (A1 * A2) * (2 * my * B2 - 2 * my * B1)) / (d * d);
          const double dEyy = -(A1 * A2) * B1 / (d * d), dExy = 2 * A1 * B2 * B1 / (d * d);
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
              const size_t q = (size_t)p * H * W + (i + a) * W + j + b;
              g[q] += k * (dmy + 2.0 * yp[(i + a) * W + j + b] * dEyy + xp[(i + a) * W + j + b] * dExy) / 9.0;
            }
        }
      }
  }
  if (g) {
    for (size_t q = 0; q < (size_t)B * C * H * W; ++q) gy[q] = (float)g[q];
    free(g);
  }
  return total / N;
}
