// dvf_loss_kernel.cuh -- device code of the fused masked reconstruction loss (forward + backward in
// ONE pass over HBM).  See dvf_loss.cu for the host side and the reference citations.
//
// One launch covers every pyramid level and every source view of a loss call:
//   * the launch is a list of 256-pixel units; every resident CTA owns an equal share of it and walks it in
//     PIECES, contiguous runs of target pixels of ONE image of ONE level, so the per-image projection P and
//     K^-1 are uniform over a piece (kept in shared memory, broadcast reads);
//   * each thread owns adjacent pixels of the run, so every depth / target load and depth-gradient store of a
//     warp is contiguous; in the image kernel these streaming inputs arrive through a TMA-filled ring;
//   * per pixel the depth / target values are read once and shared by all V views; the 4 bilinear
//     taps per channel are read-only gathers (neighbouring lanes hit neighbouring texels; L1 absorbs
//     the x0/x1 and y0/y1 reuse);
//   * the hot path is branch-free straight-line code; pixels whose operands leave the range in which
//     the shared-reciprocal divisions are exact (|q| > 2^100, NaN) are redone by a cold out-of-line
//     routine using __fdiv_rn;
//   * loss term and the 12 entries of dL/dP are accumulated per thread, folded with a 16-slot
//     butterfly, then per piece, and the LAST piece of an image (atomic ticket) adds the partials in
//     a fixed order in fp64 -> deterministic, no output needs pre-zeroing, no second launch.
#pragma once
#include <type_traits>

#include "dvf_internal.h"
#include "dvf_math.cuh"
#include "dvf_math2.cuh"
#include "dvf_pose.cuh"
#include "dvf_reduce.cuh"
#include "dvf_tma.cuh"

namespace dvf {

constexpr int kLossThreads = 128;

#ifdef DVF_TRACE   // experiment builds: per-CTA timestamps of the first piece (profiles/trace_pieces.py)
static __device__ unsigned long long g_trace[8192 * 8];   // one copy per translation unit
__device__ __forceinline__ void trace_mark(int slot) {
  if (threadIdx.x == 0 && blockIdx.x < 8192) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[blockIdx.x * 8 + slot] = t;
  }
}
#define DVF_MARK(slot) trace_mark(slot)
#else
#define DVF_MARK(slot)
#endif

struct LevelDev {
  int H, W, HW;
  FastDiv divW;
  Geo geo;
  float inv_n;
  float ds;           // downscale of this level (pose mode)
  int allow_fast;
  int prefetch_rows;  // L2 prefetch of the source band ahead of the ring: rows of parallax covered, 0 = off
  const float* depth;
  const float* tgt;
  const float* src[DVF_MAX_VIEWS];
  const float* expl;
  long long expl_bstride;
  const float* P;
  const float* Kinv;
  float* gdepth;
  float* gexpl;
  float* gsrc[DVF_MAX_VIEWS];
  float* gtgt;
  float* gP;
  int block_begin, blocks_per_image, iters;
  int px_per_cta;         // contiguous target pixels owned by one CTA (multiple of kPlanUnit)
  // balanced (persistent) split of the image kernel: the launch is a list of 256-pixel units, image-major, then
  // level, then position, each (image, level) preceded by `piece_overhead` empty units that stand for its fixed
  // cost; CTA c of G owns units [c*T/G, (c+1)*T/G) and may cross level and image boundaries
  int unit_base;          // offset of this level inside an image's unit list (including the overhead units)
  int units_per_image;    // 256-pixel units of one image of this level
  int slots_per_image;    // partial-sum slots reserved per image (>= CTAs that can touch one image)
  float* partials;        // [B*slots_per_image][V][kRedSlots]
  double* img_terms;      // [B][V]
  unsigned* img_counter;  // [B]   zero between launches
  unsigned* lvl_counter;  // [1]   zero between launches
};

struct LossParams {
  int n_levels, B, C, V;
  int need_grad, rotation;
  int total_units;         // balanced split: units of the launch = B * units_per_image_all
  int units_per_image_all; // units of one image over all levels, overhead units included
  int piece_overhead;      // empty units in front of every (image, level)
  float* terms;  // [n_levels*V]
  int mean_batch;          // batch size in the denominators (>= B: this launch may hold a shard of a larger batch)
  int ctas_per_sm;         // > 0: grid override of the balanced kernels (tuning)
  int grad_bf16;           // NHWC bf16 maps: map gradients are bf16
  int pdl;                 // programmatic dependent launch: 1 = DVF_FLAG_PDL, 2 = also DVF_FLAG_PDL_CHAINED
  // producer glue folded into the kernel (dvf_loss_desc): `depth` holds disparities, depth = 1 / (disp + disp_eps)
  // (unsupervise.py:99, train.py:188) and gdepth receives d/d disparity; images are multiplied by img_scale on load
  // (unsupervise.py:101: 0.004 * img)
  int disparity;
  float disp_eps, img_scale;   // img_scale 1 = none
  // balanced split in 32-bit arithmetic (set by launch_balanced once the grid G is known): CTA c owns units
  // [c*q + floor(c*r/G), ...) with T = q*G + r -- the same values as floor(c*T/G) without 64-bit divisions, which cost
  // ~1 us per piece when every issue slot of the SM is contended (profiles/trace_pieces.py)
  int split_q, split_r, grid;
  float ctas_per_unit;     // G / T
  FastDiv div_grid, div_upi;   // division by G and by units_per_image_all
  const float* upstream;   // device scalar multiplying every gradient, or nullptr (= 1)
  int* nan_flags;          // nullptr, or a word that collects bit l*V+v when term (l, v) is NaN
  // fused exchange of the loss terms (dvf_loss_desc.peer_terms): row peer_rank of every peer's [n_peers][n_levels*V] buffer
  int n_peers, peer_rank;
  float* peer_terms[DVF_MAX_PEERS];
  // pose mode (dvf_photo_loss_fused_pose): P / K^-1_s are derived in the CTA prologue, d pose in the epilogue
  const float* pose_vec;   // [B,V,6] or nullptr
  const float* K;          // [B,3,3]
  const float* Kinv;       // [B,3,3]
  float* gvec;             // [B,V,6] or nullptr
  double* gM_ws;           // [n_levels][B*V][12]  dL/d pose_mat per level
  unsigned* pose_counter;  // [B] zero between launches
  LevelDev lv[DVF_MAX_LEVELS];
};

// ---- cold, out-of-line exact versions --------------------------------------------------------
template <bool kZeros>
__device__ __noinline__ Proj project_exact(const float* P /*smem*/, Cam c, const Geo* g) {
  Proj o;
  project<true, kZeros>(P, c, *g, o);
  return o;
}
static __device__ __noinline__ ChainGrad chain_backward_exact(const float* P /*smem*/, Cam c, Proj p, Loc L, float gx, float gy,
                                                              const Geo* g) {
  ChainGrad o;
  chain_backward<true>(P, c, p, L, gx, gy, *g, o);
  return o;
}

// out-of-line so that their register / stack needs (sinf/cosf slow paths, fp64) stay out of the pixel loop
static __device__ __noinline__ float trig_of(float a, int want_sin) { return want_sin ? torch_sinf(a) : torch_cosf(a); }
static __device__ __noinline__ void level_gM(const float* K, float ds, const float* gP /*12, smem*/, double* dst) {
  float Ks[9];
  scaled_K(K, ds, Ks);
  double gM[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) gM[q] = 0.0;
  accumulate_gM(Ks, gP, gM);
#pragma unroll
  for (int q = 0; q < 12; ++q) __stcg(dst + q, gM[q]);
}
static __device__ __noinline__ double dtrig_of(double a, int want_cos) { return want_cos ? cos(a) : sin(a); }
// d(sum gR * Rx Ry Rz)/d angle `which` (0 = x, 1 = y, 2 = z), fp64; tr = (sx, cx, sy, cy, sz, cz)
static __device__ __noinline__ float euler_angle_grad(const double* gM /*3x4*/, const double* tr, int which) {
  const double sx = tr[0], cx = tr[1], sy = tr[2], cy = tr[3], sz = tr[4], cz = tr[5];
  double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  if (which == 0) { const double d[9] = {0, 0, 0, 0, -sx, -cx, 0, cx, -sx}; for (int q = 0; q < 9; ++q) Rx[q] = d[q]; }
  if (which == 1) { const double d[9] = {-sy, 0, cy, 0, 0, 0, -cy, 0, -sy}; for (int q = 0; q < 9; ++q) Ry[q] = d[q]; }
  if (which == 2) { const double d[9] = {-sz, -cz, 0, cz, -sz, 0, 0, 0, 0}; for (int q = 0; q < 9; ++q) Rz[q] = d[q]; }
  double t1[9], t2[9], gR[9];
  dmm3(Rx, Ry, t1);
  dmm3(t1, Rz, t2);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) gR[r * 3 + c] = gM[r * 4 + c];
  return (float)ddot9(gR, t2);
}

// CTA prologue: projection matrices of image b at this level into shared memory -- either read from the
// arrays the caller computed (dvf_pose_proj_fwd) or derived here from the 6-DoF vectors (same routines).
template <int kV>
__device__ __forceinline__ void load_matrices(const LossParams& prm, const LevelDev& lv, int b, float (*s_P)[12], float* s_M,
                                              bool new_image = true) {
  const int tid = threadIdx.x;
  if (prm.pose_vec) {
    // Three short stages, each spread over threads so that the prologue costs one global-load latency plus a
    // few hundred cycles: (1) fetch vec / K / K^-1, (2) sin / cos (torch-CPU order, dvf_pose.cuh) and the level's scaled
    // intrinsics, (3) every thread of the first kV*12 composes R and takes one entry of P = K_s @ [R|t].
    // Operation order per entry is identical to dvf_pose_proj_fwd.  new_image = false (CTA-uniform): the previous piece
    // of this CTA belonged to the same image, so vec, sin / cos, K and K^-1 are still in shared memory and only the
    // level-dependent part is redone (no global load, no trigonometry: the pieces of the small pyramid levels).
    __shared__ float s_vec[kV][6];
    __shared__ float s_trig[kV][6];
    __shared__ float s_K0[9], s_Ki0[9];
    __shared__ float s_Ks[9];
    const bool euler = prm.rotation == DVF_ROT_EULER;
    if (new_image) {
      if (tid < kV * 6) s_vec[tid / 6][tid % 6] = prm.pose_vec[((size_t)b * kV + tid / 6) * 6 + tid % 6];
      else if (tid >= 32 && tid < 41) s_K0[tid - 32] = prm.K[b * 9 + (tid - 32)];
      else if (tid >= 64 && tid < 73) s_Ki0[tid - 64] = prm.Kinv[b * 9 + (tid - 64)];
      __syncthreads();
    }
    if (tid < kV * 6) {
      if (euler && new_image) {
        const int v = tid / 6, q = tid % 6;           // q: 0 cz, 1 sz, 2 cy, 3 sy, 4 cx, 5 sx
        s_trig[v][q] = trig_of(s_vec[v][3 + (2 - q / 2)], q & 1);
      }
    } else if (tid >= 32 && tid < 41) {
      const int q = tid - 32;
      s_Ks[q] = (q < 6 && lv.ds != 1.0f) ? div(s_K0[q], lv.ds) : s_K0[q];               // rows 0-1 / downscale
    } else if (tid >= 64 && tid < 73) {
      const int q = tid - 64;
      s_M[q] = ((q % 3) < 2 && lv.ds != 1.0f) ? mul(s_Ki0[q], lv.ds) : s_Ki0[q];         // columns 0-1 * downscale
    }
    __syncthreads();
    if (tid < kV * 12) {
      const int v = tid / 12, r = (tid % 12) / 4, c = tid % 4;
      float R[9];
      if (euler) euler_compose(s_vec[v][5], s_trig[v][0], s_trig[v][1], s_trig[v][2], s_trig[v][3], s_trig[v][4], s_trig[v][5], R);
      else rotation_fwd(&s_vec[v][3], prm.rotation, R);
      const float p0 = c < 3 ? R[0 * 3 + (c < 3 ? c : 0)] : s_vec[v][0];
      const float p1 = c < 3 ? R[1 * 3 + (c < 3 ? c : 0)] : s_vec[v][1];
      const float p2 = c < 3 ? R[2 * 3 + (c < 3 ? c : 0)] : s_vec[v][2];
      s_P[v][r * 4 + c] = add(add(mul(s_Ks[r * 3 + 0], p0), mul(s_Ks[r * 3 + 1], p1)), mul(s_Ks[r * 3 + 2], p2));
    }
  } else {
    if (tid < kV * 12) s_P[tid / 12][tid % 12] = lv.P[((size_t)b * kV + tid / 12) * 12 + tid % 12];
    if (tid >= 64 && tid < 73) s_M[tid - 64] = lv.Kinv[b * 9 + (tid - 64)];
  }
}

// Programmatic dependent launch (DVF_FLAG_PDL): every CTA lets the NEXT kernel of the stream be scheduled as soon as
// SM slots free up, and runs WITHOUT waiting for the previous kernel: by contract (include/dvf_b200.h) that kernel
// shares no buffer with this launch, workspace included.  A CTA only waits for it right before it exits, so that
// kernels still complete in stream order (whoever waits for this grid has then waited for its predecessors too);
// by then the predecessor is long done.  A first version waited before the first workspace access instead (shared
// workspaces allowed): CTAs that had started early on freed SM slots then idled for up to 40 us (median 5 us) holding
// their slots -- profiles/trace_pieces.py.  Both instructions are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_let_successor_start(const LossParams& prm) {
  if (prm.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_predecessor(const LossParams& prm) {
  if (prm.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// depth from the network's disparity exactly as torch evaluates 1 / (disp + eps): a rounded add, then reciprocal()
// (python's 1 / tensor is tensor.reciprocal() * 1); backward of reciprocal: -g * (depth * depth)
__device__ __forceinline__ float depth_of_disp(float disp, float eps) { return div(1.0f, add(disp, eps)); }
__device__ __forceinline__ float gdisp_of_gdepth(float g, float depth) { return mul(-g, mul(depth, depth)); }

// sign(d)/N with sign(0) = sign(NaN) = 0, gated by `gate`
__device__ __forceinline__ float signed_unit(float d, float inv_n, bool gate) {
  const bool nz = (d < 0.0f || d > 0.0f) && gate;
  return nz ? copysignf(inv_n, d) : 0.0f;
}

// -sign(d) * n with sign(0) = sign(NaN) = 0 (n >= 0, already gated: 0 for pixels without a valid sample).  One LOP3 for
// the sign transfer (inverted sign bit of d onto |n|), one compare, one select.
__device__ __forceinline__ float neg_signed_unit(float d, float n) {
  const uint32_t r = (__float_as_uint(n) & 0x7fffffffu) | (~__float_as_uint(d) & 0x80000000u);
  return (d < 0.0f || d > 0.0f) ? __uint_as_float(r) : 0.0f;
}

// Shared tail of the loss kernels: CTA fold of acc[kV][16], partial write, ticket, image fold, level fold.
// `part` of `n_parts`: which of the CTA pieces of image b this is (fixed fold order => deterministic sums).
// Does not return early: the image kernel calls it once per piece of its unit range.
template <int kV, int kThreadsT>
__device__ __forceinline__ void reduce_and_finish(float (&acc)[kV][kRedSlots], const LossParams& prm, const LevelDev& lv,
                                                  int l, int part, int n_parts, int b, int C) {
  __shared__ float s_red[kThreadsT / 32][kV][kRedSlots];
  __shared__ int s_flag;
  __shared__ double s_term[kThreadsT / 32];
  __shared__ float s_gp[kV][12];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int v = 0; v < kV; ++v) {
    const float r = butterfly16(acc[v], lane);
    if ((lane & 1) == 0) s_red[warp][v][butterfly_slot(lane)] = r;
  }
  __syncthreads();
  if (n_parts == 1) {
    // the whole (image, level) was this CTA's: its sums are final -- no partials in global memory, no ticket
    if (tid < kV * kRedSlots) {
      const int v = tid / kRedSlots, sl = tid % kRedSlots;
      float t = 0.0f;
#pragma unroll
      for (int w8 = 0; w8 < kThreadsT / 32; ++w8) t += s_red[w8][v][sl];
      const double sum = (double)t;   // same value as the fp64 fold of a single partial
      if (sl < 12) {
        if (lv.gP) lv.gP[((size_t)b * kV + v) * 12 + sl] = (float)sum;
        s_gp[v][sl] = (float)sum;
      } else if (sl == 12) {
        lv.img_terms[(size_t)b * kV + v] = sum;
      }
    }
    if (tid == 0) s_flag = 1;
  } else {
    float* const img_part = lv.partials + (size_t)b * lv.slots_per_image * kV * kRedSlots;
    float* const my_part = img_part + (size_t)part * kV * kRedSlots;
    if (tid < kV * kRedSlots) {
      const int v = tid / kRedSlots, sl = tid % kRedSlots;
      float t = 0.0f;
#pragma unroll
      for (int w8 = 0; w8 < kThreadsT / 32; ++w8) t += s_red[w8][v][sl];
      __stcg(my_part + tid, t);
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();   // cumulative: publishes the partial stores of the whole CTA (ordered before it by the barrier)
      s_flag = (atomicAdd(lv.img_counter + b, 1u) == (unsigned)(n_parts - 1));
      if (s_flag) lv.img_counter[b] = 0u;   // last arrival: restore the counter for the next launch
    }
  }
  __syncthreads();
  if (s_flag) {   // CTA-uniform
    if (n_parts > 1) {
      // ---- last piece of image b: fixed-order fp64 fold of the partials ----------------
      float* const img_part = lv.partials + (size_t)b * lv.slots_per_image * kV * kRedSlots;
      __threadfence();
      for (int pair = tid >> 3; pair < kV * kRedSlots; pair += kThreadsT / 8) {
        const int v = pair / kRedSlots, sl = pair % kRedSlots;
        const double sum = group8_sum(img_part + v * kRedSlots + sl, n_parts, kV * kRedSlots, tid & 7);
        if ((tid & 7) == 0) {
          if (sl < 12) {
            if (lv.gP) lv.gP[((size_t)b * kV + v) * 12 + sl] = (float)sum;
            s_gp[v][sl] = (float)sum;
          } else if (sl == 12) {
            lv.img_terms[(size_t)b * kV + v] = sum;
          }
        }
      }
    }
    __threadfence();
    __syncthreads();
    if (prm.gvec) {
      // pose mode: dL/d pose_mat of this level = K_s^T @ dL/dP (fp64); the level that finishes LAST for image b
      // adds the levels in fixed order and runs the analytic backward of pose_vec2mat
      if (tid < kV) {
        level_gM(prm.K + b * 9, lv.ds, &s_gp[tid][0], prm.gM_ws + (((size_t)l * prm.B + b) * kV + tid) * 12);
        __threadfence();
      }
      __syncthreads();
      if (tid == 0) {
        const bool last_level = atomicAdd(prm.pose_counter + b, 1u) == (unsigned)(prm.n_levels - 1);
        if (last_level) prm.pose_counter[b] = 0u;
        s_flag = last_level;
      }
      __syncthreads();
      if (s_flag) {
        // this sits at the tail of the kernel: spread it over threads instead of one long fp64 chain per view
        __threadfence();
        __shared__ double s_gM[kV][12];
        __shared__ double s_dtrig[kV][6];
        const bool euler = prm.rotation == DVF_ROT_EULER;
        if (tid < kV * 12) {
          const int v = tid / 12, q = tid % 12;
          double s = 0.0;
          for (int ll = 0; ll < prm.n_levels; ++ll) s += __ldcg(prm.gM_ws + (((size_t)ll * prm.B + b) * kV + v) * 12 + q);
          s_gM[v][q] = s;
        } else if (euler && tid >= 64 && tid < 64 + kV * 6) {
          const int v = (tid - 64) / 6, q = (tid - 64) % 6;        // q: 0 sx, 1 cx, 2 sy, 3 cy, 4 sz, 5 cz
          s_dtrig[v][q] = dtrig_of((double)prm.pose_vec[((size_t)b * kV + v) * 6 + 3 + q / 2], q & 1);
        }
        __syncthreads();
        if (tid < kV * 6) {
          const int v = tid / 6, q = tid % 6;
          float* gv = prm.gvec + ((size_t)b * kV + v) * 6;
          if (q < 3) {
            gv[q] = (float)s_gM[v][q * 4 + 3];                      // translation: last column of dL/d pose_mat
          } else if (euler) {
            gv[q] = euler_angle_grad(&s_gM[v][0], &s_dtrig[v][0], q - 3);
          } else if (q == 3) {
            float g6[6];
            posemat_bwd(&s_gM[v][0], prm.pose_vec + ((size_t)b * kV + v) * 6, prm.rotation, g6);
            gv[3] = g6[3]; gv[4] = g6[4]; gv[5] = g6[5];
          }
        }
      }
      __syncthreads();
    }
    if (tid == 0) s_flag = (atomicAdd(lv.lvl_counter, 1u) == (unsigned)(prm.B - 1));
    __syncthreads();
    if (s_flag) {
      // ---- last image of the level: loss terms ---------------------------------------------
      __threadfence();
      for (int v = 0; v < kV; ++v) {
        double s = 0.0;
        for (int bb = tid; bb < prm.B; bb += kThreadsT) s += __ldcg(lv.img_terms + (size_t)bb * kV + v);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) s_term[warp] = s;
        __syncthreads();
        if (tid == 0) {
          double t = 0.0;
          for (int w8 = 0; w8 < kThreadsT / 32; ++w8) t += s_term[w8];
          const float term = (float)(t / ((double)prm.mean_batch * (double)C * (double)lv.HW));
          prm.terms[l * kV + v] = term;
          if (prm.nan_flags && term != term) atomicOr(prm.nan_flags, 1 << ((l * kV + v) & 31));
          // all-gather over NVLink: one remote store per peer, fire and forget (nobody waits for it here)
          for (int q = 0; q < prm.n_peers; ++q)
            asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(prm.peer_terms[q] + (size_t)prm.peer_rank * prm.n_levels * kV + l * kV + v),
                         "f"(term)
                         : "memory");
        }
        __syncthreads();
      }
      if (tid == 0) *lv.lvl_counter = 0u;
    }
  }
  __syncthreads();   // s_flag / s_red may be rewritten by the caller's next piece
}

// ================================================================================================
// C == 3 (images), packed: each thread owns TWO horizontally adjacent target pixels and runs the whole
// floating-point chain on (pixel A, pixel B) pairs with FFMA2 / FMUL2 / FADD2 (dvf_math2.cuh).
// Same results bit for bit as the scalar arithmetic of dvf_math.cuh (used by the other kernels); ~half the
// floating-point issue slots.  No scatter / target-gradient outputs here (those requests go through the
// generic kernel below).
//
// kTma: the streaming inputs (depth, 3 target planes, explainability weights) of a CTA's run of pixels
// are contiguous in memory; they are fetched by 1-D bulk asynchronous copies (TMA) into a kStages-deep
// shared-memory ring, several chunks ahead of the computation, and the source rows around each chunk are
// prefetched into L2.  This keeps tens of KB of HBM reads in flight per SM at zero register cost -- the
// kernel is otherwise limited by how many loads its warps can have outstanding.  Requires HW % 4 == 0 and
// 16-byte aligned tensors (checked on the host); otherwise the same kernel runs with plain loads.
// ================================================================================================
constexpr int kPlanUnit = 512;   // granularity (pixels) of the host's work split for the generic / NHWC kernels
constexpr int kUnitPx = 2 * kLossThreads;   // image kernel: one unit = one pixel pair per thread of the CTA
constexpr int kStages = 4;

// first unit of CTA c (c in [0, G]) = floor(c*T/G)
__device__ __forceinline__ int split_start(int c, const LossParams& prm) {
  return c * prm.split_q + (int)fastdiv((uint32_t)c * (uint32_t)prm.split_r, prm.div_grid);   // c*r < G*G < 2^31
}
// CTA that owns global unit g when CTA c owns [floor(c*T/G), floor((c+1)*T/G)): float estimate (exact to well below
// one CTA: g < 2^31, G <= 2^13), then fixed up against the exact integer bounds
__device__ __forceinline__ int cta_of_unit(int g, const LossParams& prm) {
  int c = (int)(((float)g + 0.5f) * prm.ctas_per_unit);
  c = min(max(c, 0), prm.grid - 1);
  while (c > 0 && split_start(c, prm) > g) --c;
  while (c + 1 < prm.grid && split_start(c + 1, prm) <= g) ++c;
  return c;
}

// Resident CTAs per SM the register allocation is sized for.  One or two views: 4 (128 registers; 5 CTAs = 96 registers
// spill: 75.0 us against 67.6 on C2; 3 CTAs = 168 registers: 74.2).  Three or four views keep too much state for 128
// registers (132 / 332 bytes of spills): 3 CTAs per SM without spills are faster there (V = 4: 241.5 -> 220.8 us).
#ifdef DVF_C3_MINBLOCKS   // experiment builds
constexpr int c3_min_blocks(int) { return DVF_C3_MINBLOCKS; }
#else
constexpr int c3_min_blocks(int views) { return views >= 3 ? 3 : 4; }
#endif
// kGlue: the producer glue (disparity -> depth, image scale; dvf_loss_desc) is compiled in.  A separate variant because the
// three CTA-uniform branches it needs cost the plain path 1.8 % when they are always present (profiles/r2_summary.md).
//
// kExt: the three streamed "target" planes hold an UPSTREAM GRADIENT d L / d warped instead of a target image, and nothing
// of the loss is evaluated: the kernel is then the backward of the materialised warp (dvf_inverse_warp_bwd without d img:
// inverse_warp.py:160-193 differentiated w.r.t. depth and P) with the ring, the balanced split and the deterministic dL/dP
// fold of the loss kernel.  V = 1, no masks, one level.
template <int kV, bool kZeros, bool kExpl, bool kGrad, bool kTma, bool kGlue = false, int kMinBlocks = c3_min_blocks(kV),
          bool kExt = false>
__global__ void __launch_bounds__(kLossThreads, kMinBlocks) photo_loss_c3x2_kernel(const __grid_constant__ LossParams prm) {
  constexpr int kC = 3;
  static_assert(!kExt || (kV == 1 && !kExpl && kGrad && !kGlue), "kExt: backward of the materialised warp");
#ifdef DVF_ACC_SMEM
  constexpr bool kAccSmem = kV == 1;
#else
  constexpr bool kAccSmem = false;
#endif
  constexpr int kPlanes = 1 + kC + (kExpl ? kV : 0);   // streamed planes per chunk
  // pixel pairs per thread between two ring hand-overs (= CTA barriers).  Measured on C2 with the balanced split:
  // 2 pairs x 4 stages 69.6 us, 3 x 3 69.4 us, 4 x 2 67.5 us; with the masks of up to two views 3 pairs x 2 stages.
#ifndef DVF_PAIRS_PER_CHUNK   // experiment builds (profiles/ab.sh) override these
#define DVF_PAIRS_PER_CHUNK 4
#endif
#ifndef DVF_PAIRS_PER_CHUNK_MASKS   // 5 or 6 planes (masks of one or two views).  C3 (V = 2 + masks): 1 -> 405 us, 2 -> 400, 3 -> 390
#define DVF_PAIRS_PER_CHUNK_MASKS 3
#endif
  constexpr int kPairsPerChunk = kPlanes <= 4 ? DVF_PAIRS_PER_CHUNK : (kPlanes <= 6 ? DVF_PAIRS_PER_CHUNK_MASKS : 1);
  constexpr int kChunk = kUnitPx * kPairsPerChunk;    // pixels per chunk
  // ring depth: as deep as 40 KB of static shared memory allow (4 CTAs per SM stay resident), at least 2
#ifndef DVF_RING_BYTES
#define DVF_RING_BYTES 38000   // 2 stages of 4 planes x 1024 px, or of 6 planes x 768 px
#endif
  constexpr int kStagesFit = DVF_RING_BYTES / (kPlanes * kChunk * 4);
  constexpr int kSt = kStagesFit >= kStages ? kStages : (kStagesFit < 2 ? 2 : kStagesFit);
  __shared__ __align__(16) float s_P[kV][12];
  __shared__ __align__(16) float s_M[12];
  __shared__ __align__(128) float s_ring[kTma ? kSt : 1][kTma ? kPlanes : 1][kTma ? kChunk : 4];
  __shared__ __align__(8) uint64_t s_full[kSt];

  const int tid = threadIdx.x;
  if (kTma && tid == 0) {
#pragma unroll
    for (int s = 0; s < kSt; ++s) mbar_init(&s_full[s], 1);
    mbar_fence_init();
  }

  // balanced split: this CTA owns units [w, w_end) of the launch-wide unit list; it walks them (image, level) by
  // (image, level) -- "pieces", each with its own matrices, accumulators and partial-sum slot
  int w = split_start((int)blockIdx.x, prm);
  const int w_end = split_start((int)blockIdx.x + 1, prm);
  int ring_k = 0;   // chunks this CTA has pushed through the ring so far (stage / parity bookkeeping)
  pdl_let_successor_start(prm);
  DVF_MARK(0);
  int piece_no = 0;
  int prev_b = -1;   // image whose pose / intrinsics sit in shared memory
  const bool disp_mode = kGlue && prm.disparity != 0, scale_imgs = kGlue && prm.img_scale != 1.0f;   // CTA-uniform

  while (w < w_end) {
  const int b = (int)fastdiv((uint32_t)w, prm.div_upi);
  int l = 0;
  while (l + 1 < prm.n_levels && w - b * prm.units_per_image_all >= prm.lv[l + 1].unit_base) ++l;
  const LevelDev& lv = prm.lv[l];
  const int real0 = b * prm.units_per_image_all + lv.unit_base + prm.piece_overhead;   // first pixel-carrying unit
  const int real1 = real0 + lv.units_per_image;
  const int piece_end = min(w_end, real1);
  const int k0 = max(w, real0) - real0, k1 = piece_end - real0;   // units [k0, k1) of image b at level l
  const bool piece_starts_here = w <= real0, piece_ends_here = piece_end == real1;
  w = piece_end;
  if (k1 <= k0) continue;   // only overhead units fell into my range (CTA-uniform)
  // which of the CTA pieces of (b, l) this is: only pieces that are cut by a CTA boundary need the search
  const int first_cta = piece_starts_here ? (int)blockIdx.x : cta_of_unit(real0, prm);
  const int last_cta = piece_ends_here ? (int)blockIdx.x : cta_of_unit(real1 - 1, prm);
  const int n_parts = last_cta - first_cta + 1;
  const int part = (int)blockIdx.x - first_cta;
  const int H = lv.H, W = lv.W, HW = lv.HW;
  const Geo geo = lv.geo;
  const Geo2 geo2 = make_geo2(geo);
  const FastDiv divW = lv.divW;
  const float inv_n = prm.upstream ? mul(lv.inv_n, __ldg(prm.upstream)) : lv.inv_n;

  // dL/dP partial sums, (A,B) lanes folded at the end.  kAccSmem: they live in thread-private shared memory
  // (column tid of s_acc) instead of 24 registers
  __shared__ f2 s_acc[kAccSmem ? 12 : 1][kAccSmem ? kLossThreads : 1];
  f2 acc2[kV][12];
#pragma unroll
  for (int v = 0; v < kV; ++v)
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      if (kAccSmem) s_acc[k][tid] = dup(0.0f);
      else acc2[v][k] = dup(0.0f);
    }
  float accl[kV];      // loss partial sums
#pragma unroll
  for (int v = 0; v < kV; ++v) accl[v] = 0.0f;

  const float* const depth_b = lv.depth + (size_t)b * HW;
  const float* const tgt0 = lv.tgt + (size_t)b * kC * HW;
  float* const gdepth_b = lv.gdepth + (size_t)b * HW;
  const bool want_gdepth = lv.gdepth != nullptr;
  const float* const expl_b = kExpl ? lv.expl + (size_t)b * lv.expl_bstride : nullptr;
  float* const gexpl_b = (kExpl && lv.gexpl) ? lv.gexpl + (size_t)b * kV * HW : nullptr;
  const float* src_b[kV];
#pragma unroll
  for (int v = 0; v < kV; ++v) src_b[v] = lv.src[v] + (size_t)b * kC * HW;

  // this piece: pixels [px_begin, px_end) of image b, walked in chunks of kChunk (the last one may be short)
  const int px_begin = k0 * kUnitPx;
  const int px_end = min(k1 * kUnitPx, HW);
  const int n_chunks = (px_end - px_begin + kChunk - 1) / kChunk;

  // producer (thread 0): bulk copies of chunk k into ring stage k % kStages + L2 prefetch of the source rows
  auto issue = [&](int k) {
    const int start = px_begin + k * kChunk;
    const uint32_t bytes = (uint32_t)(min(kChunk, px_end - start) * 4);   // multiple of 16 (HW % 4 == 0)
    const int st = (ring_k + k) % kSt;
    mbar_expect_tx(&s_full[st], bytes * kPlanes);
    bulk_g2s(&s_ring[st][0][0], depth_b + start, bytes, &s_full[st]);
#pragma unroll
    for (int c = 0; c < kC; ++c) bulk_g2s(&s_ring[st][1 + c][0], tgt0 + c * HW + start, bytes, &s_full[st]);
    if (kExpl) {
#pragma unroll
      for (int v = 0; v < kV; ++v) bulk_g2s(&s_ring[st][1 + kC + v][0], expl_b + v * HW + start, bytes, &s_full[st]);
    }
    // the bilinear taps of these pixels lie in the same rows of the source image give or take the parallax:
    // pull that band towards L2 so that the gathers find it there
    // Only the part of the band that the previous chunk of this piece did not cover (every source texel is requested once),
    // +- 1 row of parallax.  Measured (C2 / C3 / C5loss, us per step): whole band +- 2 rows for every chunk 46.6 / 389 / 455,
    // +- 4 rows 47.5 / - / 578, leading edge +- 3 rows 46.1 / 386 / 425, +- 2 rows 45.8 / 381 / 420, +- 1 row 45.7 / 374 / 414,
    // no prefetch at all 51.1 / 397 / 406: the requests themselves are not free, and for rows wider than 512 pixels (level 0
    // of the 256 x 832 shape) they cost more than they bring -- lv.prefetch_rows is 0 there.
    const int pf = lv.prefetch_rows;
    int lo = (k == 0 ? start - pf * W : start + pf * W), hi = start + kChunk + pf * W;
    lo = max(lo, 0) & ~3;
    hi = pf > 0 ? (min(hi, HW) & ~3) : lo;
    if (hi > lo) {
#pragma unroll
      for (int v = 0; v < kV; ++v) {
        const float* s0 = lv.src[v] + (size_t)b * kC * HW;
#pragma unroll
        for (int c = 0; c < kC; ++c) bulk_prefetch_l2(s0 + c * HW + lo, (uint32_t)(hi - lo) * 4);
      }
    }
  };
  if (kTma && tid == 0) {
    for (int k = 0; k < min(kSt, n_chunks); ++k) issue(k);
  }

  // projection matrices of this image (overlaps with the bulk copies just issued)
  if (piece_no == 0) DVF_MARK(1);
  load_matrices<kV>(prm, lv, b, s_P, s_M, b != prev_b);
  prev_b = b;
  __syncthreads();
  if (piece_no == 0) DVF_MARK(2);
  bool mats_ok = lv.allow_fast != 0;
  float M[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    M[k] = s_M[k];
    mats_ok = mats_ok && (fabsf(M[k]) <= 1048576.0f);
  }
#pragma unroll
  for (int v = 0; v < kV; ++v)
#pragma unroll
    for (int k = 0; k < 12; ++k) mats_ok = mats_ok && (fabsf(s_P[v][k]) <= 1073741824.0f);
  // with |P| <= 2^30, |Kinv| <= 2^20, |depth| <= 2^30 and pixel indices < 2^15 every intermediate of the coordinate
  // chain stays below 2^100 => the shared-reciprocal divisions are exact; anything else takes the cold exact path
  const float depth_max = mats_ok ? 1073741824.0f : -1.0f;


  // one chunk; kTail = the chunk may contain lanes past the end of the run (only the last chunk can)
  auto do_chunk = [&](int k, auto tail_tag) {
    constexpr bool kTail = decltype(tail_tag)::value;
    const int st = (ring_k + k) % kSt;
    if (kTma) mbar_wait(&s_full[st], (uint32_t)((ring_k + k) / kSt) & 1u);
    // a short last chunk (the unit granularity is one pair per thread) has fewer pairs
    const int n_pairs = kPairsPerChunk == 1 ? 1 : min(kPairsPerChunk, (px_end - (px_begin + k * kChunk) + kUnitPx - 1) / kUnitPx);
#pragma unroll 1
    for (int h = 0; h < n_pairs; ++h) {
    const int slot = h * kUnitPx + 2 * tid;          // my pair inside the chunk
    const int idxA = px_begin + k * kChunk + slot;
    const int idxB = idxA + 1;
    const bool liveA = !kTail || idxA < px_end, liveB = !kTail || idxB < px_end;
    f2 dep, tg0, tg1, tg2;
    f2 exv[kExpl ? kV : 1];
    if (kTma) {
      dep = *reinterpret_cast<const f2*>(&s_ring[st][0][slot]);
      tg0 = *reinterpret_cast<const f2*>(&s_ring[st][1][slot]);
      tg1 = *reinterpret_cast<const f2*>(&s_ring[st][2][slot]);
      tg2 = *reinterpret_cast<const f2*>(&s_ring[st][3][slot]);
      if (kExpl) {
#pragma unroll
        for (int v = 0; v < kV; ++v) exv[v] = *reinterpret_cast<const f2*>(&s_ring[st][1 + kC + v][slot]);
      }
    } else {
      const int ia = kTail ? min(idxA, HW - 1) : idxA, ib = kTail ? min(idxB, HW - 1) : idxB;
      dep = make_float2(ld_stream(depth_b + ia), ld_stream(depth_b + ib));
      tg0 = make_float2(ld_stream(tgt0 + ia), ld_stream(tgt0 + ib));
      tg1 = make_float2(ld_stream(tgt0 + HW + ia), ld_stream(tgt0 + HW + ib));
      tg2 = make_float2(ld_stream(tgt0 + 2 * HW + ia), ld_stream(tgt0 + 2 * HW + ib));
      if (kExpl) {
#pragma unroll
        for (int v = 0; v < kV; ++v) exv[v] = make_float2(ld_stream(expl_b + v * HW + ia), ld_stream(expl_b + v * HW + ib));
      }
    }
    if (disp_mode) dep = make_float2(depth_of_disp(dep.x, prm.disp_eps), depth_of_disp(dep.y, prm.disp_eps));
    if (scale_imgs) {
      tg0 = mul2(tg0, dup(prm.img_scale));
      tg1 = mul2(tg1, dup(prm.img_scale));
      tg2 = mul2(tg2, dup(prm.img_scale));
    }
    // dead lanes (past the end of the run) may read stale ring contents: give them a harmless depth; their
    // taps are never loaded (all-zero => invalid => zero gradients) and nothing of theirs is stored
    if (kTail) {
      dep = make_float2(liveA ? dep.x : 1.0f, liveB ? dep.y : 1.0f);
      // ... and finite target / weight values: stale shared memory may hold NaNs of an earlier kernel, and 0 * NaN
      // would reach the dL/dP sums through the (zero) gradient of a dead lane
      tg0 = make_float2(liveA ? tg0.x : 0.0f, liveB ? tg0.y : 0.0f);
      tg1 = make_float2(liveA ? tg1.x : 0.0f, liveB ? tg1.y : 0.0f);
      tg2 = make_float2(liveA ? tg2.x : 0.0f, liveB ? tg2.y : 0.0f);
      if (kExpl) {
#pragma unroll
        for (int v = 0; v < kV; ++v) exv[v] = make_float2(liveA ? exv[v].x : 1.0f, liveB ? exv[v].y : 1.0f);
      }
    }
    Cam2 cam;
    {
      const int ca = kTail ? min(idxA, HW - 1) : idxA, cb = kTail ? min(idxB, HW - 1) : idxB;
      const int iA = (int)fastdiv((uint32_t)ca, divW), jA = ca - iA * W;
      const int iB = (int)fastdiv((uint32_t)cb, divW), jB = cb - iB * W;
#ifdef DVF_M_SMEM
      float Mv[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) Mv[q] = *reinterpret_cast<volatile float*>(&s_M[q]);
      pixel_to_cam2(Mv, dep, make_float2((float)iA, (float)iB), make_float2((float)jA, (float)jB), cam);
#else
      pixel_to_cam2(M, dep, make_float2((float)iA, (float)iB), make_float2((float)jA, (float)jB), cam);
#endif
    }
    const bool fastA = fabsf(dep.x) <= depth_max, fastB = fabsf(dep.y) <= depth_max;   // false for NaN
    const bool slow = !(fastA && fastB);
    f2 gd = dup(0.0f);

#pragma unroll
    for (int v = 0; v < kV; ++v) {
      float P[12];
#pragma unroll
      for (int q = 0; q < 12; ++q) P[q] = s_P[v][q];
      const float* const src0 = src_b[v];

      Proj2 pr;
      Loc2 L;
      project2<kZeros>(P, cam, geo2, pr);
      if (__builtin_expect(slow, 0)) {   // redo the offending lane(s) with exact divisions
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 0 ? fastA : fastB) continue;
          Cam c1;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            c1.ray[q] = h == 0 ? cam.ray[q].x : cam.ray[q].y;
            c1.cam[q] = h == 0 ? cam.cam[q].x : cam.cam[q].y;
          }
          const Proj e = project_exact<kZeros>(&s_P[v][0], c1, &lv.geo);
          if (h == 0) {
            pr.qz.x = e.qz; pr.nZ.x = -e.Z; pr.u.x = e.u; pr.v.x = e.v; pr.xn.x = e.xn; pr.yn.x = e.yn;
          } else {
            pr.qz.y = e.qz; pr.nZ.y = -e.Z; pr.u.y = e.u; pr.v.y = e.v; pr.xn.y = e.xn; pr.yn.y = e.yn;
          }
        }
      }
      locate2<kZeros>(pr.xn, pr.yn, H, W, geo, geo2, L);

      const int oA = L.y0A * W + L.x0A, oB = L.y0B * W + L.x0B;
      const bool nwA = L.nwA && liveA, neA = L.neA && liveA, swA = L.swA && liveA, seA = L.seA && liveA;
      const bool nwB = L.nwB && liveB, neB = L.neB && liveB, swB = L.swB && liveB, seB = L.seB && liveB;
      f2 t00[kC], t01[kC], t10[kC], t11[kC];
      {
        const float* pa = ptr_off(src0, oA);
        const float* pb = ptr_off(src0, oB);
#pragma unroll
        for (int c = 0; c < kC; ++c) {
          const float* pa1 = ptr_off(pa, W);
          const float* pb1 = ptr_off(pb, W);
          t00[c] = make_float2(nwA ? __ldg(pa) : 0.0f, nwB ? __ldg(pb) : 0.0f);
          t01[c] = make_float2(neA ? __ldg(pa + 1) : 0.0f, neB ? __ldg(pb + 1) : 0.0f);
          t10[c] = make_float2(swA ? __ldg(pa1) : 0.0f, swB ? __ldg(pb1) : 0.0f);
          t11[c] = make_float2(seA ? __ldg(pa1 + 1) : 0.0f, seB ? __ldg(pb1 + 1) : 0.0f);
          if (c + 1 < kC) {
            pa = ptr_off(pa, HW);
            pb = ptr_off(pb, HW);
          }
        }
      }
      if (scale_imgs) {
#pragma unroll
        for (int c = 0; c < kC; ++c) {
          t00[c] = mul2(t00[c], dup(prm.img_scale));
          t01[c] = mul2(t01[c], dup(prm.img_scale));
          t10[c] = mul2(t10[c], dup(prm.img_scale));
          t11[c] = mul2(t11[c], dup(prm.img_scale));
        }
      }
      const f2 ex = kExpl ? exv[kExpl ? v : 0] : dup(1.0f);

      f2 d0[kC], d1[kC];
      bool anyA = true, anyB = true;
      if (!kExt) {
        const f2 wnw = mul2(L.s, L.e), wne = mul2(L.s, L.w), wsw = mul2(L.n, L.e), wse = mul2(L.n, L.w);
        const f2 w0 = bilerp2(t00[0], t01[0], t10[0], t11[0], wnw, wne, wsw, wse);
        const f2 w1 = bilerp2(t00[1], t01[1], t10[1], t11[1], wnw, wne, wsw, wse);
        const f2 w2 = bilerp2(t00[2], t01[2], t10[2], t11[2], wnw, wne, wsw, wse);
        // dead lanes have all-zero taps => any = false
        anyA = (w0.x != 0.0f) || (w1.x != 0.0f) || (w2.x != 0.0f);
        anyB = (w0.y != 0.0f) || (w1.y != 0.0f) || (w2.y != 0.0f);
        d0[0] = sub2(tg0, w0);
        d0[1] = sub2(tg1, w1);
        d0[2] = sub2(tg2, w2);
#pragma unroll
        for (int c = 0; c < kC; ++c) d1[c] = kExpl ? mul2(d0[c], ex) : d0[c];
        const float lsA = add(add(fabsf(d1[0].x), fabsf(d1[1].x)), fabsf(d1[2].x));
        const float lsB = add(add(fabsf(d1[0].y), fabsf(d1[1].y)), fabsf(d1[2].y));
        accl[v] += (anyA ? lsA : 0.0f) + (anyB ? lsB : 0.0f);
      }
      if (kGrad) {
        f2 gx = dup(0.0f), gy = dup(0.0f);
        // nsu = -sign(d1)/N (0 where d1 == 0 or the pixel has no valid sample): the gradient of the warped value
        const f2 ngate = make_float2(anyA ? inv_n : 0.0f, anyB ? inv_n : 0.0f);
        f2 nsu[kC];
#pragma unroll
        for (int c = 0; c < kC; ++c) {
          if (kExt) {   // the gradient of the warped value comes from the caller (dead lanes: zeroed above)
            bilerp_grad2(t00[c], t01[c], t10[c], t11[c], L, c == 0 ? tg0 : (c == 1 ? tg1 : tg2), gx, gy);
            continue;
          }
          nsu[c] = make_float2(neg_signed_unit(d1[c].x, ngate.x), neg_signed_unit(d1[c].y, ngate.y));
          const f2 ng = kExpl ? mul2(nsu[c], ex) : nsu[c];
          bilerp_grad2(t00[c], t01[c], t10[c], t11[c], L, ng, gx, gy);
        }
        if (kExpl) {
          if (gexpl_b) {   // d/d mask = sum_c sign(d1) * d0 / N = -(nsu . d0), summed left to right like the reference
            if (liveA)
              st_stream(gexpl_b + v * HW + idxA, -add(add(mul(nsu[0].x, d0[0].x), mul(nsu[1].x, d0[1].x)), mul(nsu[2].x, d0[2].x)));
            if (liveB)
              st_stream(gexpl_b + v * HW + idxB, -add(add(mul(nsu[0].y, d0[0].y), mul(nsu[1].y, d0[1].y)), mul(nsu[2].y, d0[2].y)));
          }
        }
        ChainGrad2 cg;
        chain_backward2<kZeros>(P, cam, pr, L, gx, gy, geo2, cg);
        if (__builtin_expect(slow, 0)) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 0 ? fastA : fastB) continue;
            Cam c1;
            Proj p1;
            Loc L1;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              c1.ray[q] = h == 0 ? cam.ray[q].x : cam.ray[q].y;
              c1.cam[q] = h == 0 ? cam.cam[q].x : cam.cam[q].y;
            }
            p1.qz = h == 0 ? pr.qz.x : pr.qz.y;
            p1.Z = h == 0 ? -pr.nZ.x : -pr.nZ.y;
            p1.rZ = 0.0f;
            p1.u = h == 0 ? pr.u.x : pr.u.y;
            p1.v = h == 0 ? pr.v.x : pr.v.y;
            p1.xn = h == 0 ? pr.xn.x : pr.xn.y;
            p1.yn = h == 0 ? pr.yn.x : pr.yn.y;
            p1.mx = kZeros && (h == 0 ? pr.xn.x : pr.xn.y) == 2.0f;
            p1.my = kZeros && (h == 0 ? pr.yn.x : pr.yn.y) == 2.0f;
            L1.x0 = L1.y0 = 0;
            L1.w = L1.e = L1.n = L1.s = 0.0f;
            L1.bnw = L1.bne = L1.bsw = L1.bse = false;
            L1.gmx = h == 0 ? L.gmx.x : L.gmx.y;
            L1.gmy = h == 0 ? L.gmy.x : L.gmy.y;
            const ChainGrad e = chain_backward_exact(&s_P[v][0], c1, p1, L1, h == 0 ? gx.x : gx.y, h == 0 ? gy.x : gy.y, &lv.geo);
            if (h == 0) {
              cg.gq[0].x = e.gq[0]; cg.gq[1].x = e.gq[1]; cg.gq[2].x = e.gq[2]; cg.gdepth.x = e.gdepth;
            } else {
              cg.gq[0].y = e.gq[0]; cg.gq[1].y = e.gq[1]; cg.gq[2].y = e.gq[2]; cg.gdepth.y = e.gdepth;
            }
          }
        }
        gd = make_float2(add(gd.x, cg.gdepth.x), add(gd.y, cg.gdepth.y));
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          if (kAccSmem) {
#pragma unroll
            for (int q = 0; q < 3; ++q) s_acc[r * 4 + q][tid] = fma2(cg.gq[r], cam.cam[q], s_acc[r * 4 + q][tid]);
            s_acc[r * 4 + 3][tid] = add2(s_acc[r * 4 + 3][tid], cg.gq[r]);
            continue;
          }
#pragma unroll
          // With several views the (A,B) halves are folded right away: 12 instead of 24 accumulator registers per
          // view (two scalar FMAs occupy the FMA pipe as long as one packed one).  Measured on C2: V=2 spills 276 B ->
          // 0 and 120.2 -> 114.9 us, V=2 with masks 136.1 -> 127.0 us, V=4 256.7 -> 241.5 us.
          if (kV > 1) {   // (no gain at V = 1: 67.7 vs 67.4 us)
#pragma unroll
            for (int q = 0; q < 3; ++q)
              acc2[v][r * 4 + q].x = fmaf(cg.gq[r].y, cam.cam[q].y, fmaf(cg.gq[r].x, cam.cam[q].x, acc2[v][r * 4 + q].x));
            acc2[v][r * 4 + 3].x += cg.gq[r].x + cg.gq[r].y;
            continue;
          }
#pragma unroll
          for (int q = 0; q < 3; ++q) acc2[v][r * 4 + q] = fma2(cg.gq[r], cam.cam[q], acc2[v][r * 4 + q]);
          acc2[v][r * 4 + 3] = add2(acc2[v][r * 4 + 3], cg.gq[r]);
        }
      }
    }  // views
    if (kGrad && want_gdepth) {
      if (disp_mode) {   // d/d disparity; the depth is recomputed from the staged disparity (two registers less to keep alive)
        const f2 dsp = kTma ? *reinterpret_cast<const f2*>(&s_ring[st][0][slot])
                            : make_float2(ld_stream(depth_b + (kTail ? min(idxA, HW - 1) : idxA)),
                                          ld_stream(depth_b + (kTail ? min(idxB, HW - 1) : idxB)));
        gd = make_float2(gdisp_of_gdepth(gd.x, depth_of_disp(dsp.x, prm.disp_eps)),
                         gdisp_of_gdepth(gd.y, depth_of_disp(dsp.y, prm.disp_eps)));
      }
      float* const gp = ptr_off(gdepth_b, idxA);
      if (liveA) st_stream(gp, gd.x);
      if (liveB) st_stream(gp + 1, gd.y);
    }
    }  // pairs of this chunk
    if (kTma) {
      __syncthreads();   // every thread has consumed ring stage k % kStages
      if (tid == 0 && k + kSt < n_chunks) issue(k + kSt);
    }
  };
  if (kTma && piece_no == 0) { mbar_wait(&s_full[ring_k % kSt], (uint32_t)(ring_k / kSt) & 1u); DVF_MARK(3); }
  for (int k = 0; k + 1 < n_chunks; ++k) do_chunk(k, std::false_type{});
  if (n_chunks > 0) {
    if ((px_end - px_begin) % kUnitPx == 0) do_chunk(n_chunks - 1, std::false_type{});
    else do_chunk(n_chunks - 1, std::true_type{});   // ragged end of the image: lanes past it are predicated off
  }
  ring_k += n_chunks;

  float acc[kV][kRedSlots];
#pragma unroll
  for (int v = 0; v < kV; ++v) {
#pragma unroll
    for (int q = 0; q < 12; ++q) acc[v][q] = kAccSmem ? s_acc[q][tid].x + s_acc[q][tid].y : acc2[v][q].x + acc2[v][q].y;
    acc[v][12] = accl[v];
    acc[v][13] = acc[v][14] = acc[v][15] = 0.0f;
  }
  if (piece_no == 0) DVF_MARK(4);
  if (piece_no == 0) DVF_MARK(5);
  reduce_and_finish<kV, kLossThreads>(acc, prm, lv, l, part, n_parts, b, kC);
  if (piece_no == 0) {
    DVF_MARK(6);
#ifdef DVF_TRACE
    if (tid == 0 && blockIdx.x < 8192) g_trace[blockIdx.x * 8 + 7] = (unsigned long long)(px_end - px_begin);
#endif
  }
  ++piece_no;
  }  // pieces of this CTA
  pdl_wait_predecessor(prm);
}

// ================================================================================================
// any channel count (feature maps in NCHW): channels are walked in a loop, twice when gradients
// to the maps are requested (pass 1 decides the value-based mask, pass 2 scatters)
// ================================================================================================
template <int kV, bool kZeros>
__global__ void __launch_bounds__(kLossThreads, 4) photo_loss_cn_kernel(const __grid_constant__ LossParams prm) {
  __shared__ __align__(16) float s_P[kV][12];
  __shared__ __align__(16) float s_M[12];

  int l = 0;
  while (l + 1 < prm.n_levels && (int)blockIdx.x >= prm.lv[l + 1].block_begin) ++l;
  const LevelDev& lv = prm.lv[l];
  const int rel = (int)blockIdx.x - lv.block_begin;
  const int b = rel / lv.blocks_per_image;
  const int chunk = rel - b * lv.blocks_per_image;
  const int tid = threadIdx.x;
  const int H = lv.H, W = lv.W, HW = lv.HW, C = prm.C;
  const Geo geo = lv.geo;
  const bool need_grad = prm.need_grad != 0;
  const bool allow_fast = lv.allow_fast != 0;
  const bool has_expl = lv.expl != nullptr;
  const float inv_n = prm.upstream ? mul(lv.inv_n, __ldg(prm.upstream)) : lv.inv_n;

  load_matrices<kV>(prm, lv, b, s_P, s_M);
  __syncthreads();

  float acc[kV][kRedSlots];
#pragma unroll
  for (int v = 0; v < kV; ++v)
#pragma unroll
    for (int k = 0; k < kRedSlots; ++k) acc[v][k] = 0.0f;

  const size_t img_off = (size_t)b * C * HW;
  const float* depth_b = lv.depth + (size_t)b * HW;
  const float* tgt_b = lv.tgt + img_off;
  float* gtgt_b = lv.gtgt ? lv.gtgt + img_off : nullptr;
  const int px_begin = chunk * lv.px_per_cta + tid;

  for (int q = 0; q < lv.px_per_cta / kLossThreads; ++q) {
    const int idx = px_begin + q * kLossThreads;
    const bool live = idx < HW;
    Cam cam;
    {
      float M[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) M[k] = s_M[k];
      const int i = (int)fastdiv((uint32_t)idx, lv.divW);
      float dv = live ? ld_stream(depth_b + idx) : 1.0f;
      if (prm.disparity) dv = depth_of_disp(dv, prm.disp_eps);
      pixel_to_cam(M, dv, i, idx - i * W, cam);
    }
    float gd = 0.0f;
#pragma unroll
    for (int v = 0; v < kV; ++v) {
      float P[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) P[k] = s_P[v][k];
      const float* src_b = lv.src[v] + img_off;
      float* gsrc_b = lv.gsrc[v] ? lv.gsrc[v] + img_off : nullptr;
      Proj pr;
      Loc L;
      const bool fast = project<false, kZeros>(P, cam, geo, pr) && allow_fast;
      if (__builtin_expect(!fast, 0)) pr = project_exact<kZeros>(&s_P[v][0], cam, &lv.geo);
      locate<kZeros>(pr.xn, pr.yn, H, W, geo, L);
      const bool bnw = L.bnw && live, bne = L.bne && live, bsw = L.bsw && live, bse = L.bse && live;
      const int o_nw = L.y0 * W + L.x0;
      const float wnw = mul(L.s, L.e), wne = mul(L.s, L.w), wsw = mul(L.n, L.e), wse = mul(L.n, L.w);
      const float ex = (has_expl && live) ? ld_stream(lv.expl + (size_t)b * lv.expl_bstride + (size_t)v * HW + idx) : 1.0f;
      float gx = 0.0f, gy = 0.0f, ge = 0.0f, lsum = 0.0f;
      bool any = false;
      // pass 1: mask, loss, d/d(ix,iy) assuming the pixel is valid
#pragma unroll 2
      for (int c = 0; c < C; ++c) {
        const float* r0 = src_b + (size_t)c * HW + o_nw;
        const float a0 = bnw ? __ldg(r0) : 0.0f, a1 = bne ? __ldg(r0 + 1) : 0.0f;
        const float a2 = bsw ? __ldg(r0 + W) : 0.0f, a3 = bse ? __ldg(r0 + W + 1) : 0.0f;
        const float t = live ? ld_stream(tgt_b + (size_t)c * HW + idx) : 0.0f;
        const float wv = bilerp(a0, a1, a2, a3, wnw, wne, wsw, wse);
        any |= (wv != 0.0f);
        const float d0 = sub(t, wv);
        const float d1 = has_expl ? mul(d0, ex) : d0;
        lsum = add(lsum, fabsf(d1));
        const float gd1 = signed_unit(d1, inv_n, true);
        const float g = has_expl ? mul(gd1, ex) : gd1;
        ge = add(ge, mul(gd1, d0));
        bilerp_grad(a0, a1, a2, a3, L, -g, gx, gy);
      }
      if (!any) {
        lsum = 0.0f;
        ge = 0.0f;
        gx = 0.0f;
        gy = 0.0f;
      }
      acc[v][12] += lsum;
      if (need_grad) {
        // pass 2: gradients to the target map and scatter to the source map
        if (live && (gsrc_b || gtgt_b)) {
          for (int c = 0; c < C; ++c) {
            float g = 0.0f;
            if (any) {
              const float* r0 = src_b + (size_t)c * HW + o_nw;
              const float a0 = bnw ? __ldg(r0) : 0.0f, a1 = bne ? __ldg(r0 + 1) : 0.0f;
              const float a2 = bsw ? __ldg(r0 + W) : 0.0f, a3 = bse ? __ldg(r0 + W + 1) : 0.0f;
              const float wv = bilerp(a0, a1, a2, a3, wnw, wne, wsw, wse);
              const float d0 = sub(ld_stream(tgt_b + (size_t)c * HW + idx), wv);
              const float d1 = has_expl ? mul(d0, ex) : d0;
              const float gd1 = signed_unit(d1, inv_n, true);
              g = has_expl ? mul(gd1, ex) : gd1;
            }
            if (gtgt_b) {   // the same thread visits the views in order: plain read-modify-write
              float* qg = gtgt_b + (size_t)c * HW + idx;
              *qg = (v == 0) ? g : add(*qg, g);
            }
            if (gsrc_b && any) {
              float* gp = gsrc_b + (size_t)c * HW + o_nw;
              if (bnw) atomicAdd(gp, mul(wnw, -g));
              if (bne) atomicAdd(gp + 1, mul(wne, -g));
              if (bsw) atomicAdd(gp + W, mul(wsw, -g));
              if (bse) atomicAdd(gp + W + 1, mul(wse, -g));
            }
          }
        }
        if (lv.gexpl && live) st_stream(lv.gexpl + ((size_t)b * kV + v) * HW + idx, ge);
        ChainGrad cg;
        chain_backward<false>(P, cam, pr, L, gx, gy, geo, cg);
        if (__builtin_expect(!fast, 0)) cg = chain_backward_exact(&s_P[v][0], cam, pr, L, gx, gy, &lv.geo);
        if (live) {
          gd = add(gd, cg.gdepth);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[v][r * 4 + k] = fmaf(cg.gq[r], cam.cam[k], acc[v][r * 4 + k]);
            acc[v][r * 4 + 3] += cg.gq[r];
          }
        }
      }
    }
    if (need_grad && live && lv.gdepth) {
      if (prm.disparity) gd = gdisp_of_gdepth(gd, depth_of_disp(ld_stream(depth_b + idx), prm.disp_eps));
      st_stream(lv.gdepth + (size_t)b * HW + idx, gd);
    }
  }
  reduce_and_finish<kV, kLossThreads>(acc, prm, lv, l, chunk, lv.blocks_per_image, b, C);
}

// Launch of the image kernel: one CTA per resident slot of this variant (occupancy query, cached per variant), every
// CTA an equal share of the units => a single wave without a tail.  dvf_loss_desc.ctas_per_sm overrides (tuning aid).
template <void (*kKernel)(const LossParams)>
inline void launch_balanced(const LossParams& prm_in, int cap, cudaStream_t st) {
  LossParams prm = prm_in;
  static int per_sm = 0;   // benign race: every thread computes the same value
  if (per_sm == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kKernel, kLossThreads, 0) != cudaSuccess || n < 1) n = 1;
    per_sm = n;
  }
  int use = prm.ctas_per_sm > 0 ? prm.ctas_per_sm : per_sm;
  if (prm.pdl > 1 && prm.ctas_per_sm <= 0) {
    // DVF_FLAG_PDL_CHAINED: consecutive launches overlap, so ONE launch need not fill every SM slot -- the slots it
    // leaves free are taken by its neighbours in the stream.  Fewer, larger CTAs mean fewer pieces (every CTA boundary cuts
    // one) and less fixed cost: about 64 units per CTA.  Measured, C2 (19 k units): 4 / 3 / 2 / 1 CTAs per SM -> 51.4 / 49.4 /
    // 48.7 / 48.8 us per step; a C3 shard of 32 images (9 k units): 58.0 / 56.1 / 54.9 / 54.3; C3 with 256 images: flat.
    const long long want = ((long long)prm.total_units + 64ll * num_sms() - 1) / (64ll * num_sms());
    use = (int)(want < 1 ? 1 : (want > per_sm ? per_sm : want));
  }
  long long g = (long long)use * num_sms();
  if (prm.pdl > 1 && prm.ctas_per_sm == 0) {
    // ... and a grid that is a multiple of the batch puts every CTA boundary on an image boundary or inside ONE image: no
    // CTA walks two images (the pose / intrinsics prologue runs once per CTA) and fewer pieces are cut.  Only while the grid
    // stays well above the SM count (C2: 296 -> 256 CTAs 48.7 -> 48.4 us; C3, 256 images: 592 -> 512 CTAs 415 -> 402 us; a
    // 32-image shard would drop to 128 CTAs and lose 7 %: left alone).
    const long long k = g / prm.B;
    if (k >= 1 && k * prm.B >= (3ll * num_sms()) / 2) g = k * prm.B;
  }
  if (prm.ctas_per_sm < 0) g = -(long long)prm.ctas_per_sm;   // tuning: explicit grid
  if (g > cap) g = cap;
  if (g > prm.total_units) g = prm.total_units;
  prm.grid = (int)g;
  prm.split_q = prm.total_units / (int)g;
  prm.split_r = prm.total_units % (int)g;
  prm.ctas_per_unit = (float)g / (float)prm.total_units;
  prm.div_grid = make_fastdiv((uint32_t)g);
  prm.div_upi = make_fastdiv((uint32_t)prm.units_per_image_all);
  if (prm.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)g);
    cfg.blockDim = dim3(kLossThreads);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kKernel, prm);
    return;
  }
  kKernel<<<(int)g, kLossThreads, 0, st>>>(prm);
}

template <int kV, bool kZeros>
void launch_loss_c3(const LossParams& prm, int blocks, bool expl, bool grad, bool tma, cudaStream_t st);
template <int kV, bool kZeros>
void launch_loss_cn(const LossParams& prm, int blocks, cudaStream_t st);
void launch_warp_bwd_fused(const LossParams& prm, int blocks, bool zeros, cudaStream_t st);   // kExt variants (TMA ring only)

}  // namespace dvf
