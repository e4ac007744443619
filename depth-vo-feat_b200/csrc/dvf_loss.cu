// dvf_loss.cu -- fused masked reconstruction loss, forward + backward in ONE pass over HBM.
//
// Replaces photometric_reconstruction_loss of pytorch_version/loss_functions.py:7-20,
// loss_functions_sfm.py:9-46 and loss_function_sfm_old.py:7-46 (which call inverse_warp,
// inverse_warp.py:160-193, once per view and scale and then run ~60 elementwise passes).
//
// One launch covers every pyramid level and every source view of a loss call:
//   * a CTA owns a contiguous run of target pixels of ONE image of ONE level, so the
//     per-image projection P and K^-1 are CTA-uniform;
//   * each thread walks kPPT pixels (stride = CTA width => every depth / target load
//     and every depth-gradient store is a fully coalesced 128 B line per warp);
//   * per pixel the depth / target values are read once and shared by all V views; the
//     4 bilinear taps per channel are plain read-only gathers: neighbouring lanes hit
//     neighbouring texels, L1 absorbs the x0/x1 and y0/y1 reuse;
//   * loss term and the 12 entries of dL/dP are accumulated per thread, folded with a
//     16-slot butterfly (16 shuffles instead of 13x5), then per CTA, and the LAST CTA
//     of an image (atomic ticket) adds the CTA partials in a fixed order in fp64 ->
//     deterministic, no output needs pre-zeroing, no second launch.
// Roofline: HBM-bound stream (no dense contraction => no tensor cores); algorithmic
// traffic (4 + 4C)/V + 4C + 4/V bytes per warped pixel for fp32 NCHW (32 B at C=3, V=1).
#include "dvf_internal.h"
#include "dvf_math.cuh"
#include "dvf_reduce.cuh"

namespace dvf {

constexpr int kPPT = 4;  // pixels per thread per iteration

struct LevelDev {
  int H, W, HW;
  FastDiv divW;
  float fW1, fH1, rW1, rH1, halfW, halfH, inv_n;
  int allow_fast;
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
  float* partials;        // [B*blocks_per_image][V][kRedSlots]
  double* img_terms;      // [B][V]
  unsigned* img_counter;  // [B]   zero between launches
  unsigned* lvl_counter;  // [1]   zero between launches
};

struct LossParams {
  int n_levels, B, C, V;
  int zeros_padding, need_grad;
  float* terms;  // [n_levels*V]
  LevelDev lv[DVF_MAX_LEVELS];
};

template <int kC, int kV>
__global__ void __launch_bounds__(kThreads, 2) photo_loss_kernel(const __grid_constant__ LossParams prm) {
  __shared__ float s_P[kV][12];
  __shared__ float s_red[kThreads / 32][kV][kRedSlots];
  __shared__ int s_flag;

  // ---- which level / image / run of pixels does this CTA own -------------------------
  int l = 0;
  while (l + 1 < prm.n_levels && (int)blockIdx.x >= prm.lv[l + 1].block_begin) ++l;
  const LevelDev& lv = prm.lv[l];
  const int rel = (int)blockIdx.x - lv.block_begin;
  const int b = rel / lv.blocks_per_image;
  const int chunk = rel - b * lv.blocks_per_image;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = lv.H, W = lv.W, HW = lv.HW;
  const int C = (kC > 0) ? kC : prm.C;
  const bool zeros = prm.zeros_padding != 0;
  const bool need_grad = prm.need_grad != 0;
  const bool allow_fast = lv.allow_fast != 0;

  if (tid < kV * 12) s_P[tid / 12][tid % 12] = lv.P[((size_t)b * kV + tid / 12) * 12 + tid % 12];
  float M[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) M[k] = __ldg(lv.Kinv + b * 9 + k);
  __syncthreads();

  float acc[kV][kRedSlots];
#pragma unroll
  for (int v = 0; v < kV; ++v)
#pragma unroll
    for (int k = 0; k < kRedSlots; ++k) acc[v][k] = 0.0f;

  const size_t img_off = (size_t)b * C * HW;
  const float* depth_b = lv.depth + (size_t)b * HW;
  const float* tgt_b = lv.tgt + img_off;
  const int px_begin = chunk * (kThreads * kPPT * lv.iters);

  for (int it = 0; it < lv.iters; ++it) {
    const int base = px_begin + it * (kThreads * kPPT) + tid;
    // ---- view-independent loads -----------------------------------------------------
    bool ok[kPPT];
    float dep[kPPT];
    Cam cam[kPPT];
    float gd[kPPT];
    float tg[kPPT][kC > 0 ? kC : 1];
    float gt[kPPT][kC > 0 ? kC : 1];
#pragma unroll
    for (int p = 0; p < kPPT; ++p) {
      const int idx = base + p * kThreads;
      ok[p] = idx < HW;
      dep[p] = ok[p] ? ld_stream(depth_b + idx) : 1.0f;
      if (kC > 0) {
#pragma unroll
        for (int c = 0; c < kC; ++c) {
          tg[p][c] = ok[p] ? ld_stream(tgt_b + (size_t)c * HW + idx) : 0.0f;
          gt[p][c] = 0.0f;
        }
      }
    }
#pragma unroll
    for (int p = 0; p < kPPT; ++p) {
      const int idx = base + p * kThreads;
      const int i = (int)fastdiv((uint32_t)idx, lv.divW);
      const int j = idx - i * W;
      pixel_to_cam(M, dep[p], i, j, cam[p]);
      gd[p] = 0.0f;
    }

    // ---- per view -------------------------------------------------------------------
#pragma unroll
    for (int v = 0; v < kV; ++v) {
      float P[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) P[k] = s_P[v][k];
      const float* src_b = lv.src[v] + img_off;
      float* gsrc_b = lv.gsrc[v] ? lv.gsrc[v] + img_off : nullptr;
#pragma unroll
      for (int p = 0; p < kPPT; ++p) {
        const int idx = base + p * kThreads;
        Proj pr;
        Loc L;
        if (zeros) {
          project<true>(P, cam[p], lv.fW1, lv.fH1, lv.rW1, lv.rH1, allow_fast, pr);
          locate<true>(pr.xn, pr.yn, H, W, lv.halfW, lv.halfH, L);
        } else {
          project<false>(P, cam[p], lv.fW1, lv.fH1, lv.rW1, lv.rH1, allow_fast, pr);
          locate<false>(pr.xn, pr.yn, H, W, lv.halfW, lv.halfH, L);
        }
        const bool live = ok[p];
        const bool bnw = L.bnw && live, bne = L.bne && live, bsw = L.bsw && live, bse = L.bse && live;
        const int o_nw = L.y0 * W + L.x0;
        const float wnw = mul(L.s, L.e), wne = mul(L.s, L.w), wsw = mul(L.n, L.e), wse = mul(L.n, L.w);
        const float ex = (lv.expl && live) ? ld_stream(lv.expl + (size_t)b * lv.expl_bstride + (size_t)v * HW + idx) : 1.0f;
        const bool has_expl = lv.expl != nullptr;
        float gx = 0.0f, gy = 0.0f, ge = 0.0f, lsum = 0.0f;
        bool any = false;

        if (kC > 0) {
          float vnw[kC > 0 ? kC : 1], vne[kC > 0 ? kC : 1], vsw[kC > 0 ? kC : 1], vse[kC > 0 ? kC : 1];
#pragma unroll
          for (int c = 0; c < kC; ++c) {
            const float* pl = src_b + (size_t)c * HW + o_nw;
            vnw[c] = bnw ? __ldg(pl) : 0.0f;
            vne[c] = bne ? __ldg(pl + 1) : 0.0f;
            vsw[c] = bsw ? __ldg(pl + W) : 0.0f;
            vse[c] = bse ? __ldg(pl + W + 1) : 0.0f;
          }
          float gwc[kC > 0 ? kC : 1];
#pragma unroll
          for (int c = 0; c < kC; ++c) {
            const float wv = bilerp(vnw[c], vne[c], vsw[c], vse[c], wnw, wne, wsw, wse);
            any |= (wv != 0.0f);
            const float d0 = sub(tg[p][c], wv);
            const float d1 = has_expl ? mul(d0, ex) : d0;
            lsum += fabsf(d1);
            // sign(d1)/N, sign(0) = 0
            const float gd1 = (d1 < 0.0f || d1 > 0.0f) ? copysignf(lv.inv_n, d1) : 0.0f;  // sign(0)=sign(NaN)=0
            const float gd0 = has_expl ? mul(gd1, ex) : gd1;
            ge = add(ge, mul(gd1, d0));
            gwc[c] = gd0;  // dL/d(tgt - warped)
          }
          if (!any) {
            lsum = 0.0f;
            ge = 0.0f;
          }
          acc[v][12] += lsum;
          if (need_grad) {
#pragma unroll
            for (int c = 0; c < kC; ++c) {
              const float g = any ? gwc[c] : 0.0f;
              gt[p][c] = add(gt[p][c], g);
              bilerp_grad(vnw[c], vne[c], vsw[c], vse[c], L, -g, gx, gy);
              if (gsrc_b) {
                float* gp = gsrc_b + (size_t)c * HW + o_nw;
                if (bnw) atomicAdd(gp, mul(wnw, -g));
                if (bne) atomicAdd(gp + 1, mul(wne, -g));
                if (bsw) atomicAdd(gp + W, mul(wsw, -g));
                if (bse) atomicAdd(gp + W + 1, mul(wse, -g));
              }
            }
          }
        } else {
          // generic channel count (feature maps in NCHW): pass 1 = mask, loss, d/d(ix,iy)
          for (int c = 0; c < C; ++c) {
            const float* pl = src_b + (size_t)c * HW + o_nw;
            const float a0 = bnw ? __ldg(pl) : 0.0f, a1 = bne ? __ldg(pl + 1) : 0.0f;
            const float a2 = bsw ? __ldg(pl + W) : 0.0f, a3 = bse ? __ldg(pl + W + 1) : 0.0f;
            const float wv = bilerp(a0, a1, a2, a3, wnw, wne, wsw, wse);
            any |= (wv != 0.0f);
            const float t = live ? ld_stream(tgt_b + (size_t)c * HW + idx) : 0.0f;
            const float d0 = sub(t, wv);
            const float d1 = has_expl ? mul(d0, ex) : d0;
            lsum += fabsf(d1);
            const float gd1 = (d1 < 0.0f || d1 > 0.0f) ? copysignf(lv.inv_n, d1) : 0.0f;  // sign(0)=sign(NaN)=0
            const float gd0 = has_expl ? mul(gd1, ex) : gd1;
            ge = add(ge, mul(gd1, d0));
            bilerp_grad(a0, a1, a2, a3, L, -gd0, gx, gy);
          }
          if (!any) {
            lsum = 0.0f;
            ge = 0.0f;
            gx = 0.0f;
            gy = 0.0f;
          }
          acc[v][12] += lsum;
          // pass 2 = gradients to the target and (scatter) to the source maps
          if (need_grad && live && (gsrc_b || lv.gtgt)) {
            float* gtgt_b = lv.gtgt ? lv.gtgt + img_off : nullptr;
            for (int c = 0; c < C; ++c) {
              float g = 0.0f;
              if (any) {
                const float* pl = src_b + (size_t)c * HW + o_nw;
                const float a0 = bnw ? __ldg(pl) : 0.0f, a1 = bne ? __ldg(pl + 1) : 0.0f;
                const float a2 = bsw ? __ldg(pl + W) : 0.0f, a3 = bse ? __ldg(pl + W + 1) : 0.0f;
                const float wv = bilerp(a0, a1, a2, a3, wnw, wne, wsw, wse);
                const float t = ld_stream(tgt_b + (size_t)c * HW + idx);
                const float d0 = sub(t, wv);
                const float d1 = has_expl ? mul(d0, ex) : d0;
                const float gd1 = (d1 < 0.0f || d1 > 0.0f) ? copysignf(lv.inv_n, d1) : 0.0f;  // sign(0)=sign(NaN)=0
                g = has_expl ? mul(gd1, ex) : gd1;
              }
              if (gtgt_b) {
                // views are visited in order by the same thread: plain read-modify-write
                float* q = gtgt_b + (size_t)c * HW + idx;
                *q = (v == 0) ? g : add(*q, g);
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
        }

        if (need_grad) {
          if (lv.gexpl && live) st_stream(lv.gexpl + ((size_t)b * kV + v) * HW + idx, ge);
          ChainGrad cg;
          chain_backward(P, cam[p], pr, L, gx, gy, lv.fW1, lv.fH1, lv.rW1, lv.rH1, cg);
          if (live) {
            gd[p] = add(gd[p], cg.gdepth);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
              for (int k = 0; k < 3; ++k) acc[v][r * 4 + k] = fmaf(cg.gq[r], cam[p].cam[k], acc[v][r * 4 + k]);
              acc[v][r * 4 + 3] += cg.gq[r];
            }
          }
        }
      }  // pixels
    }    // views

    if (need_grad) {
#pragma unroll
      for (int p = 0; p < kPPT; ++p) {
        const int idx = base + p * kThreads;
        if (ok[p]) {
          if (lv.gdepth) st_stream(lv.gdepth + (size_t)b * HW + idx, gd[p]);
          if (kC > 0 && lv.gtgt) {
#pragma unroll
            for (int c = 0; c < kC; ++c) st_stream(lv.gtgt + img_off + (size_t)c * HW + idx, gt[p][c]);
          }
        }
      }
    }
  }  // iterations

  // ---- CTA reduction of {dP[12], loss} per view --------------------------------------
#pragma unroll
  for (int v = 0; v < kV; ++v) {
    const float r = butterfly16(acc[v], lane);
    if ((lane & 1) == 0) s_red[warp][v][butterfly_slot(lane)] = r;
  }
  __syncthreads();
  float* my_part = lv.partials + (size_t)rel * kV * kRedSlots;
  if (tid < kV * kRedSlots) {
    const int v = tid / kRedSlots, s = tid % kRedSlots;
    float t = 0.0f;
#pragma unroll
    for (int w8 = 0; w8 < kThreads / 32; ++w8) t += s_red[w8][v][s];
    __stcg(my_part + tid, t);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = (atomicAdd(lv.img_counter + b, 1u) == (unsigned)(lv.blocks_per_image - 1));
  __syncthreads();
  if (!s_flag) return;

  // ---- last CTA of image b: fixed-order fp64 fold of the CTA partials ----------------
  __threadfence();
  const float* img_part = lv.partials + (size_t)b * lv.blocks_per_image * kV * kRedSlots;
  for (int pair = tid >> 3; pair < kV * kRedSlots; pair += kThreads / 8) {
    const int v = pair / kRedSlots, s = pair % kRedSlots;
    const double sum = group8_sum(img_part + v * kRedSlots + s, lv.blocks_per_image, kV * kRedSlots, tid & 7);
    if ((tid & 7) == 0) {
      if (s < 12) {
        if (lv.gP) lv.gP[((size_t)b * kV + v) * 12 + s] = (float)sum;
      } else if (s == 12) {
        lv.img_terms[(size_t)b * kV + v] = sum;
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    lv.img_counter[b] = 0u;
    s_flag = (atomicAdd(lv.lvl_counter, 1u) == (unsigned)(prm.B - 1));
  }
  __syncthreads();
  if (!s_flag) return;

  // ---- last image of the level: loss terms ---------------------------------------------
  __threadfence();
  __shared__ double s_term[kThreads / 32];
  for (int v = 0; v < kV; ++v) {
    double s = 0.0;
    for (int bb = tid; bb < prm.B; bb += kThreads) s += __ldcg(lv.img_terms + (size_t)bb * kV + v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_term[warp] = s;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w8 = 0; w8 < kThreads / 32; ++w8) t += s_term[w8];
      prm.terms[l * kV + v] = (float)(t / ((double)prm.B * (double)C * (double)HW));
    }
    __syncthreads();
  }
  if (tid == 0) *lv.lvl_counter = 0u;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct Plan {
  int blocks_per_image[DVF_MAX_LEVELS];
  int iters[DVF_MAX_LEVELS];
  int block_begin[DVF_MAX_LEVELS];
  int total_blocks;
  size_t off_partials[DVF_MAX_LEVELS], off_terms[DVF_MAX_LEVELS], off_cnt[DVF_MAX_LEVELS], off_lcnt[DVF_MAX_LEVELS];
  size_t bytes;
};

static int make_plan(const dvf_loss_desc* d, const dvf_level* levels, Plan& pl) {
  if (!d || !levels) return DVF_EINVAL_NULL;
  if (d->B <= 0 || d->C <= 0 || d->V <= 0 || d->V > DVF_MAX_VIEWS || d->n_levels <= 0 || d->n_levels > DVF_MAX_LEVELS)
    return DVF_EINVAL_SHAPE;
  long long total_px = 0;
  for (int l = 0; l < d->n_levels; ++l) {
    if (levels[l].H <= 0 || levels[l].W <= 0) return DVF_EINVAL_SHAPE;
    if ((long long)levels[l].H * levels[l].W >= (1ll << 30)) return DVF_EINVAL_SHAPE;
    total_px += (long long)levels[l].H * levels[l].W * d->B;
  }
  // aim for >= ~12 CTAs per SM over the whole launch, 1..4 chunks of 1024 px per CTA
  const long long chunk = (long long)kThreads * kPPT;
  const long long want_blocks = 12ll * num_sms();
  int iters = (int)(total_px / (chunk * want_blocks));
  if (iters < 1) iters = 1;
  if (iters > 4) iters = 4;
  size_t off = 0;
  int begin = 0;
  for (int l = 0; l < d->n_levels; ++l) {
    const long long HW = (long long)levels[l].H * levels[l].W;
    long long chunks = (HW + chunk - 1) / chunk;
    int it = iters;
    if (it > chunks) it = (int)chunks;
    pl.iters[l] = it;
    pl.blocks_per_image[l] = (int)((chunks + it - 1) / it);
    pl.block_begin[l] = begin;
    begin += pl.blocks_per_image[l] * d->B;
    pl.off_partials[l] = off;
    off += align_up((size_t)pl.blocks_per_image[l] * d->B * d->V * kRedSlots * sizeof(float), 256);
    pl.off_terms[l] = off;
    off += align_up((size_t)d->B * d->V * sizeof(double), 256);
    pl.off_cnt[l] = off;
    off += align_up((size_t)d->B * sizeof(unsigned), 256);
    pl.off_lcnt[l] = off;
    off += 256;
  }
  pl.total_blocks = begin;
  pl.bytes = off;
  return DVF_OK;
}

template <int kC, int kV>
static void launch_loss(const LossParams& prm, int blocks, cudaStream_t st) {
  photo_loss_kernel<kC, kV><<<blocks, kThreads, 0, st>>>(prm);
}

}  // namespace dvf

using namespace dvf;

DVF_EXPORT size_t dvf_photo_loss_workspace_bytes(const dvf_loss_desc* d, const dvf_level* levels) {
  Plan pl;
  if (make_plan(d, levels, pl) != DVF_OK) return 0;
  return pl.bytes;
}

DVF_EXPORT int dvf_photo_loss_fused(const dvf_loss_desc* d, const dvf_level* levels, float* terms,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  Plan pl;
  int st = make_plan(d, levels, pl);
  if (st != DVF_OK) return st;
  if (!terms) return DVF_EINVAL_NULL;
  if (d->dtype != DVF_F32 || d->layout != DVF_NCHW) return DVF_EUNSUPPORTED;
  if (d->padding != DVF_PAD_ZEROS && d->padding != DVF_PAD_BORDER) return DVF_EINVAL_DTYPE;
  if (!workspace || workspace_bytes < pl.bytes) return DVF_EWORKSPACE;
  if (!aligned(workspace, 256)) return DVF_EINVAL_ALIGN;

  LossParams prm;
  prm.n_levels = d->n_levels;
  prm.B = d->B;
  prm.C = d->C;
  prm.V = d->V;
  prm.zeros_padding = d->padding == DVF_PAD_ZEROS;
  prm.terms = terms;
  bool need_grad = false;
  char* ws = static_cast<char*>(workspace);
  for (int l = 0; l < d->n_levels; ++l) {
    const dvf_level& s = levels[l];
    LevelDev& t = prm.lv[l];
    if (!s.depth || !s.tgt || !s.P || !s.Kinv) return DVF_EINVAL_NULL;
    if (!aligned(s.depth, 4) || !aligned(s.tgt, 4) || !aligned(s.P, 4) || !aligned(s.Kinv, 4)) return DVF_EINVAL_ALIGN;
    t.H = s.H;
    t.W = s.W;
    t.HW = s.H * s.W;
    t.divW = make_fastdiv((uint32_t)s.W);
    t.fW1 = (float)(s.W - 1);
    t.fH1 = (float)(s.H - 1);
    t.allow_fast = (s.W > 1 && s.H > 1);
    t.rW1 = t.allow_fast ? (float)(1.0 / (double)t.fW1) : 0.0f;
    t.rH1 = t.allow_fast ? (float)(1.0 / (double)t.fH1) : 0.0f;
    t.halfW = (float)s.W / 2.0f;
    t.halfH = (float)s.H / 2.0f;
    t.inv_n = 1.0f / (float)((double)d->B * d->C * s.H * s.W);
    t.depth = s.depth;
    t.tgt = static_cast<const float*>(s.tgt);
    t.expl = s.expl;
    t.expl_bstride = s.expl_bstride;
    t.P = s.P;
    t.Kinv = s.Kinv;
    t.gdepth = s.gdepth;
    t.gexpl = s.gexpl;
    t.gtgt = static_cast<float*>(s.gtgt);
    t.gP = s.gP;
    for (int v = 0; v < DVF_MAX_VIEWS; ++v) {
      t.src[v] = v < d->V ? static_cast<const float*>(s.src[v]) : nullptr;
      t.gsrc[v] = v < d->V ? static_cast<float*>(s.gsrc[v]) : nullptr;
      if (v < d->V && !s.src[v]) return DVF_EINVAL_NULL;
      if (t.gsrc[v]) need_grad = true;
    }
    if (s.gdepth || s.gexpl || s.gtgt || s.gP) need_grad = true;
    if (s.gexpl && !s.expl) return DVF_EINVAL_NULL;
    t.block_begin = pl.block_begin[l];
    t.blocks_per_image = pl.blocks_per_image[l];
    t.iters = pl.iters[l];
    t.partials = reinterpret_cast<float*>(ws + pl.off_partials[l]);
    t.img_terms = reinterpret_cast<double*>(ws + pl.off_terms[l]);
    t.img_counter = reinterpret_cast<unsigned*>(ws + pl.off_cnt[l]);
    t.lvl_counter = reinterpret_cast<unsigned*>(ws + pl.off_lcnt[l]);
  }
  prm.need_grad = need_grad;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  const int nb = pl.total_blocks;
  if (d->C == 3) {
    switch (d->V) {
      case 1: launch_loss<3, 1>(prm, nb, cs); break;
      case 2: launch_loss<3, 2>(prm, nb, cs); break;
      case 3: launch_loss<3, 3>(prm, nb, cs); break;
      default: launch_loss<3, 4>(prm, nb, cs); break;
    }
  } else {
    switch (d->V) {
      case 1: launch_loss<0, 1>(prm, nb, cs); break;
      case 2: launch_loss<0, 2>(prm, nb, cs); break;
      case 3: launch_loss<0, 3>(prm, nb, cs); break;
      default: launch_loss<0, 4>(prm, nb, cs); break;
    }
  }
  return launch_status();
}
