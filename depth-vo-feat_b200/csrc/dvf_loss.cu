// dvf_loss.cu -- host side of the fused masked reconstruction loss (forward + backward in ONE pass).
//
// Replaces photometric_reconstruction_loss of pytorch_version/loss_functions.py:7-20,
// loss_functions_sfm.py:9-46 and loss_function_sfm_old.py:7-46 (which call inverse_warp,
// inverse_warp.py:160-193, once per view and scale and then run ~60 elementwise passes each).
// Device code: dvf_loss_kernel.cuh (instantiated in dvf_loss_inst_*.cu so the variants build in parallel).
// Roofline: HBM-bound stream (no dense contraction => no tensor cores); algorithmic traffic
// (4 + 4C + 4)/V + 4C bytes per warped pixel for fp32 NCHW (32 B at C=3, V=1).
#include "dvf_loss_nhwc.cuh"

namespace dvf {

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct Plan {
  int blocks_per_image[DVF_MAX_LEVELS];
  int iters[DVF_MAX_LEVELS];
  int block_begin[DVF_MAX_LEVELS];
  int total_blocks;
  // image kernel (balanced split, dvf_loss_kernel.cuh): 256-pixel units, image-major
  int units_per_image[DVF_MAX_LEVELS], unit_base[DVF_MAX_LEVELS], slots_c3[DVF_MAX_LEVELS];
  long long total_units;
  int units_per_image_all, piece_overhead;
  int max_ctas_c3;
  size_t off_partials[DVF_MAX_LEVELS], off_terms[DVF_MAX_LEVELS], off_cnt[DVF_MAX_LEVELS], off_lcnt[DVF_MAX_LEVELS];
  size_t off_gM, off_pcnt;
  size_t bytes;
};

static bool uses_c3_kernel(const dvf_loss_desc* d, const dvf_level* levels) {
  if (d->C != 3 || d->layout != DVF_NCHW || d->dtype != DVF_F32) return false;
  for (int l = 0; l < d->n_levels; ++l) {   // the register-resident image kernel has no scatter / d-target outputs
    if (levels[l].gtgt) return false;
    for (int v = 0; v < d->V && v < DVF_MAX_VIEWS; ++v)
      if (levels[l].gsrc[v]) return false;
  }
  return true;
}

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
  // aim for ~6 CTAs per SM over the whole launch (4 resident): fixed per-CTA cost vs tail balance
  const long long chunk = kPlanUnit;
  const long long want_blocks = 6ll * num_sms();   // measured flat between 4 and 16 CTAs per SM (profiles/r1_summary.md)
  int iters = (int)(total_px / (chunk * want_blocks));
  if (iters < 1) iters = 1;
  if (iters > 64) iters = 64;
  // image kernel: every resident CTA slot gets an equal share of the 256-pixel units of the launch.  The grid is
  // decided at launch time (occupancy of the variant); the partial-sum slots are sized for the largest one.
  pl.max_ctas_c3 = 8 * num_sms();
  {
    // fixed cost of a piece in units.  Image kernel, measured on C2 (V = 1): 2 -> 73.7 us, 3 -> 69.6, 4 -> 67.4,
    // 5 -> 67.4, 6 -> 68.5, 8 -> 69.7; with V = 2 a unit is twice the work: 3 -> 120.1 us, 4 -> 122.6, 6 -> 125.4.
    // A unit of the channels-last kernel is C/kVec times more work again, so its pieces cost less than one unit.
    // round 2 (programmatic dependent launch, cheaper pieces), C2: 3 -> 53.8 us, 4 -> 52.2, 5 -> 51.4, 6 -> 51.4, 8 -> 51.6,
    // 10 -> 52.0; V = 2 (C3 / C5loss shapes) is flat between 2 and 6.
    // chained launches (DVF_FLAG_PDL_CHAINED, smaller grids): 3 -> 49.6, 4 -> 48.9, 5 -> 49.3, 6 -> 49.7.
    const bool chained = (d->flags & DVF_FLAG_PDL) && (d->flags & DVF_FLAG_PDL_CHAINED);
    const int by_views = d->V == 1 ? (chained ? 4 : 5) : (d->V == 2 ? 3 : 2);
    pl.piece_overhead = d->piece_overhead > 0 ? d->piece_overhead : (d->layout == DVF_NHWC ? 0 : by_views);
  }
  long long per_image = 0;
  for (int l = 0; l < d->n_levels; ++l) {
    const long long HW = (long long)levels[l].H * levels[l].W;
    pl.units_per_image[l] = (int)((HW + kUnitPx - 1) / kUnitPx);
    pl.unit_base[l] = (int)per_image;
    per_image += pl.units_per_image[l] + pl.piece_overhead;
  }
  pl.units_per_image_all = (int)per_image;
  pl.total_units = per_image * d->B;
  if (pl.total_units >= (1ll << 31)) return DVF_EINVAL_SHAPE;
  for (int l = 0; l < d->n_levels; ++l) {
    // an image of U units is cut by at most ceil(U / floor(T/G)) + 1 CTAs, and by at most U of them
    const long long G = pl.total_units < pl.max_ctas_c3 ? pl.total_units : pl.max_ctas_c3;
    const long long per = pl.total_units / G;   // >= 1
    long long slots = (pl.units_per_image[l] + per - 1) / per + 1;
    if (slots > pl.units_per_image[l]) slots = pl.units_per_image[l];
    pl.slots_c3[l] = (int)slots;
  }
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
    const int slots = pl.blocks_per_image[l] > pl.slots_c3[l] ? pl.blocks_per_image[l] : pl.slots_c3[l];
    off += align_up((size_t)slots * d->B * d->V * kRedSlots * sizeof(float), 256);
    pl.off_terms[l] = off;
    off += align_up((size_t)d->B * d->V * sizeof(double), 256);
    pl.off_cnt[l] = off;
    off += align_up((size_t)d->B * sizeof(unsigned), 256);
    pl.off_lcnt[l] = off;
    off += 256;
  }
  pl.off_gM = off;
  off += align_up((size_t)d->n_levels * d->B * d->V * 12 * sizeof(double), 256);
  pl.off_pcnt = off;
  off += align_up((size_t)d->B * sizeof(unsigned), 256);
  pl.total_blocks = begin;
  pl.bytes = off;
  return DVF_OK;
}

}  // namespace dvf

using namespace dvf;

DVF_EXPORT size_t dvf_photo_loss_workspace_bytes(const dvf_loss_desc* d, const dvf_level* levels) {
  Plan pl;
  if (make_plan(d, levels, pl) != DVF_OK) return 0;
  return pl.bytes;
}

// ext_grad: levels[0].tgt holds d L / d warped and the image kernel runs as the backward of the materialised warp (kExt)
static int run_loss(const dvf_loss_desc* d, const dvf_level* levels, const dvf_pose_args* pose, float* terms,
                    void* workspace, size_t workspace_bytes, void* stream, bool ext_grad = false) {
  Plan pl;
  int st = make_plan(d, levels, pl);
  if (st != DVF_OK) return st;
  if (!terms) return DVF_EINVAL_NULL;
  if (d->dtype != DVF_F32 && d->dtype != DVF_BF16) return DVF_EINVAL_DTYPE;
  if (d->layout != DVF_NCHW && d->layout != DVF_NHWC) return DVF_EINVAL_DTYPE;
  const bool nhwc = d->layout == DVF_NHWC;
  const bool bf16 = d->dtype == DVF_BF16;
  if (bf16 && !nhwc) return DVF_EUNSUPPORTED;                 // bf16 maps are channels-last only
  if (nhwc) {
    const int vec = bf16 ? 8 : 4;
    const int lpp = d->C / vec;
    if (d->C % vec != 0 || lpp < 1 || lpp > 32 || (lpp & (lpp - 1)) != 0) return DVF_EUNSUPPORTED;   // C = vec * 2^k
    for (int l = 0; l < d->n_levels; ++l)   // the kernel addresses one image with 32-bit byte offsets
      if ((long long)levels[l].H * levels[l].W * d->C * 4 >= (1ll << 31)) return DVF_EUNSUPPORTED;
  }
  if (d->padding != DVF_PAD_ZEROS && d->padding != DVF_PAD_BORDER) return DVF_EINVAL_DTYPE;
  if (!workspace || workspace_bytes < pl.bytes) return DVF_EWORKSPACE;
  if (!aligned(workspace, 256)) return DVF_EINVAL_ALIGN;

  LossParams prm;
  prm.n_levels = d->n_levels;
  prm.B = d->B;
  prm.C = d->C;
  prm.V = d->V;
  const bool zeros = d->padding == DVF_PAD_ZEROS;
  prm.rotation = 0;
  prm.pose_vec = nullptr;
  prm.K = prm.Kinv = nullptr;
  prm.gvec = nullptr;
  if (pose) {
    if (!pose->vec || !pose->K || !pose->Kinv || !pose->downscale) return DVF_EINVAL_NULL;
    if (pose->rotation != DVF_ROT_EULER && pose->rotation != DVF_ROT_QUAT) return DVF_EINVAL_DTYPE;
    prm.rotation = pose->rotation;
    prm.pose_vec = pose->vec;
    prm.K = pose->K;
    prm.Kinv = pose->Kinv;
    prm.gvec = pose->gvec;
  }
  prm.terms = terms;
  if (d->mean_batch < 0 || (d->mean_batch > 0 && d->mean_batch < d->B)) return DVF_EINVAL_SHAPE;
  const int mean_batch = d->mean_batch > 0 ? d->mean_batch : d->B;
  prm.mean_batch = mean_batch;
  prm.upstream = d->upstream;
  prm.nan_flags = (d->flags & DVF_FLAG_NAN_CHECK) ? d->nan_flags : nullptr;
  if ((d->flags & DVF_FLAG_NAN_CHECK) && !d->nan_flags) return DVF_EINVAL_NULL;
  prm.ctas_per_sm = d->ctas_per_sm;
  prm.disparity = (d->flags & DVF_FLAG_DISPARITY) ? 1 : 0;
  prm.disp_eps = d->disp_eps;
  prm.img_scale = d->img_scale == 0.0f ? 1.0f : d->img_scale;
  if (prm.img_scale != 1.0f && !(d->C == 3 && d->layout == DVF_NCHW)) return DVF_EUNSUPPORTED;   // images only
  // the image kernel carries the glue in zeros-padding variants only (the reference's default, and what its callers use)
  if ((prm.disparity || prm.img_scale != 1.0f) && d->padding != DVF_PAD_ZEROS && uses_c3_kernel(d, levels)) return DVF_EUNSUPPORTED;
  if (d->n_peers < 0 || d->n_peers > DVF_MAX_PEERS || (d->n_peers > 0 && (d->peer_rank < 0 || d->peer_rank >= d->n_peers)))
    return DVF_EINVAL_SHAPE;
  if (d->n_peers > 0 && !d->peer_terms) return DVF_EINVAL_NULL;
  prm.n_peers = d->n_peers;
  prm.peer_rank = d->peer_rank;
  for (int q = 0; q < DVF_MAX_PEERS; ++q) {
    prm.peer_terms[q] = q < d->n_peers ? d->peer_terms[q] : nullptr;
    if (q < d->n_peers && (!prm.peer_terms[q] || !aligned(prm.peer_terms[q], 4))) return DVF_EINVAL_NULL;
  }
  prm.pdl = (d->flags & DVF_FLAG_PDL) ? ((d->flags & DVF_FLAG_PDL_CHAINED) ? 2 : 1) : 0;
  if (d->grad_dtype != DVF_F32 && !(d->grad_dtype == DVF_BF16 && nhwc && bf16)) return DVF_EUNSUPPORTED;
  prm.grad_bf16 = d->grad_dtype == DVF_BF16;
  bool need_grad = false;
  char* ws = static_cast<char*>(workspace);
  for (int l = 0; l < d->n_levels; ++l) {
    const dvf_level& s = levels[l];
    LevelDev& t = prm.lv[l];
    if (!s.depth || !s.tgt) return DVF_EINVAL_NULL;
    if (!pose && (!s.P || !s.Kinv)) return DVF_EINVAL_NULL;
    if (!aligned(s.depth, 4) || !aligned(s.tgt, 4) || !aligned(s.P, 4) || !aligned(s.Kinv, 4)) return DVF_EINVAL_ALIGN;
    if (nhwc && (!aligned(s.tgt, 16) || !aligned(s.gtgt, 16))) return DVF_EINVAL_ALIGN;
    t.H = s.H;
    t.W = s.W;
    t.HW = s.H * s.W;
    t.divW = make_fastdiv((uint32_t)s.W);
    t.geo = make_geo(s.H, s.W, (d->flags & DVF_FLAG_ALIGN_CORNERS) != 0);
    t.allow_fast = (s.W > 1 && s.H > 1 && s.W - 1 <= kMaxConstDiv && s.H - 1 <= kMaxConstDiv);
    t.prefetch_rows = s.W <= 512 ? 1 : 0;
    t.ds = pose ? pose->downscale[l] : 1.0f;
    t.inv_n = 1.0f / (float)((double)mean_batch * d->C * s.H * s.W);
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
      if (nhwc && v < d->V && (!aligned(s.src[v], 16) || !aligned(s.gsrc[v], 16))) return DVF_EINVAL_ALIGN;
      if (t.gsrc[v]) need_grad = true;
    }
    if (s.gdepth || s.gexpl || s.gtgt || s.gP || (pose && pose->gvec)) need_grad = true;
    if (s.gexpl && !s.expl) return DVF_EINVAL_NULL;
    t.block_begin = pl.block_begin[l];
    t.blocks_per_image = pl.blocks_per_image[l];
    t.iters = pl.iters[l];
    t.px_per_cta = pl.iters[l] * kPlanUnit;
    t.slots_per_image = pl.blocks_per_image[l];
    t.unit_base = pl.unit_base[l];
    t.units_per_image = pl.units_per_image[l];
    t.partials = reinterpret_cast<float*>(ws + pl.off_partials[l]);
    t.img_terms = reinterpret_cast<double*>(ws + pl.off_terms[l]);
    t.img_counter = reinterpret_cast<unsigned*>(ws + pl.off_cnt[l]);
    t.lvl_counter = reinterpret_cast<unsigned*>(ws + pl.off_lcnt[l]);
  }
  prm.need_grad = need_grad;
  prm.gM_ws = reinterpret_cast<double*>(ws + pl.off_gM);
  prm.pose_counter = reinterpret_cast<unsigned*>(ws + pl.off_pcnt);
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  if (d->flags & DVF_FLAG_ZERO_GSRC) {
    const size_t gesz = 4;   // source-map gradients are fp32
    for (int l = 0; l < d->n_levels; ++l)
      for (int v = 0; v < d->V; ++v)
        if (levels[l].gsrc[v]) {
          const cudaError_t e = cudaMemsetAsync(levels[l].gsrc[v], 0, (size_t)d->B * d->C * levels[l].H * levels[l].W * gesz, cs);
          if (e != cudaSuccess) return (int)e;
        }
  }
  prm.total_units = (int)pl.total_units;
  prm.units_per_image_all = pl.units_per_image_all;
  prm.piece_overhead = pl.piece_overhead;
  int nb = pl.total_blocks;
#define DVF_DISPATCH_V(FN, Z)                                  \
  switch (d->V) {                                              \
    case 1: FN<1, Z>(prm, nb, cs); break;                      \
    case 2: FN<2, Z>(prm, nb, cs); break;                      \
    case 3: FN<3, Z>(prm, nb, cs); break;                      \
    default: FN<4, Z>(prm, nb, cs); break;                     \
  }
  if (nhwc) {
    for (int l = 0; l < d->n_levels; ++l) prm.lv[l].slots_per_image = pl.slots_c3[l];
    nb = pl.max_ctas_c3;
#define DVF_DISPATCH_NHWC(Z)                                        \
  switch (d->V) {                                                   \
    case 1: launch_loss_nhwc<1, Z>(prm, nb, bf16, cs); break;       \
    case 2: launch_loss_nhwc<2, Z>(prm, nb, bf16, cs); break;       \
    case 3: launch_loss_nhwc<3, Z>(prm, nb, bf16, cs); break;       \
    default: launch_loss_nhwc<4, Z>(prm, nb, bf16, cs); break;      \
  }
    if (zeros) { DVF_DISPATCH_NHWC(true) } else { DVF_DISPATCH_NHWC(false) }
#undef DVF_DISPATCH_NHWC
  } else if (uses_c3_kernel(d, levels)) {
    // balanced split: the launcher picks the grid (resident CTAs of the variant, at most max_ctas_c3)
    for (int l = 0; l < d->n_levels; ++l) prm.lv[l].slots_per_image = pl.slots_c3[l];
    nb = pl.max_ctas_c3;
    bool expl = false;
    for (int l = 0; l < d->n_levels; ++l) expl |= (levels[l].expl != nullptr);
    for (int l = 0; l < d->n_levels; ++l)
      if (expl && !levels[l].expl) return DVF_EINVAL_NULL;   // all levels or none
    // bulk-copy (TMA) path: contiguous runs must start on 16-byte boundaries for every image and level
    bool tma = (d->flags & DVF_FLAG_NO_TMA) == 0;
    for (int l = 0; l < d->n_levels; ++l) {
      const dvf_level& s = levels[l];
      tma = tma && ((s.H * s.W) % 4 == 0) && aligned(s.depth, 16) && aligned(s.tgt, 16);
      if (s.expl) tma = tma && aligned(s.expl, 16) && (s.expl_bstride % 4 == 0);
      for (int v = 0; v < d->V; ++v) tma = tma && aligned(s.src[v], 16);
    }
    if (ext_grad) {
      if (!tma || expl || d->V != 1 || d->n_levels != 1 || pose) return DVF_EUNSUPPORTED;
      launch_warp_bwd_fused(prm, nb, zeros, cs);
      return launch_status();
    }
#define DVF_DISPATCH_C3(Z)                                                          \
  switch (d->V) {                                                                   \
    case 1: launch_loss_c3<1, Z>(prm, nb, expl, need_grad, tma, cs); break;         \
    case 2: launch_loss_c3<2, Z>(prm, nb, expl, need_grad, tma, cs); break;         \
    case 3: launch_loss_c3<3, Z>(prm, nb, expl, need_grad, tma, cs); break;         \
    default: launch_loss_c3<4, Z>(prm, nb, expl, need_grad, tma, cs); break;        \
  }
    if (zeros) { DVF_DISPATCH_C3(true) } else { DVF_DISPATCH_C3(false) }
#undef DVF_DISPATCH_C3
  } else {
    if (zeros) { DVF_DISPATCH_V(launch_loss_cn, true) } else { DVF_DISPATCH_V(launch_loss_cn, false) }
  }
#undef DVF_DISPATCH_V
  return launch_status();
}

// ---- dvf_inverse_warp_bwd without d img, run by the image kernel (dvf_warp.cu decides when) ----------------------------
namespace dvf {
static void warp_bwd_as_loss(const dvf_desc* d, dvf_loss_desc& ld, dvf_level& lv) {
  ld = dvf_loss_desc{};
  ld.B = d->B;
  ld.C = 3;
  ld.V = 1;
  ld.n_levels = 1;
  ld.dtype = DVF_F32;
  ld.layout = DVF_NCHW;
  ld.padding = d->padding;
  ld.flags = d->flags & DVF_FLAG_ALIGN_CORNERS;
  ld.grad_dtype = DVF_F32;
  lv = dvf_level{};
  lv.H = d->H;
  lv.W = d->W;
}
bool warp_bwd_fused_ok(const dvf_desc* d, const void* gout, const void* img, const float* depth) {
  return d->C == 3 && ((long long)d->H * d->W) % 4 == 0 && aligned(gout, 16) && aligned(img, 16) && aligned(depth, 16);
}
size_t warp_bwd_fused_workspace_bytes(const dvf_desc* d) {
  dvf_loss_desc ld;
  dvf_level lv;
  warp_bwd_as_loss(d, ld, lv);
  Plan pl;
  if (make_plan(&ld, &lv, pl) != DVF_OK) return 0;
  return pl.bytes + 256;   // + the (unused) loss term of the launch
}
int warp_bwd_fused(const dvf_desc* d, const void* gout, const void* img, const float* depth, const float* P, const float* Kinv,
                   float* gdepth, float* gP, void* workspace, size_t workspace_bytes, void* stream) {
  dvf_loss_desc ld;
  dvf_level lv;
  warp_bwd_as_loss(d, ld, lv);
  lv.depth = depth;
  lv.tgt = gout;
  lv.src[0] = img;
  lv.P = P;
  lv.Kinv = Kinv;
  lv.gdepth = gdepth;
  lv.gP = gP;
  if (workspace_bytes < 256) return DVF_EWORKSPACE;
  float* const term = reinterpret_cast<float*>(static_cast<char*>(workspace) + workspace_bytes - 256);
  return run_loss(&ld, &lv, nullptr, term, workspace, workspace_bytes - 256, stream, true);
}
}  // namespace dvf

DVF_EXPORT int dvf_photo_loss_fused(const dvf_loss_desc* d, const dvf_level* levels, float* terms, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  return run_loss(d, levels, nullptr, terms, workspace, workspace_bytes, stream);
}

DVF_EXPORT int dvf_photo_loss_fused_pose(const dvf_loss_desc* d, const dvf_level* levels, const dvf_pose_args* pose,
                                         float* terms, void* workspace, size_t workspace_bytes, void* stream) {
  if (!pose) return DVF_EINVAL_NULL;
  return run_loss(d, levels, pose, terms, workspace, workspace_bytes, stream);
}

