/*
 * dvf_b200.h -- C ABI of libdvf_b200.so: the B200 (sm_100a) implementation of
 * Depth-VO-Feat's differentiable inverse warp and masked reconstruction losses.
 *
 * This is the drop-in boundary.  Every entry point replaces one function of the
 * reference's pytorch_version/ package (cited per entry as file:line, relative
 * to the reference checkout) and is what a binding in the reference's own
 * language (Python: ctypes) calls; see INTEGRATION.md for the stub.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no C++/torch types.
 *  - All pointers are DEVICE pointers owned by the caller (torch tensors'
 *    data_ptr()).  Nothing is allocated, nothing is synchronised; every launch
 *    goes to the cudaStream_t the caller passes (as void*).
 *  - Entries are re-entrant; there is no global state.  Scratch memory is a
 *    caller-provided workspace whose size dvf_*_workspace_bytes() reports.  A
 *    workspace must be zero-filled before its FIRST use (e.g. torch.zeros); the
 *    kernels restore the ticket counters they use, so it can be reused without
 *    clearing by later calls WITH THE SAME DESCRIPTOR (same shapes) on the same
 *    stream.  Re-zero it (or use another one) when the shapes change.  One
 *    workspace per in-flight call.
 *  - Return value: 0 = DVF_OK, <0 = argument error (dvf_status), >0 =
 *    cudaError_t of the failed launch.  dvf_strerror() names either.
 *  - Geometry (depth, poses, intrinsics, P, gradients w.r.t. them) is fp32.
 *    Images / feature maps are fp32 or bf16, NCHW or NHWC (dvf_desc).
 *  - Arithmetic profile: the coordinate chain reproduces torch-CPU fp32
 *    rounding step by step (FMA placement, true divisions), so bilinear cells
 *    and validity masks are bit-identical to the reference run on the CPU.
 */
#ifndef DVF_B200_H_
#define DVF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVF_ABI_VERSION 2
#define DVF_MAX_VIEWS 4   /* source views per target in one loss call          */
#define DVF_MAX_LEVELS 6  /* pyramid levels fused into one loss launch         */

typedef enum dvf_status {
  DVF_OK = 0,
  DVF_EINVAL_SHAPE = -1,   /* a size is <= 0 or exceeds a DVF_MAX_* limit      */
  DVF_EINVAL_DTYPE = -2,   /* unknown element type / layout / mode value       */
  DVF_EINVAL_ALIGN = -3,   /* pointer not aligned for the element type         */
  DVF_EINVAL_NULL = -4,    /* a required pointer is NULL                       */
  DVF_EUNSUPPORTED = -5,   /* valid request this build does not implement      */
  DVF_EWORKSPACE = -6      /* workspace missing or too small                   */
} dvf_status;

typedef enum dvf_dtype { DVF_F32 = 0, DVF_BF16 = 1 } dvf_dtype;
typedef enum dvf_layout { DVF_NCHW = 0, DVF_NHWC = 1 } dvf_layout;
typedef enum dvf_padding { DVF_PAD_ZEROS = 0, DVF_PAD_BORDER = 1 } dvf_padding;   /* inverse_warp.py:67, train.py:194 */
typedef enum dvf_rotation { DVF_ROT_EULER = 0, DVF_ROT_QUAT = 1 } dvf_rotation;  /* inverse_warp.py:152-155 */
/* Arithmetic profile of the pose chain, or-ed into the `rotation` argument of dvf_pose_proj_fwd / _bwd:
 * default = the reference on torch-CPU (MKL sin / cos, multiply-add-add tiny matmuls); DVF_ROT_REF_CUDA = the reference
 * on torch-CUDA eager (libdevice sinf / cosf, FMA-chain tiny matmuls).  A caller who wants torch-CUDA's bits computes
 * P with this flag and passes P to the warp / loss entries (per pixel torch-CUDA differs in two more places, see
 * DVF_FLAG_REF_CUDA); dvf_photo_loss_fused_pose evaluates the default profile only and rejects the flag
 * (DVF_EINVAL_DTYPE).  Verified bit for bit for DVF_ROT_EULER (the reference's default and what its callers use); under
 * DVF_ROT_QUAT torch-CUDA's quat2mat agrees on ~91 % of the entries only (profiles/ref_cuda_quat_probe.py: its norm
 * reduction sums in another order) -- quaternion poses keep the torch-CPU rounding of that step.                    */
#define DVF_ROT_REF_CUDA 0x100

/* Descriptor flags (dvf_desc.flags, dvf_loss_desc.flags); no environment variables are read anywhere. */
typedef enum dvf_flags {
  /* F.grid_sample(align_corners=True): ix = (x+1)/2*(W-1).  The reference never passes align_corners
   * (inverse_warp.py:191): torch <= 1.2, which it was written for, sampled this way; torch >= 1.3
   * defaults to False, which is this library's default (flag clear).                               */
  DVF_FLAG_ALIGN_CORNERS = 1,
  /* loss entries: gsrc maps are zero-filled by the entry (cudaMemsetAsync on the caller's stream)
   * before the launch instead of by the caller                                                    */
  DVF_FLAG_ZERO_GSRC = 2,
  /* loss entries: in-kernel replacement of the reference's per-(scale, view) NaN assertion
   * (loss_functions_sfm.py:34, one device synchronisation each): bit l*V+v of *nan_flags is set
   * when terms[l*V+v] is NaN; the caller reads the word whenever it likes                        */
  DVF_FLAG_NAN_CHECK = 4,
  /* loss entries: plain loads instead of the bulk-copy (TMA) ring of the image kernel (A/B timing) */
  DVF_FLAG_NO_TMA = 8,
  /* loss entries (image and channels-last kernels): programmatic dependent launch.  The launch carries the
   * programmatic-stream-serialisation attribute and lets its successor start early in turn, so that back-to-back
   * loss launches (micro-batches, the views of a step, consecutive steps of a captured graph) overlap: the CTAs of
   * launch i+1 start on the SM slots launch i frees while its last CTAs and its serial tail (ticket, fp64 fold,
   * pose backward) still run.  The kernel does not wait for the PREVIOUS kernel of the stream until just before it
   * exits.  CONTRACT: the previous kernel in the stream shares NO buffer with this launch -- it writes none of its
   * inputs and touches none of its outputs, and the two launches use different workspaces.                    */
  DVF_FLAG_PDL = 16,
  /* loss entries: producer glue folded into the kernel.  dvf_level.depth holds the network's
   * DISPARITY; the kernel uses depth = 1 / (disp + dvf_loss_desc.disp_eps) -- `1/(inv_depth+1e-4)` of
   * unsupervise.py:99, `1/disp` of train.py:188 (eps 0) -- evaluated as torch does (rounded add, reciprocal), and
   * dvf_level.gdepth receives d/d disparity = -g_depth * depth^2.                                              */
  DVF_FLAG_DISPARITY = 32,
  /* with DVF_FLAG_PDL: the caller promises that this launch sits in a CHAIN of such launches issued back to back
   * (consecutive steps / micro-batches of a captured graph).  The grid is then sized for overlap instead of for the
   * whole GPU -- about 64 units of 256 pixels per CTA, down to one CTA per SM: fewer CTA boundaries = fewer pieces =
   * less fixed cost, and the SM slots one launch leaves free are filled by its neighbours in the stream (C2: 51.4 ->
   * 48.7 us per step).  Sums over CTAs are folded in another order than without the flag (same 1e-7 class).      */
  DVF_FLAG_PDL_CHAINED = 64,
  /* dvf_inverse_warp_fwd only: per-pixel arithmetic of the reference run with torch-CUDA eager instead of torch-CPU
   * (companion of DVF_ROT_REF_CUDA): ATen's CUDA kernels divide by the scalar w-1 / h-1 through a multiplication
   * by its fp32 reciprocal (BinaryDivTrueKernel.cu) and form the bilinear weights as (x1 - ix) * (y1 - iy)
   * (GridSampler.cu).  dvf_inverse_warp_bwd rejects it (DVF_EUNSUPPORTED) and the loss entries do not look at it:
   * gradients and losses follow torch-CPU.                                                                    */
  DVF_FLAG_REF_CUDA = 128
} dvf_flags;

/* Image-tensor descriptor shared by the warp and loss entries. */
typedef struct dvf_desc {
  int32_t B, C, H, W;
  int32_t dtype;    /* dvf_dtype of img / warped / gout / gimg / tgt / src    */
  int32_t layout;   /* dvf_layout of those tensors                            */
  int32_t padding;  /* dvf_padding                                            */
  int32_t flags;    /* dvf_flags (DVF_FLAG_ALIGN_CORNERS)                     */
} dvf_desc;

int dvf_version(void);
const char* dvf_strerror(int status);

/* ---- pose_vec2mat / projection ------------------------------------------
 * Replaces pose_vec2mat (inverse_warp.py:141-157, euler2mat :77-114, quat2mat
 * :117-138), `intrinsics @ pose_mat` (inverse_warp.py:188) and the per-scale
 * intrinsics of loss_functions_sfm.py:20-21.
 *   vec        [n_pose,6]  (tx,ty,tz,rx,ry,rz); n_pose = B*V, b-major
 *   K, Kinv    [B,3,3] or NULL (then only posemat is produced)
 *   downscale  host array [n_levels] (img_H / level_h), n_levels may be 0
 *   posemat    [n_pose,3,4] out, nullable
 *   P          [n_levels,n_pose,3,4] out:  (K rows 0-1 / downscale) @ posemat
 *   Kinv_s     [n_levels,B,3,3] out: Kinv with columns 0-1 * downscale        */
int dvf_pose_proj_fwd(const float* vec, const float* K, const float* Kinv,
                      int32_t B, int32_t V, int32_t rotation,
                      const float* downscale, int32_t n_levels, float* posemat,
                      float* P, float* Kinv_s, void* stream);

/* Backward of the above w.r.t. vec (sums over levels), fp64 inside.
 *   gP [n_levels,n_pose,3,4] (nullable), gposemat [n_pose,3,4] (nullable)
 *   gvec [n_pose,6] out, written.                                            */
int dvf_pose_proj_bwd(const float* gP, const float* gposemat, const float* vec,
                      const float* K, int32_t B, int32_t V, int32_t rotation,
                      const float* downscale, int32_t n_levels, float* gvec,
                      void* stream);

/* ---- pixel2cam / cam2pixel (inverse_warp.py:26-40, :43-74) ---------------
 * Exposed because the reference exposes them; inverse_warp below fuses both. */
int dvf_pixel2cam(const float* depth, const float* Kinv, int32_t B, int32_t H,
                  int32_t W, float* cam /*[B,3,H,W]*/, void* stream);
int dvf_cam2pixel(const float* cam /*[B,3,H,W]*/, const float* rot /*[B,3,3] dense, nullable*/,
                  const float* tr /*[B,3] dense, nullable*/, int32_t B, int32_t H, int32_t W,
                  int32_t padding, float* grid /*[B,H,W,2]*/, void* stream);

/* Their backward passes (the reference differentiates them through autograd): d depth = (g_cam * ray).sum(1);
 * g_grid [B,H,W,2] -> g_cam [B,3,H,W], g_rot [B,3,3], g_tr [B,3] (each nullable; g_rot / g_tr need rot / tr).     */
int dvf_pixel2cam_bwd(const float* gcam /*[B,3,H,W]*/, const float* Kinv, int32_t B, int32_t H, int32_t W,
                      float* gdepth /*[B,H,W]*/, void* stream);
int dvf_cam2pixel_bwd(const float* ggrid /*[B,H,W,2]*/, const float* cam /*[B,3,H,W]*/, const float* rot, const float* tr,
                      int32_t B, int32_t H, int32_t W, int32_t padding, float* gcam, float* grot, float* gtr, void* stream);

/* ---- inverse_warp (inverse_warp.py:160-193) ------------------------------
 * warped = grid_sample(img, cam2pixel(pixel2cam(depth,Kinv), P), padding),
 * P = K @ pose_vec2mat(pose) [B,3,4].  valid (nullable, uint8 [B,H,W]) is the
 * value-based mask of loss_functions.py:11: any_c(warped_c != 0).             */
int dvf_inverse_warp_fwd(const dvf_desc* d, const void* img, const float* depth,
                         const float* P, const float* Kinv, void* warped,
                         uint8_t* valid, void* stream);

size_t dvf_inverse_warp_bwd_workspace_bytes(const dvf_desc* d);

/* gout = dL/dwarped.  gdepth [B,H,W] and gP [B,3,4] are written; gimg
 * (nullable, same dtype/layout as img, fp32 only) is ACCUMULATED into and must
 * be zero-filled by the caller.  With gimg == NULL, C == 3, H*W % 4 == 0 and
 * 16-byte aligned gout / img / depth the call runs on the image kernel of the
 * fused loss (bulk-copy ring, balanced persistent grid); d depth is bit-identical
 * on either route and the workspace serves both.                              */
int dvf_inverse_warp_bwd(const dvf_desc* d, const void* gout, const void* img,
                         const float* depth, const float* P, const float* Kinv,
                         float* gdepth, float* gP, void* gimg, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- fused masked reconstruction loss, forward + backward in one pass ----
 * Replaces photometric_reconstruction_loss of loss_functions.py:7-20 (V=2,
 * one level, images or feature maps), loss_functions_sfm.py:9-46 (V refs, n
 * levels, explainability masks) and loss_function_sfm_old.py:7-46.
 * For every level l and view v:
 *   warped = inverse_warp(src[v], depth, P[:,v]);  valid = any_c(warped != 0)
 *   terms[l*V+v] = mean_{B,C,H,W} | (tgt - warped) * valid [* expl[:,v]] |
 * and, for upstream d(sum of terms) = 1, the gradients w.r.t. depth (summed
 * over views), P, expl, src and tgt.  Any gradient pointer may be NULL.      */
typedef struct dvf_level {
  int32_t H, W;
  const float* depth;               /* [B,H,W]                                */
  const void* tgt;                  /* [B,C,H,W] or [B,H,W,C]                 */
  const void* src[DVF_MAX_VIEWS];   /* like tgt                               */
  const float* expl;                /* NULL or base of channel 0: element (b,v,y,x) at expl[b*expl_bstride + v*H*W + y*W + x] */
  int64_t expl_bstride;
  const float* P;                   /* [B,V,3,4]                              */
  const float* Kinv;                /* [B,3,3]                                */
  float* gdepth;                    /* [B,H,W] written                        */
  float* gexpl;                     /* [B,V,H,W] dense, written               */
  void* gsrc[DVF_MAX_VIEWS];        /* fp32, layout of src; ACCUMULATED (see DVF_FLAG_ZERO_GSRC) */
  void* gtgt;                       /* dvf_loss_desc.grad_dtype, layout of tgt; written */
  float* gP;                        /* [B,V,3,4] written                      */
} dvf_level;

typedef struct dvf_loss_desc {
  int32_t B, C, V, n_levels;
  int32_t dtype, layout, padding;
  int32_t flags;          /* dvf_flags                                                              */
  int32_t mean_batch;     /* batch size in the denominator of the means; 0 = B.  A rank that holds B
                             of the mean_batch images of a sharded batch passes the global size and
                             gets its share of the global loss and gradients (no collective needed) */
  int32_t grad_dtype;     /* dvf_dtype of gtgt: DVF_F32, or DVF_BF16 for NHWC bf16 maps (gsrc is accumulated
                             and therefore always fp32)                                            */
  int32_t piece_overhead; /* tuning: fixed cost of an (image, level) piece in 256-px units; <= 0 = default */
  int32_t ctas_per_sm;    /* tuning: resident CTAs per SM of the balanced kernels' grid; 0 = automatic, < 0 = a grid of
                             exactly -ctas_per_sm CTAs                                               */
  const float* upstream;  /* device scalar g = d(total)/d(sum of terms): every gradient is scaled by it
                             (terms are not); NULL = 1                                              */
  int32_t* nan_flags;     /* device word for DVF_FLAG_NAN_CHECK, OR-ed into (never cleared); nullable */
  /* Exchange of the loss terms between the GPUs of a node, fused into the kernel's epilogue (multi-GPU logging of
   * the loss): the CTA that finishes a level stores its terms into the buffer of EVERY peer over
   * NVLink (peer-to-peer mapped memory), so each rank ends up with all ranks' terms -- an all-gather without a
   * collective launch; the global term is the sum over the rows.  peer_terms: HOST array of n_peers device
   * pointers, entry q = rank q's buffer [n_peers][n_levels*V] floats as mapped in THIS process; this launch writes
   * row peer_rank of every buffer.  Visible to a peer once this kernel has completed (stream / event / barrier
   * order).  n_peers = 0: no exchange.                                                                         */
  int32_t n_peers, peer_rank;
  float* const* peer_terms;
  float disp_eps;         /* DVF_FLAG_DISPARITY: depth = 1 / (disp + disp_eps)                                  */
  float img_scale;        /* images (C = 3, NCHW) are multiplied by this on load: tgt and src hold the RAW images and
                             the loss is that of img_scale * img (unsupervise.py:101 passes 0.004 * img); 0 or 1 = none */
} dvf_loss_desc;
#define DVF_MAX_PEERS 8

size_t dvf_photo_loss_workspace_bytes(const dvf_loss_desc* d, const dvf_level* levels);

/* terms: device float [n_levels*V], written (deterministic reduction order). */
int dvf_photo_loss_fused(const dvf_loss_desc* d, const dvf_level* levels,
                         float* terms, void* workspace, size_t workspace_bytes,
                         void* stream);

/* Same launch with the pose chain folded in: the kernel derives P = (K rows 0-1 / downscale) @ pose_vec2mat(vec)
 * and K^-1_s itself (levels[l].P / .Kinv are ignored, may be NULL) and, when gvec is given, finishes with the
 * backward of pose_vec2mat -> d(sum of terms)/d vec, summed over levels.  One launch per training step.      */
typedef struct dvf_pose_args {
  const float* vec;        /* [B,V,6] (tx,ty,tz,rx,ry,rz)                                     */
  const float* K;          /* [B,3,3]                                                         */
  const float* Kinv;       /* [B,3,3]                                                         */
  const float* downscale;  /* HOST array [n_levels]: image H / level h (loss_functions_sfm.py:16) */
  int32_t rotation;        /* dvf_rotation                                                    */
  int32_t reserved;
  float* gvec;             /* [B,V,6] written, or NULL                                        */
} dvf_pose_args;

int dvf_photo_loss_fused_pose(const dvf_loss_desc* d, const dvf_level* levels, const dvf_pose_args* pose,
                              float* terms, void* workspace, size_t workspace_bytes, void* stream);

/* ---- neighbours of the path ----------------------------
 * F.interpolate(img,(h,w),mode='area') for the integer factors 2,4,8
 * (loss_functions_sfm.py:18-19): one pass over img writes all levels.        */
int dvf_area_pyramid(const float* img /*[BC,H,W]*/, int32_t BC, int32_t H, int32_t W,
                     int32_t n_out, float* const* outs /*host array of n_out device ptrs: /2,/4,/8*/,
                     void* stream);

/* Same operator for one arbitrary output size (h,w) -- adaptive average pooling windows
 * [floor(o*I/O), ceil((o+1)*I/O)), as F.interpolate(mode='area') uses.               */
int dvf_area_downsample(const float* img /*[BC,H,W]*/, int32_t BC, int32_t H, int32_t W,
                        int32_t h, int32_t w, float* out /*[BC,h,w]*/, void* stream);

/* Layout change of feature maps, per image a [R,S] -> [S,R] transpose: dense NCHW -> channels-last with R = C,
 * S = H*W (the reference's FeatExtractor hands NCHW maps to the loss, unsupervise.py:104-109; the feature-loss
 * kernel reads channels-last), and back with R = H*W, S = C.  elem_bytes 4 (fp32) or 2 (bf16).               */
int dvf_transpose_planes(const void* src /*[B,R,S]*/, void* dst /*[B,S,R]*/, int32_t B, int32_t R, int32_t S,
                         int32_t elem_bytes, void* stream);

/* The same transpose for fp32 with every element multiplied by *scale (device scalar, nullable = 1): the gradient maps
 * of the feature loss go back to dense NCHW and take the upstream gradient in one pass.                              */
int dvf_transpose_planes_scaled(const float* src /*[B,R,S]*/, float* dst /*[B,S,R]*/, int32_t B, int32_t R, int32_t S,
                                const float* scale, void* stream);

/* ---- regularisers next to the path -----------------------
 * smooth_loss (loss_functions.py:23-41, loss_functions_sfm.py:59-77) and explainability_loss
 * (loss_functions_sfm.py:49-56), every scale in ONE launch, value and gradient together.
 *   x      [B,H,W] fp32 (maps with several channels: fold the channels into B)
 *   g      same shape, written: d(out)/dx for upstream 1 (nullable)
 *   weight factor of this level in the sum (smooth_loss: 1/scale_factor^level; explainability: 1)
 *   out    device float[1]: sum_l weight_l * loss_l                                    */
typedef struct dvf_reg_level {
  const float* x;
  float* g;
  int32_t B, H, W;
  float weight;
} dvf_reg_level;

size_t dvf_reg_workspace_bytes(const dvf_reg_level* levels, int32_t n_levels);
int dvf_smooth_loss(const dvf_reg_level* levels, int32_t n_levels, float* out, void* workspace,
                    size_t workspace_bytes, void* stream);
int dvf_explainability_loss(const dvf_reg_level* levels, int32_t n_levels, float* out, void* workspace,
                            size_t workspace_bytes, void* stream);

/* ---- se(3) -> SE(3) exponential map ------------------------------
 * Replaces SE3_Generator_KITTI (pytorch_version/se3_generate.py:7-103; copies in model.py:33-131,
 * fixmodel.py:10-108, caffe/python/pygeometry.py:6-115).  in [B,6] fp32 = (w, u); out [B,4,4] fp64 =
 * [[R, R u],[0,1]]; backward: gout [B,4,4] fp64 -> gin [B,6] fp32.                          */
int dvf_se3_exp_fwd(const float* in, int32_t B, double* out, void* stream);
int dvf_se3_exp_bwd(const float* in, const double* gout, int32_t B, float* gin, void* stream);

/* ---- Caffe-convention layers ---------------------------------------
 * GeoTransform / PinHole / InverseWarping / AbsLoss of caffe/src/caffe/layers/{geometry_transformation,
 * pin_hole_layer,inverse_warping_layer,abs_loss_layer}.cu (transliterated, non-functional, in
 * pytorch_version/geo_transform.py:6-125; called from unsupervise_dvo.py:95-122).  Pixel-space sample positions,
 * K [N,4] = (fx,fy,cx,cy), T [N,16] = row-major 4x4, pts [N,3,H,W], coords [N,2,H,W], images [N,C,H,W], fp32.
 * Backward entries WRITE their outputs (the small T/K/img accumulators are zeroed inside); any output may be NULL. */
int dvf_caffe_geo_fwd(const float* depth, const float* T, const float* K, int32_t N, int32_t H, int32_t W, float* pts, void* stream);
int dvf_caffe_geo_bwd(const float* top_diff, const float* depth, const float* T, const float* K, int32_t N, int32_t H, int32_t W,
                      float* depth_diff, float* T_diff, float* K_diff, void* stream);
int dvf_caffe_pinhole_fwd(const float* pts, const float* K, int32_t N, int32_t H, int32_t W, float* coords, void* stream);
int dvf_caffe_pinhole_bwd(const float* coords_diff, const float* pts, const float* K, int32_t N, int32_t H, int32_t W,
                          float* pts_diff, float* K_diff, void* stream);
int dvf_caffe_warp_fwd(const float* img, const float* coords, int32_t N, int32_t C, int32_t H, int32_t W, float* out, void* stream);
int dvf_caffe_warp_bwd(const float* top_diff, const float* img, const float* coords, int32_t N, int32_t C, int32_t H, int32_t W,
                       float* img_diff, float* coords_diff, void* stream);
/* loss[0] = sum|a-b| / num;  ga = weight/num * ((d>0)-(d<=0)), gb = -ga (nullable); workspace: >= 8 bytes, 8-aligned */
int dvf_caffe_abs_loss(const float* a, const float* b, uint64_t count, int32_t num, float weight, float* loss, float* ga,
                       float* gb, void* workspace, void* stream);

/* Edge-aware smoothness of the Caffe graphs (experiments/depth/train.prototxt:4022-4234, fillers caffe/include/caffe/
 * filler.hpp:266-315): gx = exp(-0.33 * sum_c |EdgeX(img_c)|), dx = gx * EdgeX(inv_depth) over the (H-2)x(W-2) valid
 * window origins, likewise y; loss[0] = sum|dx| / N, loss[1] = sum|dy| / N (the two AbsLoss tops, un-weighted);
 * ginv [N,1,H,W] (nullable) = d(weight * (loss[0] + loss[1])) / d inv_depth (the prototxt uses loss_weight 10 for both).
 * img [N,3,H,W], inv_depth [N,1,H,W]; workspace: >= 16 bytes, 8-aligned.                                        */
int dvf_caffe_edge_smooth_loss(const float* img, const float* inv_depth, int32_t N, int32_t H, int32_t W, float weight,
                               float* loss /*[2]*/, float* ginv, void* workspace, void* stream);

/* ---- SSIM reconstruction term (north_star "masked photometric (L1/SSIM)") --------------------------------------
 * NOT in the reference (no SSIM anywhere in the checkout): new functionality, PARITY UNPINNED, specified in
 * csrc/dvf_ssim.cu -- 3x3 average-pool SSIM without padding, C1 = 0.01^2, C2 = 0.03^2, l = clamp((1-SSIM)/2, 0, 1),
 * loss = sum(m*l) / (B*C*(H-2)*(W-2)) with m = AND of `valid` (nullable, uint8 [B,H,W], dvf_inverse_warp_fwd's mask)
 * over the window.  x, y [B,C,H,W] fp32; loss: device float; gy (nullable) = d loss / d y, written;
 * workspace: >= 8 bytes, 8-aligned.                                                                              */
int dvf_ssim_loss(const float* x, const float* y, const uint8_t* valid, int32_t B, int32_t C, int32_t H, int32_t W,
                  float* loss, float* gy, void* workspace, void* stream);

/* ---- torch.sin / torch.cos of fp32 values exactly as torch-CPU evaluates them ------------------
 * (reference: inverse_warp.py:89-91,98-99,105-106 -- euler2mat's torch.cos / torch.sin, which on the
 * CPU run MKL's VML in high-accuracy mode).  Bit-identical to torch 2.11 CPU for |x| <= 10000; larger
 * arguments use CUDA's sinf / cosf.  The pose entries above use the same routine.  Either output may
 * be NULL.                                                                                        */
int dvf_torch_sincos(const float* x, int64_t n, float* sin_out, float* cos_out, void* stream);

/* ---- diagnostics -----------------------------------------------------------
 * Compares the shared-reciprocal IEEE division of the coordinate chain with
 * __fdiv_rn on n pseudo-random operand pairs (mode 0: float divisors, mode 1:
 * integer divisors with a host-style reciprocal).  *mismatches: device u64,
 * zeroed by the caller, receives the number of differing quotients.           */
int dvf_selftest_fast_div(uint64_t seed, uint64_t n, int32_t mode, unsigned long long* mismatches, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVF_B200_H_ */
