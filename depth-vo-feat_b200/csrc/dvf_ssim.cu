// dvf_ssim.cu -- SSIM reconstruction term (forward + backward in one launch).
//
// BASELINE.json's north_star names a "masked photometric (L1/SSIM)" loss; the reference itself contains no SSIM
// (`grep -ri ssim` over the checkout finds nothing), so this is NEW functionality specified here, PARITY UNPINNED
// (there is nothing to pin it to; tests compare with a torch fp32 restatement of the same definition).  Definition --
// the 3x3 average-pool SSIM of the monocular-depth literature (Godard et al.), C1 = 0.01^2, C2 = 0.03^2, no padding:
//   mu_x = avg3(x), mu_y = avg3(y), s_x = avg3(x^2) - mu_x^2, s_y = avg3(y^2) - mu_y^2, s_xy = avg3(x*y) - mu_x*mu_y
//   SSIM = (2 mu_x mu_y + C1)(2 s_xy + C2) / ((mu_x^2 + mu_y^2 + C1)(s_x + s_y + C2))
//   l    = clamp((1 - SSIM) / 2, 0, 1)                    per channel and 3x3 window, (H-2) x (W-2) windows
//   loss = sum(m * l) / (B * C * (H-2) * (W-2)),          m = 1, or AND of `valid` over the window's 9 pixels
// x = target image, y = warped image (dvf_inverse_warp_fwd, which also provides `valid`); gy = d loss / d y feeds
// dvf_inverse_warp_bwd.  HBM-bound stream: 8 B read + 4 B written per element; one CTA handles a 32 x 16 tile of one
// (image, channel) plane with a 2-pixel halo staged in shared memory: every window's statistics are evaluated once,
// every pixel gathers the <= 9 windows it belongs to.
#include "dvf_internal.h"

namespace dvf {

constexpr int kSsimTW = 32, kSsimTH = 16;                 // tile of pixels whose gradient this CTA writes
constexpr int kSsimPW = kSsimTW + 4, kSsimPH = kSsimTH + 4;   // pixels staged (halo 2)
constexpr int kSsimWW = kSsimTW + 2, kSsimWH = kSsimTH + 2;   // window origins evaluated (rows r0-2.., cols c0-2..)
constexpr float kSsimC1 = 0.01f * 0.01f, kSsimC2 = 0.03f * 0.03f;

__global__ void __launch_bounds__(kSsimTW* kSsimTH) ssim_loss_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                    const uint8_t* __restrict__ valid, int C, int H, int W,
                                                                    float inv_n, float* __restrict__ gy, double* __restrict__ acc) {
  __shared__ float s_x[kSsimPH][kSsimPW], s_y[kSsimPH][kSsimPW];
  __shared__ uint8_t s_v[kSsimPH][kSsimPW];
  __shared__ float s_k[3][kSsimWH][kSsimWW];   // per window: k*dS/dmu_y, k*dS/dE[y^2], k*dS/dE[xy]  (k = -m/(2N) inside the clamp)
  const int plane = blockIdx.z, b = plane / C;
  const int r0 = blockIdx.y * kSsimTH, c0 = blockIdx.x * kSsimTW;
  const int tid = threadIdx.y * kSsimTW + threadIdx.x;
  const float* xp = x + (size_t)plane * H * W;
  const float* yp = y + (size_t)plane * H * W;
  const uint8_t* vp = valid ? valid + (size_t)b * H * W : nullptr;
  for (int i = tid; i < kSsimPH * kSsimPW; i += kSsimTW * kSsimTH) {
    const int pr = i / kSsimPW, pc = i - pr * kSsimPW;
    const int r = r0 - 2 + pr, c = c0 - 2 + pc;
    const bool in = r >= 0 && c >= 0 && r < H && c < W;
    s_x[pr][pc] = in ? __ldg(xp + (size_t)r * W + c) : 0.0f;
    s_y[pr][pc] = in ? __ldg(yp + (size_t)r * W + c) : 0.0f;
    s_v[pr][pc] = in ? (vp ? __ldg(vp + (size_t)r * W + c) : (uint8_t)1) : (uint8_t)0;
  }
  __syncthreads();
  double lsum = 0.0;
  for (int i = tid; i < kSsimWH * kSsimWW; i += kSsimTW * kSsimTH) {
    const int wr = i / kSsimWW, wc = i - wr * kSsimWW;
    const int r = r0 - 2 + wr, c = c0 - 2 + wc;   // window origin in the image; its pixels are staged at [wr..wr+2][wc..wc+2]
    float k1 = 0.0f, k2 = 0.0f, k3 = 0.0f;
    if (r >= 0 && c >= 0 && r < H - 2 && c < W - 2) {
      float sx = 0.0f, sy = 0.0f;
      bool m = true;
#pragma unroll
      for (int dr = 0; dr < 3; ++dr)
#pragma unroll
        for (int dc = 0; dc < 3; ++dc) {
          sx += s_x[wr + dr][wc + dc];
          sy += s_y[wr + dr][wc + dc];
          m = m && s_v[wr + dr][wc + dc] != 0;
        }
      const float ninth = 1.0f / 9.0f;
      const float mx = sx * ninth, my = sy * ninth;
      // centred second moments (E[x^2] - mu^2 loses half the digits on smooth image patches)
      float sxx = 0.0f, syy = 0.0f, sxy = 0.0f;
#pragma unroll
      for (int dr = 0; dr < 3; ++dr)
#pragma unroll
        for (int dc = 0; dc < 3; ++dc) {
          const float a = s_x[wr + dr][wc + dc] - mx, bb = s_y[wr + dr][wc + dc] - my;
          sxx += a * a; syy += bb * bb; sxy += a * bb;
        }
      const float vx = sxx * ninth, vy = syy * ninth, cxy = sxy * ninth;
      const float A1 = 2.0f * mx * my + kSsimC1, A2 = 2.0f * cxy + kSsimC2;
      const float B1 = mx * mx + my * my + kSsimC1, B2 = vx + vy + kSsimC2;
      const float n = A1 * A2, d = B1 * B2;
      const float S = n / d;
      const float l = 0.5f * (1.0f - S);
      if (m) {
        const float lc = fminf(fmaxf(l, 0.0f), 1.0f);
        // each window is counted by the CTA whose pixel tile contains its origin
        if (wr >= 2 && wc >= 2) lsum += (double)lc;
        if (l > 0.0f && l < 1.0f) {
          const float k = -0.5f * inv_n;
          const float rd = 1.0f / d;
          const float dS_dmy = ((2.0f * mx * A2 - 2.0f * mx * A1) - S * (2.0f * my * B2 - 2.0f * my * B1)) * rd;
          const float dS_dEyy = -S * B1 * rd;
          const float dS_dExy = 2.0f * A1 * rd;
          k1 = k * dS_dmy * ninth;
          k2 = k * dS_dEyy * ninth;
          k3 = k * dS_dExy * ninth;
        }
      }
    }
    s_k[0][wr][wc] = k1;
    s_k[1][wr][wc] = k2;
    s_k[2][wr][wc] = k3;
  }
  __syncthreads();
  const int r = r0 + threadIdx.y, c = c0 + threadIdx.x;
  if (gy && r < H && c < W) {
    // pixel (r, c) = staged [ty+2][tx+2] belongs to the windows with origins (r-2..r, c-2..c) = s_k[.][ty..ty+2][tx..tx+2]
    float a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
    for (int dr = 0; dr < 3; ++dr)
#pragma unroll
      for (int dc = 0; dc < 3; ++dc) {
        a1 += s_k[0][threadIdx.y + dr][threadIdx.x + dc];
        a2 += s_k[1][threadIdx.y + dr][threadIdx.x + dc];
        a3 += s_k[2][threadIdx.y + dr][threadIdx.x + dc];
      }
    const float xv = s_x[threadIdx.y + 2][threadIdx.x + 2], yv = s_y[threadIdx.y + 2][threadIdx.x + 2];
    gy[(size_t)plane * H * W + (size_t)r * W + c] = a1 + 2.0f * yv * a2 + xv * a3;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  __shared__ double s_l[kSsimTW * kSsimTH / 32];
  if ((tid & 31) == 0) s_l[tid >> 5] = lsum;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kSsimTW * kSsimTH / 32; ++w) t += s_l[w];
    atomicAdd(acc, t);
  }
}
__global__ void ssim_finish(const double* acc, double n, float* out) { *out = (float)(*acc / n); }

}  // namespace dvf

using namespace dvf;

DVF_EXPORT int dvf_ssim_loss(const float* x, const float* y, const uint8_t* valid, int32_t B, int32_t C, int32_t H, int32_t W,
                             float* loss, float* gy, void* workspace, void* stream) {
  if (!x || !y || !loss || !workspace) return DVF_EINVAL_NULL;
  if (B <= 0 || C <= 0 || H < 3 || W < 3 || (long long)B * C > 65535 || (long long)H * W >= (1ll << 30)) return DVF_EINVAL_SHAPE;
  if (!aligned(workspace, 8)) return DVF_EINVAL_ALIGN;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  double* acc = static_cast<double*>(workspace);
  cudaMemsetAsync(acc, 0, sizeof(double), cs);
  const double n = (double)B * C * (H - 2) * (W - 2);
  dim3 grid((W + kSsimTW - 1) / kSsimTW, (H + kSsimTH - 1) / kSsimTH, B * C), block(kSsimTW, kSsimTH);
  ssim_loss_kernel<<<grid, block, 0, cs>>>(x, y, valid, C, H, W, (float)(1.0 / n), gy, acc);
  ssim_finish<<<1, 1, 0, cs>>>(acc, n, loss);
  return launch_status();
}
