// dvf_internal.h -- host-side helpers shared by the translation units of libdvf_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/dvf_b200.h"

#define DVF_EXPORT extern "C" __attribute__((visibility("default")))

namespace dvf {

constexpr int kThreads = 256;   // threads per CTA of the per-pixel kernels
constexpr int kRedSlots = 16;   // 12 dP entries + loss term + 3 spare, per view

inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DVF_OK : (int)e;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline int num_sms() {
  static int n = 0;   // read-only after first query; benign race
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      n = v;
    else
      return 148;
  }
  return n;
}

// dvf_inverse_warp_bwd without d img, run by the image kernel of the fused loss (dvf_loss.cu): TMA ring, balanced split
bool warp_bwd_fused_ok(const dvf_desc* d, const void* gout, const void* img, const float* depth);
size_t warp_bwd_fused_workspace_bytes(const dvf_desc* d);
int warp_bwd_fused(const dvf_desc* d, const void* gout, const void* img, const float* depth, const float* P, const float* Kinv,
                   float* gdepth, float* gP, void* workspace, size_t workspace_bytes, void* stream);

}  // namespace dvf
