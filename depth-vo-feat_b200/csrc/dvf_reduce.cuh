// dvf_reduce.cuh -- deterministic CTA / cross-CTA reductions used by the backward kernels.
//
// The reference reduces dL/dP with a [B,3,HW]@[B,HW,3] GEMM and a sum over HW (autograd of
// inverse_warp.py:55-60); the Caffe-era kernels used 12 global atomics per pixel
// (caffe/src/caffe/layers/geometry_transformation.cu:128-172).  Here: per-thread fp32
// accumulation, a 16-slot shuffle butterfly per warp, one smem fold per CTA, and a fixed-order
// fp64 fold of the CTA partials by whichever CTA of the image finishes last (atomic ticket,
// nobody waits).  Results are bit-reproducible run to run.
#pragma once
#include "dvf_internal.h"

namespace dvf {

// fold a[0..15] over the 32 lanes: afterwards lane L holds, in a[0], the warp-wide sum of
// slot ((L>>1) & 15) -- 16 shuffles in total.
__device__ __forceinline__ float butterfly16(float (&a)[kRedSlots], int lane) {
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? a[i] : a[i + 8];
      const float keep = up ? a[i + 8] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? a[i] : a[i + 4];
      const float keep = up ? a[i + 4] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? a[i] : a[i + 2];
      const float keep = up ? a[i + 2] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = lane & 2;
    const float send = up ? a[0] : a[1];
    const float keep = up ? a[1] : a[0];
    a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
  return a[0];
}
// slot held by a lane after butterfly16: bit4 -> +8, bit3 -> +4, bit2 -> +2, bit1 -> +1
__device__ __forceinline__ int butterfly_slot(int lane) { return (lane >> 1) & 15; }

__device__ __forceinline__ float ldcg_f(const float* p) { return __ldcg(p); }

// deterministic sum of n values x[k*stride] by the 8 lanes of a group (fixed order)
__device__ __forceinline__ double group8_sum(const float* x, int n, int stride, int sub) {
  double s = 0.0;
  for (int k = sub; k < n; k += 8) s += (double)ldcg_f(x + (size_t)k * stride);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  return s;
}

}  // namespace dvf
