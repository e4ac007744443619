// Microbenchmark: throughput of packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2, sm_100a) vs scalar FFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
  const float2 m = make_float2(s, s * 0.999f), c = make_float2(1e-7f, -1e-7f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // scalar: 2 FFMA per element pair
        a[i].x = __fmaf_rn(a[i].x, m.x, c.x);
        a[i].y = __fmaf_rn(a[i].y, m.y, c.y);
      } else if (MODE == 1) {  // packed FFMA2
        a[i] = __ffma2_rn(a[i], m, c);
      } else if (MODE == 2) {  // packed, mixed with an integer op per FFMA2 (does packing free issue slots?)
        a[i] = __ffma2_rn(a[i], m, c);
        asm volatile("" ::: "memory");
      } else if (MODE == 3) {  // FADD2 + FMUL2
        a[i] = __fadd2_rn(__fmul2_rn(a[i], m), c);
      }
    }
  }
  float r = 0;
  for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// interleave: N packed/scalar FMAs with integer ALU work, to see whether FFMA2 relieves the issue port
template <int MODE>
__global__ void kmix(float* out, int iters, float s, int q) {
  float2 a[4];
  int z[4];
  for (int i = 0; i < 4; ++i) { a[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i); z[i] = threadIdx.x + i; }
  const float2 m = make_float2(s, s * 0.999f), c = make_float2(1e-7f, -1e-7f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (MODE == 0) { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); a[i].y = __fmaf_rn(a[i].y, m.y, c.y); }
      else a[i] = __ffma2_rn(a[i], m, c);
      z[i] = (z[i] ^ q) + (z[i] >> 3);   // 2-3 ALU ops
    }
  }
  float r = 0;
  for (int i = 0; i < 4; ++i) r += a[i].x + a[i].y + z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  float* out; cudaMalloc(&out, blocks * threads * 4);
  const double fma_per_launch = (double)blocks * threads * iters * 16;
  float t0 = timeit([&] { k<0><<<blocks, threads>>>(out, iters, 0.9999f); });
  float t1 = timeit([&] { k<1><<<blocks, threads>>>(out, iters, 0.9999f); });
  float t3 = timeit([&] { k<3><<<blocks, threads>>>(out, iters, 0.9999f); });
  printf("scalar FFMA : %.3f ms  %.2f TFMA/s\n", t0, fma_per_launch / t0 / 1e9);
  printf("FFMA2       : %.3f ms  %.2f TFMA/s\n", t1, fma_per_launch / t1 / 1e9);
  printf("FMUL2+FADD2 : %.3f ms  %.2f Tpair-op/s\n", t3, fma_per_launch / t3 / 1e9);
  const double fma_mix = (double)blocks * threads * iters * 8;
  float m0 = timeit([&] { kmix<0><<<blocks, threads>>>(out, iters, 0.9999f, 5); });
  float m1 = timeit([&] { kmix<1><<<blocks, threads>>>(out, iters, 0.9999f, 5); });
  printf("mix scalar  : %.3f ms  %.2f TFMA/s\n", m0, fma_mix / m0 / 1e9);
  printf("mix FFMA2   : %.3f ms  %.2f TFMA/s\n", m1, fma_mix / m1 / 1e9);
  return 0;
}
