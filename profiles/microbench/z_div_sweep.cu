// Can the quotients by the per-pixel depth Z drop their second correction step too?  Exhaustive over MANTISSAS (the
// sequences are scale-invariant while everything stays normal): every divisor b in [1, 2) x every numerator a in [1, 2)
// (2^46 pairs; the quotient covers (0.5, 2), i.e. both binade cases), candidate sequences against __fdiv_rn:
//   V1: y = rcp.approx + one Newton step (dvf_math.cuh: rcp_refined),  q = a*y, q += fma(-b, q, a) * y       (one step)
//   V2: y = rcp.approx + two Newton steps,                              same one-step quotient
//   V0: y = rcp.approx alone,                                           same one-step quotient
// Prints the number of mismatching pairs per variant.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o z_div_sweep z_div_sweep.cu
// usage: z_div_sweep [first b mantissa, hex] [number of b mantissas, hex]      (default: all 2^23)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__global__ void sweep(uint32_t b_first, unsigned long long* out) {
  const uint32_t mb = b_first + blockIdx.x * blockDim.x + threadIdx.x;
  const float b = __uint_as_float(0x3f800000u | mb), nb = -b;
  float y1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(b));
  const float y0 = y1;   // V0: the raw approximation, no Newton step
  y1 = __fmaf_rn(y1, __fmaf_rn(nb, y1, 1.0f), y1);
  const float y2 = __fmaf_rn(y1, __fmaf_rn(nb, y1, 1.0f), y1);
  unsigned long long bad1 = 0, bad2 = 0, bad_y = 0, bad0 = 0, badv0 = 0;
  if (__float_as_uint(y2) != __float_as_uint(__frcp_rn(b))) bad_y = 1;
#pragma unroll 4
  for (uint32_t ma = 0; ma < (1u << 23); ++ma) {
    const float a = __uint_as_float(0x3f800000u | ma);
    const uint32_t ref = __float_as_uint(__fdiv_rn(a, b));
    float q = __fmul_rn(a, y1);
    bad0 += (__float_as_uint(q) != ref);   // control: the uncorrected quotient must mismatch often
    q = __fmaf_rn(__fmaf_rn(nb, q, a), y1, q);
    bad1 += (__float_as_uint(q) != ref);
    q = __fmul_rn(a, y2);
    q = __fmaf_rn(__fmaf_rn(nb, q, a), y2, q);
    bad2 += (__float_as_uint(q) != ref);
    q = __fmul_rn(a, y0);
    q = __fmaf_rn(__fmaf_rn(nb, q, a), y0, q);
    badv0 += (__float_as_uint(q) != ref);
  }
  if (bad1) atomicAdd(out, bad1);
  if (bad2) atomicAdd(out + 1, bad2);
  if (bad_y) atomicAdd(out + 2, bad_y);
  if (bad0) atomicAdd(out + 3, bad0);
  if (badv0) atomicAdd(out + 4, badv0);
}
int main(int argc, char** argv) {
  const uint32_t first = argc > 1 ? (uint32_t)strtoul(argv[1], nullptr, 16) : 0u;
  const uint32_t count = argc > 2 ? (uint32_t)strtoul(argv[2], nullptr, 16) : (1u << 23);
  unsigned long long* dev;
  cudaMalloc(&dev, 40);
  cudaMemset(dev, 0, 40);
  const uint32_t per_launch = 1u << 17;   // b mantissas per launch (keeps launches to a few seconds)
  for (uint32_t done = 0; done < count; done += per_launch) {
    const uint32_t n = count - done < per_launch ? count - done : per_launch;
    sweep<<<(n + 127) / 128, 128>>>(first + done, dev);
    if ((done / per_launch) % 8 == 7 || done + per_launch >= count) {
      unsigned long long h[5];
      if (cudaMemcpy(h, dev, 40, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
      printf("b mantissas %06x..%06x x all a: one Newton step %llu mismatches, two %llu; two-step reciprocals != 1/b correctly rounded: %llu; control (no correction step): %llu mismatches; V0 (raw rcp.approx, one step): %llu\n",
             first, first + done + n - 1, h[0], h[1], h[2], h[3], h[4]);
      fflush(stdout);
    }
  }
  return 0;
}
