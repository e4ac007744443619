// Exhaustive check of the ONE-correction-step division by a constant: for every divisor d in [1, dmax] and EVERY fp32
// numerator a with |a| <= 2^100 (zero, subnormals, both signs),  q0 = RN(a*r), e = fma(-b, q0, a), q1 = fma(e, r, q0)
// with r = (float)(1.0 / (double)d) -- the reciprocal dvf_math.cuh: make_geo hands the kernels -- against __fdiv_rn(a, b).
// Counted separately: numerators with |a| in [2^-100, 2^100] (the range the coordinate chain guarantees: gradients are
// >= 1e-30 in magnitude or exactly zero) and the tiny ones below it, where the remainder a - b*q is no longer exactly
// representable and the two-step sequence of dvf_math.cuh: div_by rounds differently from IEEE as well (third counter).
// Signed zeros are ignored (both sequences return +0 for -0 / b).
// Prints the divisors with at least one mismatch in the guaranteed range (the kernels keep the two-step sequence for those).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o const_div_sweep const_div_sweep.cu
// usage: const_div_sweep [dmin] [dmax] [min |a| bits, hex] [max |a| bits, hex]   (defaults 0 and 71800000 = 2^100)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
__global__ void sweep(float b, float r, uint32_t mag_lo, uint32_t mag_hi, unsigned long long* out /* [0] mismatches in range, [1] first bad numerator bits + 1,
                                                                    [2] one-step mismatches below 2^-100, [3] two-step ones */) {
  const uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) * 256u;
  unsigned long long bad = 0, tiny1 = 0, tiny2 = 0;
  uint32_t first = 0;
  const float nb = -b;
  for (uint32_t k = 0; k < 256u; ++k) {
    const uint32_t bits = base + k;
    if ((bits & 0x7fffffffu) > mag_hi || (bits & 0x7fffffffu) < mag_lo) continue;   // outside the requested magnitudes, inf, NaN
    if ((bits & 0x7fffffffu) == 0u) continue;           // +-0
    const float a = __uint_as_float(bits);
    float q = __fmul_rn(a, r);
    float e = __fmaf_rn(nb, q, a);
    q = __fmaf_rn(e, r, q);
    const float ref = __fdiv_rn(a, b);
    const bool in_range = (bits & 0x7fffffffu) >= 0x0d800000u;   // |a| >= 2^-100
    if (__float_as_uint(q) != __float_as_uint(ref)) {
      if (in_range) {
        if (!bad) first = bits;
        ++bad;
      } else {
        ++tiny1;
      }
    }
    if (!in_range) {
      e = __fmaf_rn(nb, q, a);
      q = __fmaf_rn(e, r, q);
      if (__float_as_uint(q) != __float_as_uint(ref)) ++tiny2;
    }
  }
  if (bad) {
    atomicAdd(out, bad);
    atomicMax(out + 1, (unsigned long long)first + 1ull);
  }
  if (tiny1) atomicAdd(out + 2, tiny1);
  if (tiny2) atomicAdd(out + 3, tiny2);
}
int main(int argc, char** argv) {
  const int dmin = argc > 1 ? atoi(argv[1]) : 1, dmax = argc > 2 ? atoi(argv[2]) : 8191;
  const uint32_t mag_lo = argc > 3 ? (uint32_t)strtoul(argv[3], nullptr, 16) : 0u, mag_hi = argc > 4 ? (uint32_t)strtoul(argv[4], nullptr, 16) : 0x71800000u;
  unsigned long long* dev;
  cudaMalloc(&dev, 32 * (size_t)(dmax + 1));
  cudaMemset(dev, 0, 32 * (size_t)(dmax + 1));
  for (int d = dmin; d <= dmax; ++d) {
    const float b = (float)d, r = (float)(1.0 / (double)b);
    sweep<<<1 << 16, 256>>>(b, r, mag_lo, mag_hi, dev + 4 * d);
  }
  std::vector<unsigned long long> h(4 * (size_t)(dmax + 1));
  cudaError_t e = cudaMemcpy(h.data(), dev, 32 * (size_t)(dmax + 1), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  int nbad = 0;
  unsigned long long t1 = 0, t2 = 0;
  for (int d = dmin; d <= dmax; ++d) {
    t1 += h[4 * d + 2];
    t2 += h[4 * d + 3];
    if (h[4 * d]) {
      ++nbad;
      printf("d=%d mismatches=%llu one_numerator_bits=0x%08llx\n", d, h[4 * d], h[4 * d + 1] - 1);
    }
  }
  printf("divisors %d..%d, |a| bits %08x..%08x: %d with mismatches for |a| >= 2^-100; below 2^-100: one step %llu, two steps %llu mismatches in total\n",
         dmin, dmax, mag_lo, mag_hi, nbad, t1, t2);
  return 0;
}
