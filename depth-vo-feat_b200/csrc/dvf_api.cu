// dvf_api.cu -- version / error strings of the C ABI (include/dvf_b200.h)
#include "dvf_internal.h"

DVF_EXPORT int dvf_version(void) { return DVF_ABI_VERSION; }

DVF_EXPORT const char* dvf_strerror(int status) {
  switch (status) {
    case DVF_OK: return "ok";
    case DVF_EINVAL_SHAPE: return "invalid shape (size <= 0 or above a DVF_MAX_* limit)";
    case DVF_EINVAL_DTYPE: return "invalid dtype / layout / mode value";
    case DVF_EINVAL_ALIGN: return "pointer not aligned for its element type";
    case DVF_EINVAL_NULL: return "required pointer is NULL";
    case DVF_EUNSUPPORTED: return "combination not implemented by this build";
    case DVF_EWORKSPACE: return "workspace missing or too small";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "unknown dvf status";
}

// ---- self-test of the shared-reciprocal IEEE division used by the coordinate chain ----------
#include "dvf_math.cuh"

namespace dvf {
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}
// numerators |a| in [2^-30, 2^40], divisors b in [2^-10, 2^40] (mode 0) or integers 1..65535 (mode 1)
__global__ void selftest_div_kernel(uint64_t seed, uint64_t n, int mode, unsigned long long* mismatches) {
  unsigned long long bad = 0;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t ra = mix32(seed + 2 * k), rb = mix32(seed + 2 * k + 1);
    const uint32_t ea = 97u + (ra >> 9) % 71u;                     // exponent field for 2^-30 .. 2^40
    const float a = __uint_as_float((ra & 0x80000000u) | (ea << 23) | (ra & 0x007fffffu));
    float b, r;
    if (mode == 0) {
      const uint32_t eb = 117u + (rb >> 9) % 51u;                  // 2^-10 .. 2^40
      b = __uint_as_float((eb << 23) | (rb & 0x007fffffu));
      r = rcp_refined(b);
    } else {
      b = (float)(1u + rb % 65535u);
      r = (float)(1.0 / (double)b);
    }
    if (__float_as_uint(div_by(a, b, r)) != __float_as_uint(__fdiv_rn(a, b))) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}
}  // namespace dvf

// Compares dvf::div_by against __fdiv_rn on n pseudo-random operand pairs; *mismatches (device,
// zeroed by the caller) receives the number of differing results.  Diagnostic entry point.
DVF_EXPORT int dvf_selftest_fast_div(uint64_t seed, uint64_t n, int32_t mode, unsigned long long* mismatches, void* stream) {
  if (!mismatches) return DVF_EINVAL_NULL;
  dvf::selftest_div_kernel<<<dvf::num_sms() * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, n, mode, mismatches);
  return dvf::launch_status();
}
