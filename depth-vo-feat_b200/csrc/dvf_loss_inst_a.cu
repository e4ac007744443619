// explicit instantiations of the fused loss kernels (split over files to build in parallel)
#include "dvf_loss_kernel.cuh"

namespace dvf {
template <int kV, bool kZeros, bool kTma>
static void launch_loss_c3_t(const LossParams& prm, int blocks, bool expl, bool grad, cudaStream_t st) {
  if (expl) {
    if (grad) launch_balanced<photo_loss_c3x2_kernel<kV, kZeros, true, true, kTma>>(prm, blocks, st);
    else launch_balanced<photo_loss_c3x2_kernel<kV, kZeros, true, false, kTma>>(prm, blocks, st);
  } else {
    if (grad) launch_balanced<photo_loss_c3x2_kernel<kV, kZeros, false, true, kTma>>(prm, blocks, st);
    else launch_balanced<photo_loss_c3x2_kernel<kV, kZeros, false, false, kTma>>(prm, blocks, st);
  }
}
template <int kV, bool kZeros>
void launch_loss_c3(const LossParams& prm, int blocks, bool expl, bool grad, bool tma, cudaStream_t st) {
  if (tma) launch_loss_c3_t<kV, kZeros, true>(prm, blocks, expl, grad, st);
  else launch_loss_c3_t<kV, kZeros, false>(prm, blocks, expl, grad, st);
}
template void launch_loss_c3<1, true>(const LossParams&, int, bool, bool, bool, cudaStream_t);
template void launch_loss_c3<2, true>(const LossParams&, int, bool, bool, bool, cudaStream_t);
template void launch_loss_c3<3, true>(const LossParams&, int, bool, bool, bool, cudaStream_t);
template void launch_loss_c3<4, true>(const LossParams&, int, bool, bool, bool, cudaStream_t);
}  // namespace dvf

#ifdef DVF_TRACE
// experiment builds only (not declared in include/dvf_b200.h): per-CTA timestamps of the last zeros-padding image-kernel
// launch (this translation unit's copy of g_trace)
DVF_EXPORT int dvf_debug_trace_read(unsigned long long* host_out, int n_ctas) {
  return (int)cudaMemcpyFromSymbol(host_out, dvf::g_trace, sizeof(unsigned long long) * 8 * n_ctas);
}
#endif
