// explicit instantiations of the fused loss kernels (split over files to build in parallel)
#include "dvf_loss_kernel.cuh"

namespace dvf {
template <int kV, bool kZeros, bool kTma>
static void launch_loss_c3_t(const LossParams& prm, int blocks, bool expl, bool grad, cudaStream_t st) {
  if (prm.disparity != 0 || prm.img_scale != 1.0f) {   // producer glue: its own variants (zeros padding only, checked by the host)
    if constexpr (kZeros) {
      if (expl) {
        if (grad) launch_balanced<photo_loss_c3x2_kernel<kV, kZeros, true, true, kTma, true>>(prm, blocks, st);
        else launch_balanced<photo_loss_c3x2_kernel<kV, kZeros, true, false, kTma, true>>(prm, blocks, st);
      } else {
        if (grad) launch_balanced<photo_loss_c3x2_kernel<kV, kZeros, false, true, kTma, true>>(prm, blocks, st);
        else launch_balanced<photo_loss_c3x2_kernel<kV, kZeros, false, false, kTma, true>>(prm, blocks, st);
      }
    }
    return;
  }
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
template void launch_loss_c3<1, false>(const LossParams&, int, bool, bool, bool, cudaStream_t);
template void launch_loss_c3<2, false>(const LossParams&, int, bool, bool, bool, cudaStream_t);
template void launch_loss_c3<3, false>(const LossParams&, int, bool, bool, bool, cudaStream_t);
template void launch_loss_c3<4, false>(const LossParams&, int, bool, bool, bool, cudaStream_t);
}  // namespace dvf
