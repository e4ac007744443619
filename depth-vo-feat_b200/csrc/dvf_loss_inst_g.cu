// explicit instantiations of the fused loss kernels (split over files to build in parallel): the kExt variants, i.e. the
// backward of the materialised warp run by the image kernel (dvf_inverse_warp_bwd without d img)
#include "dvf_loss_kernel.cuh"

namespace dvf {
void launch_warp_bwd_fused(const LossParams& prm, int blocks, bool zeros, cudaStream_t st) {
  if (zeros) launch_balanced<photo_loss_c3x2_kernel<1, true, false, true, true, false, c3_min_blocks(1), true>>(prm, blocks, st);
  else launch_balanced<photo_loss_c3x2_kernel<1, false, false, true, true, false, c3_min_blocks(1), true>>(prm, blocks, st);
}
}  // namespace dvf
