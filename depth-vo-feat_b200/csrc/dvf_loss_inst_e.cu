// explicit instantiations of the channels-last feature-map loss kernels
#include "dvf_loss_nhwc.cuh"

namespace dvf {
template <int kV, bool kZeros>
void launch_loss_nhwc(const LossParams& prm, int blocks, bool bf16, cudaStream_t st) {
  if (bf16) launch_balanced<photo_loss_nhwc_kernel<kV, kZeros, true>>(prm, blocks, st);
  else launch_balanced<photo_loss_nhwc_kernel<kV, kZeros, false>>(prm, blocks, st);
}
template void launch_loss_nhwc<1, true>(const LossParams&, int, bool, cudaStream_t);
template void launch_loss_nhwc<2, true>(const LossParams&, int, bool, cudaStream_t);
template void launch_loss_nhwc<3, true>(const LossParams&, int, bool, cudaStream_t);
template void launch_loss_nhwc<4, true>(const LossParams&, int, bool, cudaStream_t);
template void launch_loss_nhwc<1, false>(const LossParams&, int, bool, cudaStream_t);
template void launch_loss_nhwc<2, false>(const LossParams&, int, bool, cudaStream_t);
template void launch_loss_nhwc<3, false>(const LossParams&, int, bool, cudaStream_t);
template void launch_loss_nhwc<4, false>(const LossParams&, int, bool, cudaStream_t);
}  // namespace dvf
