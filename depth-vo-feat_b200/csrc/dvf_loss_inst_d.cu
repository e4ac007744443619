// explicit instantiations of the fused loss kernels (split over files to build in parallel)
#include "dvf_loss_kernel.cuh"

namespace dvf {
template <int kV, bool kZeros>
void launch_loss_cn(const LossParams& prm, int blocks, cudaStream_t st) {
  photo_loss_cn_kernel<kV, kZeros><<<blocks, kLossThreads, 0, st>>>(prm);
}
template void launch_loss_cn<1, false>(const LossParams&, int, cudaStream_t);
template void launch_loss_cn<2, false>(const LossParams&, int, cudaStream_t);
template void launch_loss_cn<3, false>(const LossParams&, int, cudaStream_t);
template void launch_loss_cn<4, false>(const LossParams&, int, cudaStream_t);
}  // namespace dvf
