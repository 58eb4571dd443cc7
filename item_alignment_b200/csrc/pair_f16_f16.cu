// Instantiations of the pair kernels for inputs __half, gradients __half.
#include "pair_kernels.cuh"
namespace ia {
int launch_pair_f16_f16(int mode, bool cosloss, int measure, const PairParams& p, bool vec_ok, cudaStream_t s) {
  return launch_pair<__half, __half>(mode, cosloss, measure, p, vec_ok, s);
}
}  // namespace ia
