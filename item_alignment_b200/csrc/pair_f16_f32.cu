// Instantiations of the pair kernels for inputs __half, gradients float.
#include "pair_kernels.cuh"
namespace ia {
int launch_pair_f16_f32(int mode, bool cosloss, int measure, const PairParams& p, bool vec_ok, cudaStream_t s) {
  return launch_pair<__half, float>(mode, cosloss, measure, p, vec_ok, s);
}
}  // namespace ia
