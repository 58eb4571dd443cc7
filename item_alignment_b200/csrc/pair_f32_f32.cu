// Instantiations of the pair kernels for inputs float, gradients float.
#include "pair_kernels.cuh"
namespace ia {
int launch_pair_f32_f32(int mode, bool cosloss, int measure, const PairParams& p, bool vec_ok, cudaStream_t s) {
  return launch_pair<float, float>(mode, cosloss, measure, p, vec_ok, s);
}
}  // namespace ia
