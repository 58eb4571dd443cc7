// Instantiations of the pair kernels for inputs __nv_bfloat16, gradients float.
#include "pair_kernels.cuh"
namespace ia {
int launch_pair_bf16_f32(int mode, bool cosloss, int measure, const PairParams& p, bool vec_ok, cudaStream_t s) {
  return launch_pair<__nv_bfloat16, float>(mode, cosloss, measure, p, vec_ok, s);
}
}  // namespace ia
