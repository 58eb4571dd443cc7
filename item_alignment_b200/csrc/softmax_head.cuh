// Shared declarations of the softmax-head kernels (softmax_head.cu: CUDA-core kernels; softmax_head_mma.cu: the
// tensor-core training kernel for 16-bit rows).
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace ia {

struct HeadParams {
  const void* x;
  const void* y;
  int64_t ldx, ldy;
  const float* w;   // [2, 2h]
  const float* b;   // [2]
  const int64_t* labels;
  int64_t n;
  int h;
  float* logits;    // [n,2] or null
  float* probs;     // [n,2] or null
  float* loss_out;  // scalar
  void* dx;
  void* dy;
  int64_t lddx, lddy;
  float grad_scale;   // upstream / n
  double loss_scale;  // 1/n
  void* workspace;    // [kWorkspaceBytes | float partial[grid][2h+2]]
  int stages;         // ring depth (TRAIN)
  int group;          // adjacent pairs per ring stage (TRAIN): 2 for rows <= 3 KB, else 1
  int load_mode;      // bit 1: ld.global.cs (never set in production; pins the load order of the forward kernel)
  const float* upstream;   // optional DEVICE scalar d(total)/d(loss) folded into every gradient (autograd backward)
  int upstream_skip_one;   // with upstream: leave at once when *upstream == 1 (the gradients already written are exact)
};

constexpr int kHeadMaxGrid = 592;  // 148 SMs x 4

// dW[1][j] = sum over blocks of partial[b][j] in a fixed order; dW[0] = -dW[1]; same for db (softmax_head.cu)
__global__ void __launch_bounds__(256) softmax_head_finalize(const float* partials, int nblocks, int h2, float* dw, float* db, const float* upstream,
                                      int skip_one);

// Tensor-core training kernel (softmax_head_mma.cu).  Returns IA_ERR_UNSUPPORTED when the shape is not eligible (the caller
// then uses the CUDA-core kernel).  dtype / grad_dtype as in ia_softmax_head_fwd_bwd.
bool softmax_head_mma_eligible(int dtype, const HeadParams& p);
int launch_softmax_head_mma(int dtype, int grad_dtype, const HeadParams& p, cudaStream_t stream, float* dw, float* db);

}  // namespace ia
