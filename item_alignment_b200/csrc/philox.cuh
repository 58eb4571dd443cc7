// Counter-based dropout masks (Philox4x32-10, the generator family torch's CUDA dropout uses) with a layout chosen for the GEMM
// epilogues of this library: the mask of element (row, col) of a [rows, cols] tensor is lane (row & 7) of the 8 x 16-bit output of
//     philox4x32_10(key = seed, counter = (g_lo, g_hi, stream, step)),   g = (row >> 3) * cols + col
// so a thread that owns ONE column and many rows (a TMEM lane in the projection epilogues) pays one call per 8 rows, and an
// elementwise thread that owns an 8 x 8 block pays one call per column.  Element kept <=> lane value >= round(p * 65536)
// (drop probability quantised to 2^-16); kept values are scaled by 1 / (1 - p) like nn.Dropout (reference base.py:50-53,67-75).
// `stream` separates the tensors of one step (0: f1, 1: f2, 2: x, 3: y), `step` is the caller's call counter.
// torch's own RNG stream cannot be reproduced inside a GEMM epilogue; the masks are checked by replaying this function on the
// host (oracle/formula.philox_keep_mask) and statistically.
#pragma once
#include <stdint.h>

namespace ia {

struct DropoutParams {
  uint32_t seed_lo, seed_hi;
  uint32_t step;        // call counter (Philox counter word 3)
  uint32_t thr16;       // drop if lane < thr16; 0 = dropout off
  float scale;          // 1 / (1 - p)
};

__host__ __device__ inline uint32_t dropout_threshold16(float p) {
  const float t = p * 65536.0f + 0.5f;
  return t <= 0.f ? 0u : (t >= 65535.f ? 65535u : (uint32_t)t);
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// 8 keep bits (bit j = row 8 * row_group + j at column col): one Philox call
__device__ __forceinline__ uint32_t dropout_keep8(const DropoutParams& d, uint32_t stream, uint64_t row_group, uint32_t cols, uint32_t col) {
  const uint64_t g = row_group * (uint64_t)cols + col;
  const uint4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), stream, d.step, d.seed_lo, d.seed_hi);
  uint32_t m = 0;
  m |= (uint32_t)((r.x & 0xffffu) >= d.thr16) << 0; m |= (uint32_t)((r.x >> 16) >= d.thr16) << 1;
  m |= (uint32_t)((r.y & 0xffffu) >= d.thr16) << 2; m |= (uint32_t)((r.y >> 16) >= d.thr16) << 3;
  m |= (uint32_t)((r.z & 0xffffu) >= d.thr16) << 4; m |= (uint32_t)((r.z >> 16) >= d.thr16) << 5;
  m |= (uint32_t)((r.w & 0xffffu) >= d.thr16) << 6; m |= (uint32_t)((r.w >> 16) >= d.thr16) << 7;
  return m;
}

}  // namespace ia
