// Issue-rate microbenchmark: tcgen05.mma 128 x N x 16 (one CTA) vs 256 x N x 16 (CTA pair, cta_group::2) on whatever bytes sit in
// shared memory (no TMA, no epilogue): cycles per MMA on one SM / one pair.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -I item_alignment_b200/csrc -o /tmp/mma_rate scripts/micro/mma_rate.cu
#include <cstdio>
#include "ptx_sm100.cuh"
using namespace ia;

template <int CTAS, int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CTAS == 1 ? 0 : (int)cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 4 * 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (warp == 0) { if (CTAS == 1) tmem_alloc(&tmem_ptr, 512); else tmem_alloc_pair(&tmem_ptr, 512); }
  tc_fence_before();
  if (CTAS == 1) __syncthreads(); else cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (warp == 1 && lane == 0 && rank == 0) {
    constexpr uint32_t idesc = umma_idesc(128 * CTAS, N, 1);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int stage = it & 3;
      const uint32_t a_addr = smem_u32(smem + stage * 49152);
      const uint64_t a_desc = umma_smem_desc_sw128(a_addr), b_desc = umma_smem_desc_sw128(a_addr + 16384);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        if (CTAS == 1) umma_f16(tmem_base + (it & 1) * 256, a_desc + 2 * k4, b_desc + 2 * k4, idesc, 1);
        else umma_f16_pair(tmem_base + (it & 1) * 256, a_desc + 2 * k4, b_desc + 2 * k4, idesc, 1);
      }
    }
    if (CTAS == 1) umma_commit(&bar); else umma_commit_pair(&bar, 1);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x / CTAS] = t1 - t0;
  }
  tc_fence_before();
  if (CTAS == 1) __syncthreads(); else cluster_sync();
  if (warp == 0) { tc_fence_after(); if (CTAS == 1) tmem_dealloc(tmem_base, 512); else tmem_dealloc_pair(tmem_base, 512); }
}

template <int CTAS, int N>
static void run(const char* name, int grid, int iters) {
  long long* out;
  cudaMalloc(&out, sizeof(long long) * 256);
  auto kernel = rate_kernel<CTAS, N>;
  const int smem = 4 * 49152 + 1024;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, kernel, iters, out);
    cudaEventRecord(e1);
    if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const double cyc = (double)h[0] / (4.0 * iters);
    const double flop = 2.0 * 128 * CTAS * N * 16 * 4.0 * iters * (grid / CTAS);
    if (rep) printf("%-28s grid %3d: %.1f cycles per MMA (%.0f FLOP/clk/SM), %.3f ms, %.0f TFLOP/s, %.2f GHz\n", name, grid, cyc,
                    2.0 * 128 * N * 16 / cyc, ms, flop / ms / 1e9, h[0] / ms / 1e6);
  }
  cudaFree(out);
}

int main() {
  const int iters = 20000;
  for (int grid : {2, 148}) {
    run<1, 256>("1 CTA  128x256x16", grid, iters);
    run<2, 256>("pair   256x256x16", grid, iters);
    run<1, 128>("1 CTA  128x128x16", grid, iters);
    run<2, 128>("pair   256x128x16", grid, iters);
    run<2, 64>("pair   256x64x16", grid, iters);
  }
  return 0;
}
