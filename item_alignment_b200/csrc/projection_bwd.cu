// Training side of the head projection (reference src/models/base.py:50-53,67-75 in train() mode and its
// autograd backward):   x = drop(tanh(dense(drop(f))))
//
//   ia_dropout_fwd            f' = f * mask / (1 - p)                       (input dropout; Philox masks of philox.cuh)
//   ia_project_tanh_dropout_fwd   (projection.cu) GEMM + bias + tanh + output dropout in the epilogue
//   ia_tanh_dropout_bwd       d_pre = g * [out != 0] / (1 - p) * (1 - (out (1 - p))^2)   (generic upstream gradient; the fused
//                             pair-loss kernel writes d_pre itself: ia_pair_score_loss_fwd_bwd(..., act_bwd_scale))
//   ia_project_dgrad          (projection.cu) df = (d_pre . W) * input mask / (1 - p): the forward kernel on (d_pre, W^T)
//   ia_project_wgrad          dW = d_pre1^T f1' + d_pre2^T f2'   -- THIS FILE: a tcgen05 GEMM whose contraction runs over the
//                             ROWS (131 072 at the benchmark shape), so both operands are MN-major in shared memory:
//                             TMA boxes of [64 rows x 64 columns] (128B swizzle) are exactly the canonical MN-major SW128
//                             atoms (8 rows x 128 B, stride 1024 B between 8-row groups = SBO, 8 KB between 64-column
//                             groups = LBO); instruction descriptor with a_major = b_major = MN.  Split-K over the CTAs
//                             (the 1024 x 1024 output has only 32 tiles of 128 x 256), fp32 partial tiles reduced in a fixed
//                             order by wgrad_reduce_kernel (deterministic), db = column sums of d_pre in the same reduce.
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "philox.cuh"
#include "ptx_sm100.cuh"

namespace ia {

int make_tmap(CUtensorMap* map, int dtype, const void* base, int64_t rows, int64_t d, int64_t ld, int box_rows);

// ------------------------------------------------------------------------------------------------ elementwise kernels
template <typename T> __device__ __forceinline__ float h2f(unsigned short b);
template <> __device__ __forceinline__ float h2f<__nv_bfloat16>(unsigned short b) { return __uint_as_float((uint32_t)b << 16); }
template <> __device__ __forceinline__ float h2f<__half>(unsigned short b) { return __half2float(__ushort_as_half(b)); }
template <typename T> __device__ __forceinline__ unsigned short f2h(float v);
template <> __device__ __forceinline__ unsigned short f2h<__nv_bfloat16>(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
template <> __device__ __forceinline__ unsigned short f2h<__half>(float v) { return __half_as_ushort(__float2half_rn(v)); }

// out = x * mask / (1 - p); a thread owns an 8 x 8 block (8 rows of one 16-byte vector): one Philox call per column
template <typename T>
__global__ void __launch_bounds__(256) dropout_fwd_kernel(const T* __restrict__ x, int64_t ldx, int64_t rows, int cols, DropoutParams d,
                                                          uint32_t stream_id, T* __restrict__ out, int64_t ldo) {
  const int cvec = cols >> 3;
  const int64_t blocks = ((rows + 7) >> 3) * cvec;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < blocks; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t rg = b / cvec;
    const int c0 = (int)(b % cvec) * 8;
    uint32_t keep[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) keep[j] = dropout_keep8(d, stream_id, (uint64_t)rg, (uint32_t)cols, (uint32_t)(c0 + j));
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int64_t row = rg * 8 + r;
      if (row >= rows) break;
      const uint4 v = *reinterpret_cast<const uint4*>(x + row * ldx + c0);
      const unsigned short* e = reinterpret_cast<const unsigned short*>(&v);
      uint4 o;
      unsigned short* oe = reinterpret_cast<unsigned short*>(&o);
#pragma unroll
      for (int j = 0; j < 8; ++j) oe[j] = ((keep[j] >> r) & 1u) ? f2h<T>(h2f<T>(e[j]) * d.scale) : (unsigned short)0;
      *reinterpret_cast<uint4*>(out + row * ldo + c0) = o;
    }
  }
}

// d_pre = g * keep / (1 - p) * (1 - t^2), t = tanh output = out * (1 - p) where kept.  With dropout active the mask is read
// off the output itself (a dropped element is exactly 0; a kept tanh value is 0 only if the pre-activation was exactly 0).
template <typename T>
__global__ void __launch_bounds__(256) tanh_dropout_bwd_kernel(const T* __restrict__ g, int64_t ldg, const T* __restrict__ out, int64_t ldo,
                                                               int64_t rows, int cols, float keep_scale, T* __restrict__ dpre, int64_t ldd) {
  const int cvec = cols >> 3;
  const int64_t total = rows * cvec;
  const float inv = keep_scale > 0.f ? 1.0f / keep_scale : 1.0f, sc = keep_scale > 0.f ? keep_scale : 1.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / cvec;
    const int c0 = (int)(i % cvec) * 8;
    const uint4 gv = *reinterpret_cast<const uint4*>(g + row * ldg + c0);
    const uint4 ov = *reinterpret_cast<const uint4*>(out + row * ldo + c0);
    const unsigned short* ge = reinterpret_cast<const unsigned short*>(&gv);
    const unsigned short* oe = reinterpret_cast<const unsigned short*>(&ov);
    uint4 r;
    unsigned short* re = reinterpret_cast<unsigned short*>(&r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float o = h2f<T>(oe[j]), t = o * inv;
      const bool keep = keep_scale > 0.f ? (o != 0.f) : true;
      re[j] = keep ? f2h<T>(__fmul_rn(__fmul_rn(h2f<T>(ge[j]), sc), __fmaf_rn(-t, t, 1.0f))) : (unsigned short)0;
    }
    *reinterpret_cast<uint4*>(dpre + row * ldd + c0) = r;
  }
}

// dst[c][r] = src[r][c] (16-bit elements), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose16_kernel(const unsigned short* __restrict__ src, int rows, int cols, int64_t lds,
                                                          unsigned short* __restrict__ dst, int64_t ldd) {
  __shared__ unsigned short tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8)
    if (r0 + j < rows && c0 + tx < cols) tile[j][tx] = src[(int64_t)(r0 + j) * lds + c0 + tx];
  __syncthreads();
  for (int j = ty; j < 32; j += 8)
    if (c0 + j < cols && r0 + tx < rows) dst[(int64_t)(c0 + j) * ldd + r0 + tx] = tile[tx][j];
}

// ------------------------------------------------------------------------------------------------ wgrad (tcgen05, MN-major)
namespace wg {
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;      // M: output columns h of the head, N: input columns k_in, K: rows
constexpr int BOX_BYTES = BK * 64 * 2;                      // one TMA box: 64 rows x 64 columns of 16-bit = 8 KB
constexpr int A_BYTES = (BM / 64) * BOX_BYTES, B_BYTES = (BN / 64) * BOX_BYTES, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 192;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
}  // namespace wg

// MN-major operand tile staged by TMA with 128-byte swizzle: 8-row groups (K) 1024 B apart (SBO), 64-column groups (MN) 8 KB apart (LBO)
__device__ __forceinline__ uint64_t umma_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct WgradParams {
  int h, k_in, kb_per_side, kb_total, splits, kb_per_split, n_mt, n_nt;
  float* partial;       // [splits][h][k_in] fp32
};

template <typename T>
__global__ void __launch_bounds__(wg::THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmap_d1, const __grid_constant__ CUtensorMap tmap_d2,
             const __grid_constant__ CUtensorMap tmap_f1, const __grid_constant__ CUtensorMap tmap_f2, const WgradParams p) {
  using namespace wg;
  constexpr int AB_FORMAT = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_d1); tma_prefetch_desc(&tmap_d2); tma_prefetch_desc(&tmap_f1); tma_prefetch_desc(&tmap_f2);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // item = (m tile, n tile, split)
  const int item = blockIdx.x;
  const int split = item % p.splits;
  const int nt = (item / p.splits) % p.n_nt, mt = item / (p.splits * p.n_nt);
  const int kb0 = split * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const bool second = kb >= p.kb_per_side;
        const int row = (second ? kb - p.kb_per_side : kb) * BK;
        const CUtensorMap* md = second ? &tmap_d2 : &tmap_d1;
        const CUtensorMap* mf = second ? &tmap_f2 : &tmap_f1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
        uint8_t* s0 = smem + stage * STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < BM / 64; ++j) tma_load_2d(s0 + j * BOX_BYTES, md, &full_bar[stage], mt * BM + j * 64, row);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_2d(s0 + A_BYTES + j * BOX_BYTES, mf, &full_bar[stage], nt * BN + j * 64, row);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(BM, BN, AB_FORMAT) | (1u << 15) | (1u << 16);    // A and B MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t s0 = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t ad = umma_smem_desc_mn_sw128(s0, BOX_BYTES);
        const uint64_t bd = umma_smem_desc_mn_sw128(s0 + A_BYTES, BOX_BYTES);
#pragma unroll
        for (int k4 = 0; k4 < BK / 16; ++k4)     // 16 rows = two 8-row groups = 2048 B (>> 4 = 128) per K step
          umma_f16(tmem_base, ad + 128 * k4, bd + 128 * k4, idesc, (kb > kb0 || k4 > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  } else {
    // epilogue: TMEM lane = output row h, columns = k_in; fp32 partial tile -> workspace
    const int e = warp & 3;
    const int hrow = mt * BM + e * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    float* dst = p.partial + ((size_t)split * p.h + (size_t)(hrow < p.h ? hrow : 0)) * p.k_in + (size_t)nt * BN;
    const uint32_t taddr = tmem_base + ((uint32_t)(e * 32) << 16);
    if (kb1 > kb0) {
#pragma unroll 1
      for (int g = 0; g < BN / 32; ++g) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + g * 32, r);
        tmem_ld_wait();
        if (hrow < p.h) {
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const int col = nt * BN + g * 32 + c;
            if (col + 3 < p.k_in) {
              *reinterpret_cast<float4*>(dst + g * 32 + c) = make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]),
                                                                         __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
            } else {
              for (int q = 0; q < 4; ++q)
                if (col + q < p.k_in) dst[g * 32 + c + q] = __uint_as_float(r[c + q]);
            }
          }
        }
      }
    } else if (hrow < p.h) {      // a split without K blocks (cannot happen with the host's plan; keeps the reduce well defined)
      for (int c = 0; c < BN; ++c)
        if (nt * BN + c < p.k_in) dst[c] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// dW[i] = sum over splits (index order) of partial[s][i]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int64_t count, float* __restrict__ dw) {
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < count; i += (int64_t)gridDim.x * blockDim.x * 4) {
    float4 a = *reinterpret_cast<const float4*>(partial + i);
    for (int s = 1; s < splits; ++s) {
      const float4 b = *reinterpret_cast<const float4*>(partial + (size_t)s * count + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    *reinterpret_cast<float4*>(dw + i) = a;
  }
}

// db[c] = sum over the rows of d1 and d2 of column c.  Stage 1: a CTA takes 256 rows; warp w accumulates rows w, w+8, ... in
// registers (a lane owns 8 consecutive columns per 256-column group: 16-byte loads, a warp instruction reads 512 contiguous
// bytes, all loads independent), the eight warps are added in index order through shared memory.  Stage 2: one warp per column
// sums the CTA partials (lane-strided, then a shuffle tree).  Fixed orders everywhere: deterministic.
constexpr int kColsumRows = 256, kColsumVecs = 4;    // 4 x 256 columns per pass
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ d1, const T* __restrict__ d2, int64_t ldd, int64_t rows, int cols,
                                                             float* __restrict__ partial) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * kColsumRows;
  for (int cbase = 0; cbase < cols; cbase += 256 * kColsumVecs) {
    float acc[kColsumVecs][8];
#pragma unroll
    for (int j = 0; j < kColsumVecs; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
#pragma unroll 2
    for (int rr = warp; rr < kColsumRows; rr += 8) {
      const int64_t r = r0 + rr;
      if (r >= 2 * rows) break;
      const T* src = r < rows ? d1 + r * ldd : d2 + (r - rows) * ldd;
#pragma unroll
      for (int j = 0; j < kColsumVecs; ++j) {
        const int c = cbase + 256 * j + 8 * lane;
        if (c < cols) {
          const uint4 v = ldg_stream(reinterpret_cast<const uint4*>(src + c));
          const unsigned short* e16 = reinterpret_cast<const unsigned short*>(&v);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[j][e] += h2f<T>(e16[e]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kColsumVecs; ++j) {
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 8; ++e) red[warp][8 * lane + e] = acc[j][e];
      __syncthreads();
      float t = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < 8; ++w2) t += red[w2][threadIdx.x];
      const int c = cbase + 256 * j + threadIdx.x;
      if (c < cols) partial[(size_t)blockIdx.x * cols + c] = t;
    }
  }
}
__global__ void __launch_bounds__(256) colsum_finish_kernel(const float* __restrict__ partial, int n_part, int cols, float* __restrict__ db) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= cols) return;
  float a = 0.f;
  for (int i = lane; i < n_part; i += 32) a += partial[(size_t)i * cols + c];
  a = warp_sum(a);
  if (lane == 0) db[c] = a;
}

struct WgradPlan {
  int n_mt, n_nt, kb_per_side, kb_total, splits, kb_per_split, colsum_ctas;
  size_t partial_bytes, colsum_bytes, bytes;
};
static WgradPlan wgrad_plan(int64_t n, int64_t h, int64_t k_in) {
  WgradPlan pl;
  pl.n_mt = (int)((h + wg::BM - 1) / wg::BM);
  pl.n_nt = (int)((k_in + wg::BN - 1) / wg::BN);
  pl.kb_per_side = (int)((n + wg::BK - 1) / wg::BK);
  pl.kb_total = 2 * pl.kb_per_side;
  const int tiles = pl.n_mt * pl.n_nt;
  int splits = sm_count() / tiles;
  if (splits < 1) splits = 1;
  if (splits > pl.kb_total / 8) splits = pl.kb_total / 8 > 0 ? pl.kb_total / 8 : 1;     // at least 8 K blocks per split
  pl.kb_per_split = (pl.kb_total + splits - 1) / splits;
  pl.splits = (pl.kb_total + pl.kb_per_split - 1) / pl.kb_per_split;
  pl.partial_bytes = ((size_t)pl.splits * h * k_in * sizeof(float) + 255) & ~(size_t)255;
  pl.colsum_ctas = (int)((2 * n + kColsumRows - 1) / kColsumRows);
  pl.colsum_bytes = ((size_t)pl.colsum_ctas * h * sizeof(float) + 255) & ~(size_t)255;
  pl.bytes = pl.partial_bytes + pl.colsum_bytes;
  return pl;
}

static int check16(int dtype) {
  if (dtype != IA_BF16 && dtype != IA_F16) { set_error("projection training kernels run bf16 / fp16 tensors"); return IA_ERR_UNSUPPORTED; }
  return IA_OK;
}

}  // namespace ia

using namespace ia;

extern "C" {

int ia_dropout_fwd(int dtype, const void* x, int64_t ldx, int64_t rows, int64_t cols, float p_drop, uint64_t seed, uint32_t step,
                   uint32_t stream_id, void* out, int64_t ldo, ia_stream_t stream) {
  int rc = check16(dtype);
  if (rc != IA_OK) return rc;
  if (rows < 0 || cols <= 0 || cols % 8 != 0 || ldx < cols || ldo < cols || ldx % 8 != 0 || ldo % 8 != 0 || (rows > 0 && (x == nullptr || out == nullptr)) ||
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) || cols > (1 << 30)) {
    set_error("dropout: rows of 16-byte aligned 8-element vectors expected");
    return IA_ERR_INVALID;
  }
  if (!(p_drop >= 0.f) || !(p_drop < 1.f)) { set_error("dropout probability must be in [0, 1)"); return IA_ERR_INVALID; }
  if (rows == 0) return IA_OK;
  DropoutParams d{(uint32_t)seed, (uint32_t)(seed >> 32), step, dropout_threshold16(p_drop), 1.0f / (1.0f - p_drop)};
  const int64_t blocks = ((rows + 7) / 8) * (cols / 8);
  const int64_t want = (blocks + 255) / 256;
  const int grid = (int)(want < 16 * sm_count() ? want : 16 * sm_count());
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == IA_BF16) dropout_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, ldx, rows, (int)cols, d, stream_id, (__nv_bfloat16*)out, ldo);
  else dropout_fwd_kernel<__half><<<grid, 256, 0, s>>>((const __half*)x, ldx, rows, (int)cols, d, stream_id, (__half*)out, ldo);
  IA_LAUNCH_CHECK();
  return IA_OK;
}

int ia_tanh_dropout_bwd(int dtype, const void* g, int64_t ldg, const void* out, int64_t ldo, int64_t rows, int64_t cols, float keep_scale,
                        void* dpre, int64_t ldd, ia_stream_t stream) {
  int rc = check16(dtype);
  if (rc != IA_OK) return rc;
  if (rows < 0 || cols <= 0 || cols % 8 != 0 || ldg < cols || ldo < cols || ldd < cols || (ldg | ldo | ldd) % 8 != 0 ||
      (rows > 0 && (g == nullptr || out == nullptr || dpre == nullptr)) ||
      ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dpre)) & 15)) {
    set_error("tanh backward: rows of 16-byte aligned 8-element vectors expected");
    return IA_ERR_INVALID;
  }
  if (rows == 0) return IA_OK;
  const int64_t want = (rows * (cols / 8) + 255) / 256;
  const int grid = (int)(want < 16 * sm_count() ? want : 16 * sm_count());
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == IA_BF16)
    tanh_dropout_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)g, ldg, (const __nv_bfloat16*)out, ldo, rows, (int)cols, keep_scale, (__nv_bfloat16*)dpre, ldd);
  else
    tanh_dropout_bwd_kernel<__half><<<grid, 256, 0, s>>>((const __half*)g, ldg, (const __half*)out, ldo, rows, (int)cols, keep_scale, (__half*)dpre, ldd);
  IA_LAUNCH_CHECK();
  return IA_OK;
}

int ia_transpose16(const void* src, int64_t rows, int64_t cols, int64_t lds, void* dst, int64_t ldd, ia_stream_t stream) {
  if (rows <= 0 || cols <= 0 || lds < cols || ldd < rows || src == nullptr || dst == nullptr || rows > (1 << 30) || cols > (1 << 30)) { set_error("transpose: bad arguments"); return IA_ERR_INVALID; }
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  transpose16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned short*)src, (int)rows, (int)cols, lds, (unsigned short*)dst, ldd);
  IA_LAUNCH_CHECK();
  return IA_OK;
}

size_t ia_project_wgrad_workspace_bytes(int64_t n, int64_t h, int64_t k_in) {
  if (n <= 0 || h <= 0 || k_in <= 0) return 0;
  return wgrad_plan(n, h, k_in).bytes;
}

int ia_project_wgrad(int dtype, const void* d1, const void* d2, int64_t ldd, const void* f1, const void* f2, int64_t ldf, int64_t n,
                     int64_t h, int64_t k_in, float* dw, float* db, void* workspace, size_t workspace_bytes, ia_stream_t stream) {
  int rc = check16(dtype);
  if (rc != IA_OK) return rc;
  if (n <= 0 || h <= 0 || k_in <= 0 || h % 8 != 0 || k_in % 8 != 0 || ldd < h || ldf < k_in || ldd % 8 != 0 || ldf % 8 != 0 || dw == nullptr ||
      d1 == nullptr || d2 == nullptr || f1 == nullptr || f2 == nullptr || k_in % 4 != 0 ||
      ((reinterpret_cast<uintptr_t>(d1) | reinterpret_cast<uintptr_t>(d2) | reinterpret_cast<uintptr_t>(f1) | reinterpret_cast<uintptr_t>(f2) |
        reinterpret_cast<uintptr_t>(dw)) & 15)) {
    set_error("wgrad: h, k_in and leading dimensions must be multiples of 8, pointers 16-byte aligned");
    return IA_ERR_INVALID;
  }
  if (n > (int64_t)wg::BK * 0x3fffffff) { set_error("wgrad: n too large"); return IA_ERR_UNSUPPORTED; }
  const WgradPlan pl = wgrad_plan(n, h, k_in);
  if (workspace == nullptr || workspace_bytes < pl.bytes || (reinterpret_cast<uintptr_t>(workspace) & 255)) {
    set_error("wgrad: workspace of %zu bytes (256-byte aligned) required", pl.bytes);
    return IA_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  CUtensorMap md1, md2, mf1, mf2;
  if ((rc = make_tmap(&md1, dtype, d1, n, h, ldd, wg::BK)) != IA_OK) return rc;
  if ((rc = make_tmap(&md2, dtype, d2, n, h, ldd, wg::BK)) != IA_OK) return rc;
  if ((rc = make_tmap(&mf1, dtype, f1, n, k_in, ldf, wg::BK)) != IA_OK) return rc;
  if ((rc = make_tmap(&mf2, dtype, f2, n, k_in, ldf, wg::BK)) != IA_OK) return rc;
  WgradParams p;
  p.h = (int)h; p.k_in = (int)k_in; p.kb_per_side = pl.kb_per_side; p.kb_total = pl.kb_total; p.splits = pl.splits;
  p.kb_per_split = pl.kb_per_split; p.n_mt = pl.n_mt; p.n_nt = pl.n_nt;
  p.partial = reinterpret_cast<float*>(workspace);
  const int grid = pl.n_mt * pl.n_nt * pl.splits;
  static bool configured[kMaxDevices][2] = {};
  const int slot = device_slot();
  if (dtype == IA_BF16) {
    if (!configured[slot][0]) { IA_CUDA_CHECK(cudaFuncSetAttribute(wgrad_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BYTES)); configured[slot][0] = true; }
    wgrad_kernel<__nv_bfloat16><<<grid, wg::THREADS, wg::SMEM_BYTES, s>>>(md1, md2, mf1, mf2, p);
  } else {
    if (!configured[slot][1]) { IA_CUDA_CHECK(cudaFuncSetAttribute(wgrad_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BYTES)); configured[slot][1] = true; }
    wgrad_kernel<__half><<<grid, wg::THREADS, wg::SMEM_BYTES, s>>>(md1, md2, mf1, mf2, p);
  }
  IA_LAUNCH_CHECK();
  const int64_t count = h * k_in;
  const int64_t want = (count / 4 + 255) / 256;
  wgrad_reduce_kernel<<<(int)(want < 8 * sm_count() ? want : 8 * sm_count()), 256, 0, s>>>(p.partial, pl.splits, count, dw);
  IA_LAUNCH_CHECK();
  if (db != nullptr) {
    float* cpart = reinterpret_cast<float*>(static_cast<char*>(workspace) + pl.partial_bytes);
    if (dtype == IA_BF16) colsum_partial_kernel<__nv_bfloat16><<<pl.colsum_ctas, 256, 0, s>>>((const __nv_bfloat16*)d1, (const __nv_bfloat16*)d2, ldd, n, (int)h, cpart);
    else colsum_partial_kernel<__half><<<pl.colsum_ctas, 256, 0, s>>>((const __half*)d1, (const __half*)d2, ldd, n, (int)h, cpart);
    IA_LAUNCH_CHECK();
    colsum_finish_kernel<<<(int)((h + 7) / 8), 256, 0, s>>>(cpart, pl.colsum_ctas, (int)h, db);
    IA_LAUNCH_CHECK();
  }
  return IA_OK;
}

}  // extern "C"
