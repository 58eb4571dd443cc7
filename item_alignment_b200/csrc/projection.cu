// Head projection in front of the score: x = tanh(f1 . W^T + b), y = tanh(f2 . W^T + b) with the
// shared `dense` of VecSimClassificationHead (reference src/models/base.py:47-49,67-75; dropout inactive), and,
// optionally, the pair score of (x, y) computed in the same epilogue so that inference never writes the embeddings.
//
// Persistent, warp-specialised tcgen05 GEMM, one CTA per SM, 192 threads:
//   warp 0 (one lane)  TMA producer: per K block one 128x64 box of f1, one of f2 and one 128x64 box of W
//                      (128B swizzle) -> a 48 KB stage, 4 stages, full/empty mbarriers;
//   warp 1 (one lane)  tcgen05.mma M=128 x N=128 x K=16, two per K step (f1.W^T and f2.W^T share the W tile),
//                      accumulators [x | y] = 256 TMEM columns, double-buffered in the 512 columns;
//   warps 2-5          epilogue, one row per thread: tcgen05.ld -> + bias -> tanh -> round to the output type ->
//                      128-bit stores, and (SCORE) the row sums of the pair score from the ROUNDED values, so the
//                      result equals scoring the written embeddings.
// A CTA owns (row block, part) = `cts_per_part` consecutive 128-column tiles of one 128-row block; row sums of the parts
// are combined in a fixed order by project_score_finalize (deterministic, no atomics).
#include <cuda.h>

#include <type_traits>

#include "pair_kernels.cuh"
#include "ptx_sm100.cuh"

namespace ia {

int make_tmap(CUtensorMap* map, int dtype, const void* base, int64_t rows, int64_t d, int64_t ld, int box_rows);

namespace proj {
constexpr int BM = 128, BN = 128, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, W_BYTES = BN * BK * 2, STAGE_BYTES = 2 * A_BYTES + W_BYTES;
constexpr int THREADS = 192;
constexpr int kMaxH = 4096;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + kMaxH * 4 + 256 + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
constexpr int kNoScore = -1;
}  // namespace proj

struct ProjParams {
  int64_t n;
  int h, kblocks, n_rb, n_ct, parts, cts_per_part;
  const float* bias;   // [h] or null
  void* x;             // [n, h] outputs in the input type; null = do not write the embeddings
  void* y;
  int64_t ldx, ldy;
  float4* sums;        // SCORE: [parts][n] partial (xy, xx, yy, dist)
};

// tanh to ~1e-6 relative (the output is rounded to 8 or 11 mantissa bits right after): 1 - 2/(e^{2a}+1) away from
// zero, odd polynomial below 0.1 where that form cancels.  2 MUFU + ~8 ALU per value, hidden under the MMAs.
__device__ __forceinline__ float tanh_f32(float v) {
  const float a = fabsf(v);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  const float big = fmaf(-2.0f, r, 1.0f);
  const float v2 = v * v;
  const float small = a * fmaf(v2, fmaf(v2, 0.13333334f, -0.33333334f), 1.0f);
  return copysignf(a < 0.1f ? small : big, v);
}

template <typename T> __device__ __forceinline__ uint4 pack8(const float* f);
template <> __device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float* f) {
  uint4 v;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(f[0], f[1]); v.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[2], f[3]); v.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[4], f[5]); v.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[6], f[7]); v.w = *reinterpret_cast<uint32_t*>(&t);
  return v;
}
template <> __device__ __forceinline__ uint4 pack8<__half>(const float* f) {
  uint4 v;
  __half2 t;
  t = __floats2half2_rn(f[0], f[1]); v.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2half2_rn(f[2], f[3]); v.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2half2_rn(f[4], f[5]); v.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2half2_rn(f[6], f[7]); v.w = *reinterpret_cast<uint32_t*>(&t);
  return v;
}

template <typename T, int MEASURE>
__global__ void __launch_bounds__(proj::THREADS, 1)
project_kernel(const __grid_constant__ CUtensorMap tmap_f1, const __grid_constant__ CUtensorMap tmap_f2,
               const __grid_constant__ CUtensorMap tmap_w, const ProjParams p) {
  using namespace proj;
  constexpr bool SCORE = MEASURE != kNoScore;
  constexpr int AB_FORMAT = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* bias_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + kMaxH * 4);
  uint64_t* full_bar = bars;                // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;      // [STAGES] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;  // [2] MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;     // [2] epilogue -> MMA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_f1);
    tma_prefetch_desc(&tmap_f2);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  // bias (zero beyond h, so padded columns come out as tanh(0) = 0)
  for (int i = threadIdx.x; i < p.n_ct * BN; i += THREADS) bias_s[i] = (p.bias != nullptr && i < p.h) ? __ldg(p.bias + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int n_items = p.n_rb * p.parts;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int rb = item / p.parts, part = item % p.parts;
        const int ct0 = part * p.cts_per_part, ct1 = min(ct0 + p.cts_per_part, p.n_ct);
        for (int ct = ct0; ct < ct1; ++ct) {
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            uint8_t* s0 = smem + stage * STAGE_BYTES;
            tma_load_2d(s0, &tmap_f1, &full_bar[stage], kb * BK, rb * BM);
            tma_load_2d(s0 + A_BYTES, &tmap_f2, &full_bar[stage], kb * BK, rb * BM);
            tma_load_2d(s0 + 2 * A_BYTES, &tmap_w, &full_bar[stage], kb * BK, ct * BN);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(BM, BN, AB_FORMAT);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int part = item % p.parts;
        const int ct0 = part * p.cts_per_part, ct1 = min(ct0 + p.cts_per_part, p.n_ct);
        for (int ct = ct0; ct < ct1; ++ct) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t dx_tmem = tmem_base + acc * (2 * BN);
          const uint32_t dy_tmem = dx_tmem + BN;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t s0 = smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t a1 = umma_smem_desc_sw128(s0);
            const uint64_t a2 = umma_smem_desc_sw128(s0 + A_BYTES);
            const uint64_t wd = umma_smem_desc_sw128(s0 + 2 * A_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < BK / 16; ++k4) {
              umma_f16(dx_tmem, a1 + 2 * k4, wd + 2 * k4, idesc, (kb | k4) != 0);
              umma_f16(dy_tmem, a2 + 2 * k4, wd + 2 * k4, idesc, (kb | k4) != 0);
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull_bar[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: bias + tanh + round (+ row sums)
    const int e = warp & 3;                 // TMEM lane quarter this warp may access
    const int row_local = e * 32 + lane;
    T* xo = reinterpret_cast<T*>(p.x);
    T* yo = reinterpret_cast<T*>(p.y);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int rb = item / p.parts, part = item % p.parts;
      const int ct0 = part * p.cts_per_part, ct1 = min(ct0 + p.cts_per_part, p.n_ct);
      const int64_t row = (int64_t)rb * BM + row_local;
      const bool row_ok = row < p.n;
      RowSums s{0.f, 0.f, 0.f, 0.f};
      for (int ct = ct0; ct < ct1; ++ct) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(e * 32) << 16) + acc * (2 * BN);
#pragma unroll 1
        for (int g = 0; g < BN / 32; ++g) {
          uint32_t rx[32], ry[32];
          tmem_ld_32x32(taddr + g * 32, rx);
          tmem_ld_32x32(taddr + BN + g * 32, ry);
          tmem_ld_wait();
          const int col0 = ct * BN + g * 32;
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            float fx[8], fy[8];
            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + col0 + c);
            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + col0 + c + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              fx[j] = tanh_f32(__uint_as_float(rx[c + j]) + bb[j]);
              fy[j] = tanh_f32(__uint_as_float(ry[c + j]) + bb[j]);
            }
            const uint4 vx = pack8<T>(fx), vy = pack8<T>(fy);
            const bool col_ok = col0 + c < p.h;   // h % 8 == 0: a chunk is entirely inside or outside
            if (row_ok && col_ok) {
              if (xo != nullptr) stg_stream(reinterpret_cast<uint4*>(xo + row * p.ldx + col0 + c), vx);
              if (yo != nullptr) stg_stream(reinterpret_cast<uint4*>(yo + row * p.ldy + col0 + c), vy);
            }
            if (SCORE && col_ok) {
              unpack<T>(vx, fx);   // score the values the embeddings hold after rounding, like the reference does
              unpack<T>(vy, fy);
              accumulate<MEASURE, false>(fx, fy, 8, s);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (SCORE && row_ok) p.sums[(size_t)part * p.n + row] = make_float4(s.xy, s.xx, s.yy, s.dist);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// combine the parts' row sums in index order -> score, probability map (base.py:79-86), threshold label
template <int MEASURE>
__global__ void __launch_bounds__(256) project_score_finalize(const float4* __restrict__ sums, int parts, int64_t n,
                                                              float* __restrict__ sim, float* __restrict__ probs,
                                                              double threshold, uint8_t* __restrict__ labels) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  RowSums s{0.f, 0.f, 0.f, 0.f};
  for (int q = 0; q < parts; ++q) {
    const float4 v = __ldg(sums + (size_t)q * n + row);
    s.xy += v.x; s.xx += v.y; s.yy += v.z; s.dist += v.w;
  }
  float nx, ny;
  const float sc = score_from_sums<MEASURE>(s, nx, ny);
  const float pr = prob_of<MEASURE>(sc);
  sim[row] = sc;
  if (probs != nullptr) probs[row] = pr;
  if (labels != nullptr) labels[row] = (uint8_t)((double)pr >= threshold);
}

// (row block, part) plan: fewest waves x (tiles per item + fixed cost), ties -> fewer parts
static void plan_parts(int n_rb, int n_ct, int kblocks, int ctas, int* parts, int* cts_per_part) {
  double best = 1e300;
  int best_parts = 1;
  for (int q = 1; q <= n_ct; ++q) {
    const int cpp = (n_ct + q - 1) / q;
    const int real_parts = (n_ct + cpp - 1) / cpp;
    if (real_parts != q) continue;
    const long long items = (long long)n_rb * q;
    const long long waves = (items + ctas - 1) / ctas;
    const double cost = (double)waves * ((double)cpp * kblocks * 512.0 + 2000.0);
    if (cost < best * 0.999) { best = cost; best_parts = q; }
  }
  *parts = best_parts;
  *cts_per_part = (n_ct + best_parts - 1) / best_parts;
}

struct ProjPlan {
  int n_rb, n_ct, kblocks, parts, cts_per_part, ctas;
};

static int make_plan(int64_t n, int64_t k_in, int64_t h, ProjPlan* pl) {
  if (n <= 0 || k_in <= 0 || h <= 0) { set_error("projection: n, k_in and h must be positive"); return IA_ERR_INVALID; }
  if ((k_in % 8) != 0 || (h % 8) != 0) { set_error("projection: k_in and h must be multiples of 8 (16-byte rows)"); return IA_ERR_INVALID; }
  if (h > proj::kMaxH) { set_error("projection: h = %lld exceeds %d", (long long)h, proj::kMaxH); return IA_ERR_UNSUPPORTED; }
  if (n > (int64_t)proj::BM * 0x7fffff) { set_error("projection: n too large"); return IA_ERR_UNSUPPORTED; }
  pl->n_rb = (int)((n + proj::BM - 1) / proj::BM);
  pl->n_ct = (int)((h + proj::BN - 1) / proj::BN);
  pl->kblocks = (int)((k_in + proj::BK - 1) / proj::BK);
  pl->ctas = sm_count();
  plan_parts(pl->n_rb, pl->n_ct, pl->kblocks, pl->ctas, &pl->parts, &pl->cts_per_part);
  return IA_OK;
}

template <typename T, int MEASURE>
static int launch_project(const CUtensorMap& m1, const CUtensorMap& m2, const CUtensorMap& mw, const ProjParams& p, int ctas,
                          cudaStream_t st) {
  auto kern = project_kernel<T, MEASURE>;
  static bool configured = false;   // per instantiation
  if (!configured) {
    IA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, proj::SMEM_BYTES));
    configured = true;
  }
  const int items = p.n_rb * p.parts;
  kern<<<items < ctas ? items : ctas, proj::THREADS, proj::SMEM_BYTES, st>>>(m1, m2, mw, p);
  IA_LAUNCH_CHECK();
  return IA_OK;
}

template <typename T>
static int project_dispatch(int measure, const CUtensorMap& m1, const CUtensorMap& m2, const CUtensorMap& mw, const ProjParams& p,
                            int ctas, cudaStream_t st) {
  switch (measure) {
    case proj::kNoScore: return launch_project<T, proj::kNoScore>(m1, m2, mw, p, ctas, st);
    case IA_INNER: return launch_project<T, IA_INNER>(m1, m2, mw, p, ctas, st);
    case IA_COSINE: return launch_project<T, IA_COSINE>(m1, m2, mw, p, ctas, st);
    case IA_L1: return launch_project<T, IA_L1>(m1, m2, mw, p, ctas, st);
    case IA_L2: return launch_project<T, IA_L2>(m1, m2, mw, p, ctas, st);
  }
  set_error("Unsupported similarty measure: %d", measure);
  return IA_ERR_INVALID;
}

static int project_common(int measure, int dtype, const void* f1, const void* f2, int64_t ldf1, int64_t ldf2, int64_t n,
                          int64_t k_in, const void* w, int64_t ldw, const float* bias, int64_t h, void* x, void* y, int64_t ldx,
                          int64_t ldy, float4* sums, const ProjPlan& pl, cudaStream_t st) {
  if (dtype != IA_BF16 && dtype != IA_F16) {
    set_error("projection: only bf16 / fp16 features run on the tensor cores (fp32 would need TF32 and change the numerics)");
    return IA_ERR_UNSUPPORTED;
  }
  if (f1 == nullptr || f2 == nullptr || w == nullptr) { set_error("projection: null input"); return IA_ERR_INVALID; }
  if ((ldf1 % 8) || (ldf2 % 8) || (ldw % 8) || ldf1 < k_in || ldf2 < k_in || ldw < k_in) {
    set_error("projection: leading dimensions must be >= k_in and multiples of 8");
    return IA_ERR_INVALID;
  }
  if ((x != nullptr && ((ldx % 8) || ldx < h)) || (y != nullptr && ((ldy % 8) || ldy < h))) {
    set_error("projection: output leading dimensions must be >= h and multiples of 8");
    return IA_ERR_INVALID;
  }
  if (((uintptr_t)f1 | (uintptr_t)f2 | (uintptr_t)w | (uintptr_t)x | (uintptr_t)y) & 15) {
    set_error("projection: pointers must be 16-byte aligned");
    return IA_ERR_INVALID;
  }
  CUtensorMap m1, m2, mw;
  int rc;
  if ((rc = make_tmap(&m1, dtype, f1, n, k_in, ldf1, proj::BM)) != IA_OK) return rc;
  if ((rc = make_tmap(&m2, dtype, f2, n, k_in, ldf2, proj::BM)) != IA_OK) return rc;
  if ((rc = make_tmap(&mw, dtype, w, h, k_in, ldw, proj::BN)) != IA_OK) return rc;
  ProjParams p;
  p.n = n; p.h = (int)h; p.kblocks = pl.kblocks; p.n_rb = pl.n_rb; p.n_ct = pl.n_ct; p.parts = pl.parts;
  p.cts_per_part = pl.cts_per_part; p.bias = bias; p.x = x; p.y = y; p.ldx = ldx; p.ldy = ldy; p.sums = sums;
  return dtype == IA_BF16 ? project_dispatch<__nv_bfloat16>(measure, m1, m2, mw, p, pl.ctas, st)
                          : project_dispatch<__half>(measure, m1, m2, mw, p, pl.ctas, st);
}

}  // namespace ia

using namespace ia;

extern "C" {

int ia_project_tanh_fwd(int dtype, const void* f1, const void* f2, int64_t ldf1, int64_t ldf2, int64_t n, int64_t k_in,
                        const void* w, int64_t ldw, const float* bias, int64_t h, void* x, void* y, int64_t ldx, int64_t ldy,
                        ia_stream_t stream) {
  ProjPlan pl;
  int rc = make_plan(n, k_in, h, &pl);
  if (rc != IA_OK) return rc;
  if (x == nullptr || y == nullptr) { set_error("projection: null output"); return IA_ERR_INVALID; }
  return project_common(proj::kNoScore, dtype, f1, f2, ldf1, ldf2, n, k_in, w, ldw, bias, h, x, y, ldx, ldy, nullptr, pl,
                        (cudaStream_t)stream);
}

size_t ia_project_score_workspace_bytes(int64_t n, int64_t k_in, int64_t h) {
  ProjPlan pl;
  if (make_plan(n, k_in, h, &pl) != IA_OK) return 0;
  return (size_t)pl.parts * (size_t)n * sizeof(float4);
}

int ia_project_score_fwd(int measure, int dtype, const void* f1, const void* f2, int64_t ldf1, int64_t ldf2, int64_t n,
                         int64_t k_in, const void* w, int64_t ldw, const float* bias, int64_t h, void* x, void* y, int64_t ldx,
                         int64_t ldy, float* sim, float* probs, double threshold, uint8_t* labels_out, void* workspace,
                         size_t workspace_bytes, ia_stream_t stream) {
  ProjPlan pl;
  int rc = make_plan(n, k_in, h, &pl);
  if (rc != IA_OK) return rc;
  if (measure != IA_INNER && measure != IA_COSINE && measure != IA_L1 && measure != IA_L2) {
    set_error("Unsupported similarty measure: %d", measure);
    return IA_ERR_INVALID;
  }
  if (sim == nullptr) { set_error("projection: sim must not be null"); return IA_ERR_INVALID; }
  const size_t need = (size_t)pl.parts * (size_t)n * sizeof(float4);
  if (workspace == nullptr || workspace_bytes < need || ((uintptr_t)workspace & 15)) {
    set_error("projection: workspace of %zu bytes (16-byte aligned) required, got %zu", need, workspace_bytes);
    return IA_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  rc = project_common(measure, dtype, f1, f2, ldf1, ldf2, n, k_in, w, ldw, bias, h, x, y, ldx, ldy,
                      reinterpret_cast<float4*>(workspace), pl, st);
  if (rc != IA_OK) return rc;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  const float4* sums = reinterpret_cast<const float4*>(workspace);
  switch (measure) {
    case IA_INNER: project_score_finalize<IA_INNER><<<blocks, 256, 0, st>>>(sums, pl.parts, n, sim, probs, threshold, labels_out); break;
    case IA_COSINE: project_score_finalize<IA_COSINE><<<blocks, 256, 0, st>>>(sums, pl.parts, n, sim, probs, threshold, labels_out); break;
    case IA_L1: project_score_finalize<IA_L1><<<blocks, 256, 0, st>>>(sums, pl.parts, n, sim, probs, threshold, labels_out); break;
    default: project_score_finalize<IA_L2><<<blocks, 256, 0, st>>>(sums, pl.parts, n, sim, probs, threshold, labels_out); break;
  }
  IA_LAUNCH_CHECK();
  return IA_OK;
}

}  // extern "C"
