// "softmax" measure, training step (TwoTowerClassificationHead + CrossEntropyLoss forward + backward; reference
// src/models/base.py:103-117, src/models/text.py:1408-1409,1473) for 16-bit rows of h = 512 / 768 / 1024: a warp-specialised
// kernel whose three contractions run on the tensor cores (mma.sync m16n8k16), so that the SM's issue slots and its latency
// hiding are left for streaming rows.
//
// Why: the CUDA-core kernel (softmax_head.cu) spends ~900 instructions per pair at h = 1024 (4 FMA per element pair for the two
// logits, 4 more for dx / dW, the bf16 unpack twice) and measured 77 % of the HBM copy peak with its issue slots 52 % busy and
// its shared-memory pipe 42 % busy re-reading W per row group (profiles/r01/ncu_softmax_head_bf16_12warp.md).
//
// One persistent CTA per SM, 12 warps.  A stage = 16 pairs = the [16 x 2h] tile [x rows | y rows] in shared memory (row pitch
// +16 B: conflict-free ldmatrix), filled with 16-byte cp.async whose completion lands on the stage's mbarrier; three stages.
//   * FORWARD warps 0-3 (each owns h/2 columns of x or y): per 16-column tile one ldmatrix.x4 + ONE mma for the logits -- the B
//     operand is the warp's slice of W held in registers, the two class rows split into three 16-bit terms (hi + lo + lo2 =
//     the fp32 weight to 2^-24) in 6 of the 8 MMA columns.  The four partial logit tiles meet in shared memory (fixed order);
//     warp 0 evaluates softmax / CE / delta for the 16 pairs (one pair per lane), writes logits / probs and publishes
//     (p - label) and delta for the stage on an mbarrier.
//   * BACKWARD warps 4-11 (each owns h/4 columns): ldmatrix.x4.trans + ONE mma per tile for dW (A = the tile transposed,
//     B = (p - label) of the 16 pairs split three ways like W; fp32 accumulators for the warp's columns stay in registers
//     for the whole kernel), and dx = delta * (W1 - W0) for its columns: 8 FMUL + 4 pack + one 16-byte store per lane and
//     row -- the same fp32 product, rounded once, as the CUDA-core kernel.  They release the stage right after the MMAs and
//     refill it themselves (four rows per warp) before the stores.
// Forward of group k+1 overlaps backward of group k: neither role's dependent chain (MMA -> exchange -> exp / log ->
// stores) is on the other's critical path.  (A first, homogeneous version -- every warp doing every phase in lock-step with a
// block barrier per group -- measured 130 us at h = 1024: 8 warps cannot hide a 9 000-cycle dependent chain per group.)
// Deterministic: fixed CTA -> row-group assignment, fixed-order block and grid reductions (no float atomics).
#include <cstdlib>

#include "softmax_head.cuh"

namespace ia {

// cycle counters of the last launch made with IA_HEAD_DEBUG & 16 (summed over CTAs / warps): [0] forward warps waiting for rows,
// [1] forward warps at their barrier, [2] loader warps waiting for a released stage, [3] forward warps total, [4] backward
// warps waiting for deltas, [5] backward warps total, [6] warp 0 softmax section
__device__ unsigned long long g_head_stats[8];

namespace hmma {
constexpr int FW = 4;            // forward warps
constexpr int BW = 8;            // backward warps
constexpr int RB = 16;           // pairs per stage (the M / K extent of one MMA)
constexpr int THREADS = (FW + BW) * 32;
constexpr int LOADERS = BW * 32;         // threads that issue the cp.async of a stage (the backward warps: they have the slack)
}  // namespace hmma

template <typename T> struct Mma16;
template <> struct Mma16<__nv_bfloat16> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack2(float lo, float hi) {      // two floats -> packed pair, round to nearest even
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  }
  static __device__ __forceinline__ float lo_value(uint32_t p) { return __uint_as_float(p << 16); }
  static __device__ __forceinline__ float hi_value(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }
};
template <> struct Mma16<__half> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  }
  static __device__ __forceinline__ float lo_value(uint32_t p) { return __low2float(*reinterpret_cast<const __half2*>(&p)); }
  static __device__ __forceinline__ float hi_value(uint32_t p) { return __high2float(*reinterpret_cast<const __half2*>(&p)); }
};

// Packed pair of term `level` of the 16-bit expansions a = hi + lo + lo2 (+ O(2^-24 |a|)), b likewise.  The level is given
// as three all-ones / all-zeros masks (a per-thread constant of the MMA fragment layout): branch-free.
template <typename T>
__device__ __forceinline__ uint32_t split_pair(float a, float b, uint32_t m0, uint32_t m1, uint32_t m2) {
  const uint32_t hi = Mma16<T>::pack2(a, b);
  const float a1 = a - Mma16<T>::lo_value(hi), b1 = b - Mma16<T>::hi_value(hi);
  const uint32_t lo = Mma16<T>::pack2(a1, b1);
  const float a2 = a1 - Mma16<T>::lo_value(lo), b2 = b1 - Mma16<T>::hi_value(lo);
  const uint32_t lo2 = Mma16<T>::pack2(a2, b2);
  return (hi & m0) | (lo & m1) | (lo2 & m2);
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// ring depth: what fits next to ~10 KB of other shared memory (h = 1024: 3 x 64.5 KB, 768: 4 x 48.5 KB, 512: 6 x 32.5 KB)
__host__ __device__ constexpr int stages_for(int tpw) { return tpw >= 16 ? 3 : (tpw >= 12 ? 4 : 6); }

// TPW = h / 64: 16-column tiles per BACKWARD warp (a forward warp has 2 * TPW).
template <typename T, typename G, int TPW>
__global__ void __launch_bounds__(hmma::THREADS, 1) softmax_head_mma_kernel(const HeadParams p) {
  using namespace hmma;
  constexpr int E = 8;                 // elements per 16-byte vector
  constexpr int H = TPW * 64;          // p.h (checked by the launcher)
  constexpr int H2 = 2 * H;
  constexpr uint32_t ROW_BYTES = H * 2u, PITCH = ROW_BYTES + 16u;
  constexpr uint32_t LABEL_OFF = 2u * RB * PITCH, STAGE_BYTES = LABEL_OFF + RB * 8u;   // [x rows | y rows | 16 int64 labels]
  constexpr int VPR = H / 8;           // 16-byte vectors per row
  constexpr int STAGES = stages_for(TPW);
  pdl_wait();       // launched with programmatic stream serialization (launch_pdl)
  float gscale = p.grad_scale;
  if (p.upstream != nullptr) {   // gradient recomputation with the upstream scalar folded in before the one rounding
    const float u = __ldg(p.upstream);
    if (p.upstream_skip_one && u == 1.0f) return;
    gscale *= u;
  }
  extern __shared__ __align__(128) uint8_t smem[];
  float* part = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);          // [2][FW][RB][8] partial logits
  float2* delta = reinterpret_cast<float2*>(part + 2 * FW * RB * 8);            // [STAGES][RB] (p - label, delta)
  float4* outbuf = reinterpret_cast<float4*>(delta + STAGES * RB);              // [STAGES][RB] (z0, z1, p0, p1) on their way out
  uint64_t* full = reinterpret_cast<uint64_t*>(outbuf + STAGES * RB);           // [STAGES] rows landed       (LOADERS arrivals)
  uint64_t* dready = full + STAGES;                                             // [STAGES] delta published   (1 arrival)
  uint64_t* empty = dready + STAGES;                                            // [STAGES] backward finished (BW arrivals)
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t n_groups = (p.n + RB - 1) / RB;
  const uint32_t smem_base = smem_u32(smem);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], LOADERS); mbar_init(&dready[s], 1); mbar_init(&empty[s], BW); }
    fence_mbar_init();
  }
  __syncthreads();

  // ldmatrix lane -> (row, column block) of a 16 x 16 tile.  plain: matrices (rows 0-7 | 8-15) x (cols 0-7 | 8-15) in the order
  // a0..a3 of a row-major A; transposed: A' = tile^T, so matrices 1 and 2 trade places.
  const int lm = lane >> 3, lr = lane & 7;
  float loss_acc = 0.f, db_acc = 0.f;   // forward warp 0, lanes 0-15

  // Stage fill by the BACKWARD warps (they wait for deltas a fifth of the time; the forward warps are the kernel's critical
  // resource: with the fill on forward warps 1-3 they were busy 3 500 cycles per stage): loader warp lw copies rows lw, lw+8, lw+16,
  // lw+24 of the 32-row stage (x rows 0-15, y rows 16-31), a warp instruction moves 512 contiguous bytes; loader 0 also brings
  // the 16 labels.  Every thread's arrival on the stage's mbarrier fires when its own copies have landed.  Rows past the end are
  // clamped to the last row: loaded, delta = 0, never stored.
  auto issue = [&](int lw, int stage, int64_t grp) {
    const uint32_t base = smem_base + (uint32_t)stage * STAGE_BYTES;
#pragma unroll
    for (int rr = 0; rr < 2 * RB / BW; ++rr) {
      const int r = lw + rr * BW;
      int64_t row = grp * RB + (r & 15);
      if (row > p.n - 1) row = p.n - 1;
      const uint4* src = reinterpret_cast<const uint4*>(r < 16 ? static_cast<const T*>(p.x) + row * p.ldx
                                                                : static_cast<const T*>(p.y) + row * p.ldy);
      const uint32_t dst = base + (uint32_t)r * PITCH + (uint32_t)lane * 16u;
#pragma unroll
      for (int v = 0; v < VPR / 32; ++v)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)v * 512u), "l"(src + lane + 32 * v) : "memory");
    }
    if (lw == 0 && lane < RB / 2) {
      // the group's 16 labels ride along (two per lane; bytes past the end of the array are zero-filled, never read)
      const int64_t l0 = grp * RB + 2 * lane;
      const int64_t have = p.n - l0;
      const uint32_t bytes = have >= 2 ? 16u : (have == 1 ? 8u : 0u);
      const int64_t* src = p.labels + (bytes ? l0 : 0);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(base + LABEL_OFF + (uint32_t)lane * 16u), "l"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[stage])) : "memory");
  };

  if (w < FW) {
    // =========================================================================================== forward warps
    const int side = w >> 1;                          // 0: this warp's columns belong to x, 1: to y
    const int col_base = (w & 1) * (H / 2);
    constexpr int TF = 2 * TPW;                       // tiles per forward warp
    // B fragments: MMA column n = 2 * level + class (n < 6) holds term `level` of W[class][col]; n = 6, 7 are zero
    uint32_t wb0[TF], wb1[TF];
    {
      const int cls = g & 1, level = g >> 1;
      const uint32_t m0 = level == 0 ? 0xffffffffu : 0u, m1 = level == 1 ? 0xffffffffu : 0u, m2 = level == 2 ? 0xffffffffu : 0u;
      const float* wrow = p.w + (size_t)cls * H2 + (size_t)side * H + col_base + 2 * t;
#pragma unroll
      for (int i = 0; i < TF; ++i) {
        const float2 q0 = __ldg(reinterpret_cast<const float2*>(wrow + i * 16));
        const float2 q1 = __ldg(reinterpret_cast<const float2*>(wrow + i * 16 + 8));
        wb0[i] = split_pair<T>(q0.x, q0.y, m0, m1, m2);
        wb1[i] = split_pair<T>(q1.x, q1.y, m0, m1, m2);
      }
    }
    const float bias0 = __ldg(p.b), bias1 = __ldg(p.b + 1);
    const uint32_t off_n = (uint32_t)((lr + (lm & 1) * 8) * (int)PITCH + (lm >> 1) * 16);
    const uint32_t tile_base = (uint32_t)side * RB * PITCH + (uint32_t)col_base * 2u;
    int it = 0;
    long long c_full = 0, c_bar = 0, c_soft = 0;
    const long long c_begin = clock64();
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ++it) {
      const int stage = it % STAGES;
      const uint32_t par = (uint32_t)(it / STAGES) & 1u;
      long long c0 = clock64();
      mbar_wait(&full[stage], par);
      c_full += clock64() - c0;
      const uint32_t tb = smem_base + (uint32_t)stage * STAGE_BYTES + tile_base + off_n;
      // partial logits of this warp's columns: D[16 pairs x 8] += tile . Wsplit.  The asm statements keep their order:
      // fragments are fetched one batch of four tiles ahead of the MMAs that consume them, and four independent accumulators
      // break the dependent-MMA chain.
      float dq[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) { dq[c][0] = 0.f; dq[c][1] = 0.f; dq[c][2] = 0.f; dq[c][3] = 0.f; }
      {
        uint32_t a[2][4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) ldsm_x4(tb + (uint32_t)c * 32u, a[0][c]);
#pragma unroll
        for (int b = 0; b < TF / 4; ++b) {
          if (b + 1 < TF / 4) {
#pragma unroll
            for (int c = 0; c < 4; ++c) ldsm_x4(tb + (uint32_t)((b + 1) * 4 + c) * 32u, a[(b + 1) & 1][c]);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) Mma16<T>::mma(dq[c], a[b & 1][c], wb0[b * 4 + c], wb1[b * 4 + c]);
        }
      }
      float* mypart = part + (size_t)((it & 1) * FW + w) * RB * 8;
      *reinterpret_cast<float2*>(mypart + g * 8 + 2 * t) =
          make_float2((dq[0][0] + dq[1][0]) + (dq[2][0] + dq[3][0]), (dq[0][1] + dq[1][1]) + (dq[2][1] + dq[3][1]));
      *reinterpret_cast<float2*>(mypart + (g + 8) * 8 + 2 * t) =
          make_float2((dq[0][2] + dq[1][2]) + (dq[2][2] + dq[3][2]), (dq[0][3] + dq[1][3]) + (dq[2][3] + dq[3][3]));
      c0 = clock64();
      named_bar_sync(1, FW * 32);
      c_bar += clock64() - c0;
      c0 = clock64();
      if (w == 0) {
        // ---- softmax / CE / delta of the 16 pairs, one pair per lane (lanes 16-31 mirror 0-15), fixed summation order.
        // This section is the kernel's only serial chain per group, so it touches no global memory: the labels came in
        // with the rows, logits / probs leave through shared memory (a backward warp stores them).
        const int pr = lane & 15;
        const int64_t row = grp * RB + pr;
        const bool live = row < p.n;
        const long long label = *reinterpret_cast<const long long*>(smem + (size_t)stage * STAGE_BYTES + LABEL_OFF + pr * 8);
        const float* src = part + (size_t)((it & 1) * FW) * RB * 8 + pr * 8;
        float4 lo4[FW];
        float2 hi2[FW];
#pragma unroll
        for (int ww = 0; ww < FW; ++ww) {
          lo4[ww] = *reinterpret_cast<const float4*>(src + ww * RB * 8);
          hi2[ww] = *reinterpret_cast<const float2*>(src + ww * RB * 8 + 4);
        }
        float z0 = 0.f, z1 = 0.f;
#pragma unroll
        for (int ww = 0; ww < FW; ++ww) {
          z0 += (lo4[ww].x + lo4[ww].z) + hi2[ww].x;      // class 0: columns 0, 2, 4 = hi, lo, lo2
          z1 += (lo4[ww].y + lo4[ww].w) + hi2[ww].y;      // class 1: columns 1, 3, 5
        }
        z0 += bias0; z1 += bias1;
        // two-class softmax with ONE exponential: e = exp(-|z1 - z0|) in (0, 1], p_max = 1 / (1 + e), p_min = e / (1 + e),
        // -log p[label] = log1p(e) + (label is the arg-max ? 0 : |z1 - z0|)
        const float dz = z1 - z0, ad = fabsf(dz);
        const float e = expf(-ad);
        const float rden = 1.0f / (1.0f + e);
        const float pmax = rden, pmin = e * rden;
        const bool one_max = dz >= 0.f;
        const float p0 = one_max ? pmin : pmax, p1 = one_max ? pmax : pmin;
        float du = 0.f;       // p - label, unscaled
        if (live) du = label != 0 ? -p0 : p1;
        if (lane < 16) {
          delta[stage * RB + pr] = make_float2(du, du * gscale);
          outbuf[stage * RB + pr] = make_float4(z0, z1, p0, p1);
          if (live) {
            loss_acc += log1pf(e) + (((label != 0) == one_max) ? 0.f : ad);     // -log softmax[label]
            if ((unsigned long long)label > 1ull) loss_acc = __int_as_float(0x7fc00000);   // label outside {0,1}: NaN, loudly
            db_acc += du * gscale;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&dready[stage]);     // release: the deltas (and, transitively, the rows) are visible
        c_soft += clock64() - c0;
      }
    }
    if ((p.load_mode & 16) && lane == 0) {
      atomicAdd(&g_head_stats[0], (unsigned long long)c_full);
      atomicAdd(&g_head_stats[1], (unsigned long long)c_bar);
      atomicAdd(&g_head_stats[3], (unsigned long long)(clock64() - c_begin));
      atomicAdd(&g_head_stats[6], (unsigned long long)c_soft);
    }
  } else {
    // =========================================================================================== backward warps
    const int bw = w - FW;
    const int side = bw >> 2;
    constexpr int SLICE = H / 4;                      // columns per backward warp
    const int col_base = (bw & 3) * SLICE;
    // dx = delta * (W1 - W0): this lane's 8 columns of the warp's slice.  V = 16 (h = 512): the two half-warps take alternate
    // rows; V = 24 (h = 768): lanes >= 24 idle.  (Whole-row stores -- a warp writing 2 KB rows, 32 weight differences per
    // lane -- measured 4 % slower: the extra registers made the compiler re-load the differences inside the loop.)
    constexpr int V = SLICE / 8;
    const int vec = (V == 16) ? (lane & 15) : lane;
    const int row_par = (V == 16) ? (lane >> 4) : 0;
    constexpr int ROW_STEP = (V == 16) ? 2 : 1;
    const bool vec_ok = vec < V;
    float wd[E];
    {
      const float* w0 = p.w + side * H + col_base + (vec_ok ? vec : 0) * E;
#pragma unroll
      for (int e = 0; e < E; e += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(w0 + e)), c = __ldg(reinterpret_cast<const float4*>(w0 + H2 + e));
        wd[e] = c.x - a.x; wd[e + 1] = c.y - a.y; wd[e + 2] = c.z - a.z; wd[e + 3] = c.w - a.w;
      }
      // pin the eight differences in registers: left alone, the compiler re-materialises them inside the store loop (read-only
      // loads may legally be repeated) and the backward warps then stall on L2 loads in their hot loop (23 % of the kernel's
      // stall samples sat on the multiply below)
#pragma unroll
      for (int e = 0; e < E; ++e) asm volatile("" : "+f"(wd[e]));
    }
    const uint32_t m0 = g == 0 ? 0xffffffffu : 0u, m1 = g == 1 ? 0xffffffffu : 0u, m2 = g == 2 ? 0xffffffffu : 0u;
    float acc[TPW][4];
#pragma unroll
    for (int i = 0; i < TPW; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
    const uint32_t off_t = (uint32_t)((lr + (lm >> 1) * 8) * (int)PITCH + (lm & 1) * 16);
    const uint32_t tile_base = (uint32_t)side * RB * PITCH + (uint32_t)col_base * 2u;
    G* const gout = static_cast<G*>(side ? p.dy : p.dx);
    const int64_t ldg_out = side ? p.lddy : p.lddx;
    for (int s = 0; s < STAGES; ++s) {
      const int64_t grp = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (grp < n_groups) issue(bw, s, grp);
    }
    int it = 0;
    long long c_wait = 0, c_empty = 0;
    const long long c_begin = clock64();
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ++it) {
      const int stage = it % STAGES;
      const uint32_t par = (uint32_t)(it / STAGES) & 1u;
      const long long c0 = clock64();
      mbar_wait(&dready[stage], par);
      c_wait += clock64() - c0;
      mbar_wait(&full[stage], par);      // already complete (the forward warps waited on it): makes the rows visible to THIS thread
      const float2* dl = delta + stage * RB;
      // ---- dW accumulators += tile^T . split(p - label): B fragment rows 2t, 2t+1, 2t+8, 2t+9, column g = term g (g < 3)
      const float2 u01a = dl[2 * t], u01b = dl[2 * t + 1], u23a = dl[2 * t + 8], u23b = dl[2 * t + 9];
      const uint32_t db0 = split_pair<T>(u01a.x, u01b.x, m0, m1, m2);
      const uint32_t db1 = split_pair<T>(u23a.x, u23b.x, m0, m1, m2);
      const uint32_t tb = smem_base + (uint32_t)stage * STAGE_BYTES + tile_base + off_t;
      if (!(p.load_mode & 4)) {
        uint32_t a[2][4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) ldsm_x4_trans(tb + (uint32_t)c * 32u, a[0][c]);
#pragma unroll
        for (int b = 0; b < TPW / 4; ++b) {
          if (b + 1 < TPW / 4) {
#pragma unroll
            for (int c = 0; c < 4; ++c) ldsm_x4_trans(tb + (uint32_t)((b + 1) * 4 + c) * 32u, a[(b + 1) & 1][c]);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) Mma16<T>::mma(acc[b * 4 + c], a[b & 1][c], db0, db1);
        }
      }
      // the dx pass needs only the 16 deltas: take them into registers and release the stage NOW, so that its refill is in
      // flight during the stores (the ring has three stages; the earlier the release, the more bytes are in flight)
      float dr[RB / ROW_STEP];
#pragma unroll
      for (int j = 0; j < RB / ROW_STEP; ++j) dr[j] = dl[j * ROW_STEP + row_par].y;
      if (bw == 0 && lane < RB) {          // logits / probs of the group leave here, off the forward warps' serial chain
        const float4 o = outbuf[stage * RB + lane];
        const int64_t row = grp * RB + lane;
        if (row < p.n) {
          if (p.logits) *reinterpret_cast<float2*>(p.logits + 2 * row) = make_float2(o.x, o.y);
          if (p.probs) *reinterpret_cast<float2*>(p.probs + 2 * row) = make_float2(o.z, o.w);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);        // this warp is done with the stage's rows and deltas
      {
        // refill: once ALL backward warps have released the stage (the forward warps finished with it before they published the
        // deltas), this warp brings its four rows of group it + STAGES -- in flight during the stores below
        const int64_t gn = grp + (int64_t)STAGES * gridDim.x;
        if (gn < n_groups) {
          const long long e0 = clock64();
          mbar_wait(&empty[stage], par);
          c_empty += clock64() - e0;
          issue(bw, stage, gn);
        }
      }
      // ---- dx (or dy) of this warp's columns: one 16-byte (fp32 gradients: two) store per lane and row
      if (gout != nullptr && !(p.load_mode & 8)) {
        G* dst = gout + (grp * RB + row_par) * ldg_out + col_base + vec * E;
        const int rows_left = (int)min((int64_t)RB, p.n - grp * RB);      // 16 except in the last group
#pragma unroll
        for (int j = 0; j < RB / ROW_STEP; ++j) {
          if (vec_ok && j * ROW_STEP + row_par < rows_left) {
            float gx[E];
#pragma unroll
            for (int e = 0; e < E; ++e) gx[e] = dr[j] * wd[e];
            Packer<G, E>::store(dst, gx);
          }
          dst += ROW_STEP * ldg_out;
        }
      }
    }
    if ((p.load_mode & 16) && lane == 0) {
      atomicAdd(&g_head_stats[4], (unsigned long long)c_wait);
      atomicAdd(&g_head_stats[2], (unsigned long long)c_empty);
      atomicAdd(&g_head_stats[5], (unsigned long long)(clock64() - c_begin));
    }
    // ---- dW partial of this CTA (columns of this warp) -> workspace
    float* pout = reinterpret_cast<float*>(static_cast<char*>(p.workspace) + kWorkspaceBytes) + (size_t)blockIdx.x * (H2 + 2);
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
      // thread (g, t) holds D[row g][cols 2t, 2t+1] and D[row g+8][...]; columns 0, 1, 2 are the three terms
      float v0 = acc[i][0] + acc[i][1], v1 = acc[i][2] + acc[i][3];
      v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
      v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
      if (t == 0) {
        const int c = side * H + col_base + i * 16 + g;
        pout[c] = v0 * gscale;
        pout[c + 8] = v1 * gscale;
      }
    }
  }

  // ---- loss and db of this CTA (forward warp 0, lanes 0-15) -> workspace / grid reduction
  __shared__ float s_loss;
  if (w == 0) {
    float l = lane < 16 ? loss_acc : 0.f, b = lane < 16 ? db_acc : 0.f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) { l += __shfl_xor_sync(0xffffffffu, l, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) {
      s_loss = l;
      float* pout = reinterpret_cast<float*>(static_cast<char*>(p.workspace) + kWorkspaceBytes) + (size_t)blockIdx.x * (H2 + 2);
      pout[H2] = b;
    }
  }
  __syncthreads();
  pdl_trigger();    // the dW finalize launch may be scheduled now (not earlier: its waiting CTAs would share the SMs with this kernel)
  grid_sum_finish((double)s_loss, p.workspace, p.loss_out, p.loss_scale);
}

bool softmax_head_mma_eligible(int dtype, const HeadParams& p) {
  return (dtype == IA_BF16 || dtype == IA_F16) && p.labels != nullptr && (p.h == 512 || p.h == 768 || p.h == 1024) && p.n >= 4096;
}

template <typename T, typename G, int TPW>
static int launch_mma(const HeadParams& p, cudaStream_t stream, float* dw, float* db) {
  using namespace hmma;
  constexpr int STAGES = stages_for(TPW);
  auto kernel = softmax_head_mma_kernel<T, G, TPW>;
  const size_t pitch = (size_t)p.h * 2 + 16;
  const size_t smem = (size_t)STAGES * (2 * RB * pitch + RB * 8) + sizeof(float) * 2 * FW * RB * 8 + (sizeof(float2) + sizeof(float4)) * STAGES * RB +
                      8 * 3 * STAGES;
  static size_t configured[kMaxDevices] = {};
  const int slot = device_slot();
  if (smem > configured[slot]) {
    IA_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[slot] = smem;
  }
  const int64_t groups = (p.n + RB - 1) / RB;
  const int grid = (int)(groups < sm_count() ? groups : sm_count());
  HeadParams pd = p;
  static const int debug = [] { const char* e = getenv("IA_HEAD_DEBUG"); return e ? atoi(e) : 0; }();   // timing experiments only:
  pd.load_mode = debug;                                    // 4 = skip the dW MMAs, 8 = skip the dx / dy stores (wrong results),
  if (debug & 16) {                                        // 16 = collect the cycle counters of ia_softmax_head_last_stats
    void* sp = nullptr;
    IA_CUDA_CHECK(cudaGetSymbolAddress(&sp, g_head_stats));
    IA_CUDA_CHECK(cudaMemsetAsync(sp, 0, sizeof(unsigned long long) * 8, stream));
  }
  IA_PDL_LAUNCH_CHECK(launch_pdl(kernel, grid, (int)THREADS, smem, stream, pd));
  if (dw || db) {
    const float* partials = reinterpret_cast<const float*>(static_cast<const char*>(p.workspace) + kWorkspaceBytes);
    const int h2 = 2 * p.h;
    IA_PDL_LAUNCH_CHECK(launch_pdl(softmax_head_finalize, (h2 + 1 + 7) / 8, 256, 0, stream, partials, grid, h2, dw, db, p.upstream, p.upstream_skip_one));
  }
  return IA_OK;
}

template <typename T, typename G>
static int launch_mma_tpw(const HeadParams& p, cudaStream_t stream, float* dw, float* db) {
  if (p.h == 512) return launch_mma<T, G, 8>(p, stream, dw, db);
  if (p.h == 768) return launch_mma<T, G, 12>(p, stream, dw, db);
  return launch_mma<T, G, 16>(p, stream, dw, db);
}

int launch_softmax_head_mma(int dtype, int grad_dtype, const HeadParams& p, cudaStream_t stream, float* dw, float* db) {
  if (dtype == IA_BF16 && grad_dtype == IA_BF16) return launch_mma_tpw<__nv_bfloat16, __nv_bfloat16>(p, stream, dw, db);
  if (dtype == IA_BF16) return launch_mma_tpw<__nv_bfloat16, float>(p, stream, dw, db);
  if (dtype == IA_F16 && grad_dtype == IA_F16) return launch_mma_tpw<__half, __half>(p, stream, dw, db);
  if (dtype == IA_F16) return launch_mma_tpw<__half, float>(p, stream, dw, db);
  set_error("softmax head (tensor-core kernel): unsupported dtype %d", dtype);
  return IA_ERR_UNSUPPORTED;
}

}  // namespace ia

extern "C" int ia_softmax_head_last_stats(uint64_t* out8) {
  if (out8 == nullptr) { ia::set_error("bad arguments"); return IA_ERR_INVALID; }
  IA_CUDA_CHECK(cudaMemcpyFromSymbol(out8, ia::g_head_stats, sizeof(unsigned long long) * 8));
  return IA_OK;
}
