// Catalog retrieval: all-pairs scores + per-query top-k (see include/ia_b200.h for the semantics).
//
//  * inner / cosine on bf16|fp16 catalogs: a genuine dense contraction -> tcgen05 tensor cores.
//    Persistent, warp-specialised CTAs: warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle),
//    warp 1 = single-thread tcgen05.mma issuer (M=128 queries x N=256 catalog rows, K=16 per MMA, fp32
//    accumulators double-buffered in the 512 TMEM columns), warps 2-5 = epilogue: tcgen05.ld the
//    accumulator rows (one query per thread), fused normalise (x 1/|c| x 1/|q|), threshold filter against
//    the running k-th best and a rare-path append/merge into the per-query sorted key list.  The Q x C
//    score matrix is never materialised.
//  * l1 / l2 (no MMA form because of |.|) and fp32 inputs: shared-memory-tiled CUDA-core kernel with the
//    same top-k epilogue.
//  * a CTA owns (query tile, catalog split); split results are merged by a small kernel.  All CTAs working
//    on the same queries share their k-th best through tau_global so thresholds tighten across splits.
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <new>
#include <type_traits>

#include "retrieval.cuh"

namespace ia {

struct RetrParams {
  int q_rows;
  int64_t c_rows;
  int d;
  int k;
  int n_qt;             // query tiles
  int n_splits;         // catalog splits
  int tiles_per_split;  // catalog tiles per split
  int n_tiles;          // catalog tiles in total
  int kblocks;          // K blocks of 64 (tensor-core kernel)
  const float* cinv;    // [c_rows] 1/max(|c|,eps)   (cosine)
  const float* qinv;    // [q_rows]                  (cosine)
  uint64_t* lists;      // [n_splits*n_qt][BM][kListCap]
  uint64_t* overflow;   // [n_splits*n_qt][BM][kBufSlots] keys still in the append buffers when an item ended
  uint32_t* tau_global; // [n_qt*BM] best published k-th goodness per query (0 = none)
  uint32_t* done;       // [n_items*4] set when a warp's quarter of an item's lists is final
  unsigned long long* stats;  // [8] appends, compactions, rare groups, rare blocks (telemetry)
  float dist_eps;       // l1 / l2: eps added to every difference (nn.PairwiseDistance 1e-6; 0 for plain norms)
  int dist_squared;     // l2: rank and report the squared distance (torchkge l2_dissimilarity)
  uint32_t row_base;    // global row id of catalog row 0 (shard offset)
  int flags;            // tuning switches for in-run A/B measurements (env IA_RETR_FLAGS, default 14): bit1 early tau
                        // load, bit2 finished-split bound, bit3 paced merges (<= 2 lanes per tile after release);
                        // bit5 (32, set by probe passes): items neither read nor publish thresholds of other items --
                        // every (query tile, row group) finds ITS OWN exact top-k
};

// A valid lower bound on the final k-th best key of a query from the FINISHED splits of its query tile:
// if F splits are final and each holds >= r = ceil(k/F) keys >= v (v = the smallest of their r-th best keys),
// then F*r >= k candidates of disjoint chunks are >= v, so no key <= v can be in the global top-k of the
// remaining chunks.  As F grows this approaches the k-th best of everything scanned so far.
__device__ __forceinline__ uint64_t finished_splits_bound(const RetrParams& p, int qt, int my_split, int quarter,
                                                          int row_local, int tile_rows) {
  const int lane = threadIdx.x & 31;
  unsigned fin_lo = 0;   // up to 32 splits tracked (enough: the bound saturates quickly)
  const int ns = p.n_splits < 32 ? p.n_splits : 32;
  if (lane < ns && lane != my_split) {
    const uint32_t* flag = p.done + ((size_t)(lane * p.n_qt + qt)) * 4 + quarter;
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    fin_lo = v != 0;
  }
  unsigned fin = __ballot_sync(kFull, fin_lo != 0);
  const int f = __popc(fin);
  if (f < 2) return 0ull;            // one finished split gives its k-th best: already in tau_global
  const int r = (p.k + f - 1) / f;
  uint64_t bound = ~0ull;
  while (fin) {
    const int sp = __ffs(fin) - 1;
    fin &= fin - 1;
    const uint64_t* list = p.lists + (((size_t)sp * p.n_qt + qt) * tile_rows + row_local) * kListCap;
    const uint64_t key = __ldcg(list + (r - 1));
    bound = key < bound ? key : bound;
  }
  return bound;                       // 0 if any finished list is shorter than r: no bound
}

// v[c] for a runtime c in [0,32): binary select tree (31 SEL), registers stay registers
__device__ __forceinline__ float select32(const float (&v)[32], int c) {
  float a[16], b8[8], c4[4], d2[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (c & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b8[i] = (c & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c4[i] = (c & 4) ? b8[2 * i + 1] : b8[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d2[i] = (c & 8) ? c4[2 * i + 1] : c4[2 * i];
  return (c & 16) ? d2[1] : d2[0];
}

// max of three floats in one instruction (sm_100: FMNMX3); NaN operands are ignored like fmaxf does
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// Largest-safe pre-filter threshold t' with:  (x * qinv >= thr)  =>  (x >= t')  for every float x, qinv > 0.
// t' = thr / qinv pulled down by 4 ulp-ish relative steps; -inf stays -inf.  qinv == 0 (padded query) -> +inf.
__device__ __forceinline__ float prefilter_threshold(float thr, float qinv) {
  if (!(qinv > 0.f)) return INFINITY;
  const float t = thr / qinv;
  return t - fabsf(t) * 4.8e-7f - 1e-37f;
}

// =============================================================================== tensor-core kernel
namespace tc {
constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int THREADS = 192;
constexpr int BUF_BYTES = BM * kBufPitch * 8;
constexpr int CINV_BYTES = 2 * BN * 4;
// CTAS = 1: one CTA stages the 128-query tile and the whole 256-row catalog tile per k-block (48 KB, ring of 4).
// CTAS = 2: a CTA PAIR (cluster of two SMs of one TPC) runs ONE 256 x 256 MMA per k-step (cta_group::2): each CTA stages ITS
// 128-query tile and HALF of the catalog tile (32 KB, ring of 6) and ends up with the 128 x 256 accumulator of its own queries
// in its own tensor memory -- a third less shared-memory fill per FLOP, half the catalog re-stream from L2 per query tile.
template <int CTAS> struct Cfg {
  static constexpr int B_ROWS = BN / CTAS;                       // catalog rows this CTA stages per k-block
  static constexpr int STAGE_BYTES = A_BYTES + B_ROWS * BK * 2;
  static constexpr int STAGES = CTAS == 1 ? 4 : 6;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BUF_BYTES + CINV_BYTES + 256 + 1024;  // + barriers + align slack
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
};
}  // namespace tc
constexpr bool kPairDefault = false;   // the pair form is selected by default where use_pair() says it pays

// KREG > 0: "small k" mode (k <= KREG): every thread keeps its query's top-KREG keys sorted in REGISTERS -- no append
// buffers, no warp merges, no lists in L2 until the item ends.  Used for k <= 16 (probe passes, top-10 workloads).
template <bool COSINE, int AB_FORMAT, int KREG, int CTAS>
__global__ void __launch_bounds__(tc::THREADS, 1)
retrieve_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                   const RetrParams p) {
  using namespace tc;
  constexpr int STAGES = Cfg<CTAS>::STAGES, STAGE_BYTES = Cfg<CTAS>::STAGE_BYTES, B_ROWS = Cfg<CTAS>::B_ROWS;
  constexpr int SCR_BYTES = 0;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzled operand tiles, keeping the pointer's shared-memory provenance
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* buf = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  float* cinv_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + BUF_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + BUF_BYTES + CINV_BYTES + SCR_BYTES);
  uint64_t* full_bar = bars;                // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;      // [STAGES] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;  // [2] MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;     // [2] epilogue -> MMA (pair: the epilogue warps of BOTH CTAs -> the leader's MMA thread)
  uint64_t* cfull_bar = tempty_bar + 2;     // [2] 1/|c| tile landed (bulk copy issued by the MMA thread)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(cfull_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // flags bit 4: the two single-thread feeder warps take the HIGHEST warp ids (the issue arbiter favours them)
  const int w_tma = (p.flags & 16) ? 4 : 0, w_mma = (p.flags & 16) ? 5 : 1;
  if (warp == w_tma && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4 * CTAS); mbar_init(&cfull_bar[a], 1); }
    fence_mbar_init();
  }
  if (warp == w_mma) { if (CTAS == 1) tmem_alloc(tmem_ptr, 512); else tmem_alloc_pair(tmem_ptr, 512); }
  tc_fence_before();
  if (CTAS == 1) __syncthreads(); else cluster_sync();     // pair: the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // Work items: (query tile [pair], catalog split).  The CTAs of a pair walk the same items; CTA `rank` owns query tile
  // CTAS * slot + rank (an odd tile count leaves the last pair's second CTA a PHANTOM tile: operands zero-filled by TMA, nothing
  // read from or written to global memory).  Lists / flags are indexed by item = split * n_qt + qt as in the single-CTA form.
  const int rank = CTAS == 1 ? 0 : (int)cluster_ctarank();
  const int unit = blockIdx.x / CTAS, n_units = gridDim.x / CTAS;
  const int n_slots = (p.n_qt + CTAS - 1) / CTAS;
  const int n_work = n_slots * p.n_splits;

  if (warp == w_tma) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = unit; work < n_work; work += n_units) {
        const int qt = (work % n_slots) * CTAS + rank, split = work / n_slots;
        const int t0 = split * p.tiles_per_split;
        const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
        for (int t = t0; t < t1; ++t) {
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* a_s = smem + stage * STAGE_BYTES;
            if (CTAS == 1) {
              mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
              tma_load_2d(a_s, &tmap_q, &full_bar[stage], kb * BK, qt * BM);
              tma_load_2d(a_s + A_BYTES, &tmap_c, &full_bar[stage], kb * BK, t * BN);
            } else {
              // both CTAs' bytes are counted on the LEADER's barrier (the one its MMA thread waits on); the leader arms it
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], CTAS * STAGE_BYTES);
              const uint32_t bar0 = mapa_u32(smem_u32(&full_bar[stage]), 0);
              tma_load_2d_pair(a_s, &tmap_q, bar0, kb * BK, qt * BM);
              tma_load_2d_pair(a_s + A_BYTES, &tmap_c, bar0, kb * BK, t * BN + rank * B_ROWS);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == w_mma) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc(BM * CTAS, BN, AB_FORMAT);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      long long c_tempty = 0, c_full = 0;
      const long long c_begin = clock64();
      for (int work = unit; work < n_work; work += n_units) {
        const int split = work / n_slots;
        const int t0 = split * p.tiles_per_split;
        const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
        for (int t = t0; t < t1; ++t) {
          long long c0 = clock64();
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);   // epilogue has drained this accumulator (and its 1/|c| tile)
          c_tempty += clock64() - c0;
          tc_fence_after();
          if (COSINE) {   // this tile's inverse catalog norms ride along: 1 KB bulk copy, lands long before the MMAs finish
            if (CTAS == 1) {
              mbar_arrive_expect_tx(&cfull_bar[acc], BN * 4);
              bulk_load_1d(cinv_s + acc * BN, p.cinv + (size_t)t * BN, BN * 4, &cfull_bar[acc]);
            } else {
#pragma unroll
              for (int c = 0; c < CTAS; ++c) {   // one copy into each CTA of the pair, armed and counted on that CTA's barrier
                const uint32_t bar_c = mapa_u32(smem_u32(&cfull_bar[acc]), c);
                mbar_arrive_expect_tx_cluster(bar_c, BN * 4);
                bulk_load_1d_cluster(mapa_u32(smem_u32(cinv_s + acc * BN), c), p.cinv + (size_t)t * BN, BN * 4, bar_c);
              }
            }
          }
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            c0 = clock64();
            mbar_wait(&full_bar[stage], phase);
            c_full += clock64() - c0;
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t a_desc = umma_smem_desc_sw128(a_addr);
            const uint64_t b_desc = umma_smem_desc_sw128(a_addr + A_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < BK / 16; ++k4) {   // +32 B (>>4 = 2) per K=16 step inside the swizzle atom
              if (CTAS == 1) umma_f16(d_tmem, a_desc + 2 * k4, b_desc + 2 * k4, idesc, (kb | k4) != 0);
              else umma_f16_pair(d_tmem, a_desc + 2 * k4, b_desc + 2 * k4, idesc, (kb | k4) != 0);
            }
            // smem slot free once these MMAs retire (pair: in both CTAs -- each producer waits on its own barrier)
            if (CTAS == 1) umma_commit(&empty_bar[stage]); else umma_commit_pair(&empty_bar[stage], 3);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (CTAS == 1) umma_commit(&tfull_bar[acc]); else umma_commit_pair(&tfull_bar[acc], 3);   // accumulator complete
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
      if (p.stats != nullptr && (p.flags & 128)) {   // flags bit 7: the MMA thread's waits replace the rare-path counters
        atomicAdd(p.stats + 2, (unsigned long long)c_tempty);
        atomicAdd(p.stats + 3, (unsigned long long)c_full);
        atomicAdd(p.stats + 7, (unsigned long long)(clock64() - c_begin));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: fused normalise + top-k
    const int e = warp & 3;                 // TMEM lane quarter this warp may access
    const int row_local = e * 32 + lane;    // query row inside the tile == TMEM lane
    uint64_t* buf_warp = buf + (size_t)(e * 32) * kBufPitch;
    TopKStats stats{0u, 0u, 0u, 0u};
    long long cyc_wait = 0, cyc_compact = 0, cyc_rare = 0;
    const long long cyc_begin = clock64();
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t tempty_leader = CTAS == 1 ? 0u : mapa_u32(smem_u32(tempty_bar), 0);
    for (int work = unit; work < n_work; work += n_units) {
      const int qt = (work % n_slots) * CTAS + rank, split = work / n_slots;
      const bool qt_ok = CTAS == 1 || qt < p.n_qt;     // false: phantom tile of an odd tile count (warp-uniform)
      const int item = split * p.n_qt + qt;
      const int t0 = split * p.tiles_per_split;
      const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
      const int q_row = qt * BM + row_local;
      const bool q_ok = q_row < p.q_rows;              // never true in a phantom tile
      float qinv = 1.f;
      if (COSINE) qinv = q_ok ? __ldg(p.qinv + q_row) : 0.f;
      uint64_t* lists_warp = p.lists + ((size_t)item * BM + e * 32) * kListCap;
      if (qt_ok)
        for (int i = lane; i < 32 * kListCap; i += 32) lists_warp[i] = 0ull;
      __syncwarp();
      uint32_t* tau_warp = p.tau_global + qt * BM + e * 32;
      TopKThread st{0ull, 0};
      if ((p.flags & 36) == 4 && qt_ok) st.thr_key = finished_splits_bound(p, qt, split, e, row_local, BM);
      uint64_t top[KREG > 0 ? KREG : 1];
#pragma unroll
      for (int i = 0; i < (KREG > 0 ? KREG : 1); ++i) top[i] = 0ull;
      uint32_t tau_published = 0;

      for (int t = t0; t < t1; ++t) {
        const int64_t j0 = (int64_t)t * BN;
        const float* cs = cinv_s + acc * BN;
        // what other CTAs have published for this query: issue the (L2-latency) load now, consume it after the waits
        uint32_t tau_seen = 0;
        if ((p.flags & 34) == 2 && qt_ok) tau_seen = *reinterpret_cast<volatile uint32_t*>(tau_warp + lane);
        const long long w0 = clock64();
        if (COSINE) mbar_wait(&cfull_bar[acc], acc_phase);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        cyc_wait += clock64() - w0;
        if (!(p.flags & 34) && qt_ok) tau_seen = *reinterpret_cast<volatile uint32_t*>(tau_warp + lane);
        {
          const uint64_t gk = (uint64_t)tau_seen << 32;
          if (gk > st.thr_key) st.thr_key = gk;
        }
        float thr_f = st.thr_key ? score_of_goodness<true>((uint32_t)(st.thr_key >> 32)) : -INFINITY;
        const uint32_t taddr = tmem_base + ((uint32_t)(e * 32) << 16) + acc * BN;
        // pre-filter threshold in the (acc * 1/|c|) domain: a hair below thr_f / (1/|q|) so that rounding can only
        // let extra candidates through; the exact test on (acc * 1/|c|) * 1/|q| follows in the rare path
        float thr_pre = thr_f;
        if (COSINE) thr_pre = prefilter_threshold(thr_f, qinv);
#pragma unroll 1
        for (int g = 0; g < ((p.flags & 64) ? 0 : BN / 32); ++g) {   // flags bit 6: skip the scan (mainloop ceiling; wrong results)
          uint32_t r[32];
          tmem_ld_32x32(taddr + g * 32, r);
          tmem_ld_wait();
          // fast path: scale by 1/|c| and take the MAXIMUM of the thread's 32 scores (three-input max: 16 instructions where a
          // compare + mask bit per score was 64): only when some lane's maximum passes its threshold is the 32-bit hit mask
          // built.  Under the power cap every epilogue instruction is paid for in tensor clock.
          float v[32];
          const float4* cs4 = reinterpret_cast<const float4*>(cs + g * 32);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const int c = 4 * c4;
            v[c + 0] = __uint_as_float(r[c + 0]); v[c + 1] = __uint_as_float(r[c + 1]);
            v[c + 2] = __uint_as_float(r[c + 2]); v[c + 3] = __uint_as_float(r[c + 3]);
            if (COSINE) {
              const float4 ci = cs4[c4];
              v[c + 0] *= ci.x; v[c + 1] *= ci.y; v[c + 2] *= ci.z; v[c + 3] *= ci.w;
            }
          }
          float mx = fmax3(v[0], v[1], v[2]);
#pragma unroll
          for (int c = 3; c + 1 < 32; c += 2) mx = fmax3(mx, v[c], v[c + 1]);     // v[3..30]
          mx = fmaxf(mx, v[31]);
          unsigned mask = 0;
          if (__any_sync(kFull, q_ok && mx >= thr_pre)) {
#pragma unroll
            for (int c = 0; c < 32; ++c) mask |= (v[c] >= thr_pre ? 1u : 0u) << c;
            if (!q_ok) mask = 0;
          }
          if (__any_sync(kFull, mask != 0)) {
            stats.rare_groups++;
            const long long r0 = clock64();
            long long c_in = 0;
            // rare path: every lane walks its own hits (usually one), all lanes in parallel
            while (__any_sync(kFull, mask != 0)) {
              stats.rare_blocks++;
              if (mask != 0) {
                const int c = __ffs(mask) - 1;
                mask &= mask - 1;
                const float vc = select32(v, c);
                const float s = COSINE ? vc * qinv : vc;
                const int64_t j = j0 + g * 32 + c;
                if (s >= thr_f && j < p.c_rows) {
                  const uint64_t key = make_key<true>(s, p.row_base + (uint32_t)j);
                  if (key > st.thr_key) {
                    stats.appends++;
                    if (KREG > 0) {
                      // sorted insertion into the register list (compare-exchange chain), new k-th best at once
                      uint64_t x = key;
                      uint64_t kth = 0;
#pragma unroll
                      for (int i = 0; i < KREG; ++i) {
                        const uint64_t hi = u64max(top[i], x);
                        x = u64min(top[i], x);
                        top[i] = hi;
                        if (i == p.k - 1) kth = hi;
                      }
                      if (kth > st.thr_key) {
                        st.thr_key = kth;
                        thr_f = score_of_goodness<true>((uint32_t)(kth >> 32));
                        thr_pre = COSINE ? prefilter_threshold(thr_f, qinv) : thr_f;
                      }
                    } else {
                      buf_warp[lane * kBufPitch + st.cnt] = key;
                      st.cnt++;
                    }
                  }
                }
              }
              if (KREG == 0 && __any_sync(kFull, st.cnt == kBufSlots)) {
                __syncwarp();
                const long long c0 = clock64();
                warp_compact(st, (p.flags & 8) ? kBufSlots : kBufSlots / 2, p.k, buf_warp, lists_warp, tau_warp, stats);
                c_in += clock64() - c0;
                thr_f = st.thr_key ? score_of_goodness<true>((uint32_t)(st.thr_key >> 32)) : -INFINITY;
                thr_pre = COSINE ? prefilter_threshold(thr_f, qinv) : thr_f;
                // drop hits the tighter threshold already rules out
                unsigned m2 = 0;
#pragma unroll
                for (int c = 0; c < 32; ++c) m2 |= (v[c] >= thr_pre ? 1u : 0u) << c;
                mask &= m2;
              }
            }
            cyc_compact += c_in;
            cyc_rare += clock64() - r0 - c_in;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                                // accumulator released: the MMA warp may overwrite it
          if (CTAS == 1) mbar_arrive(&tempty_bar[acc]); else mbar_arrive_cluster(tempty_leader + (uint32_t)acc * 8u);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        // list maintenance OUTSIDE the accumulator's critical section: half-full buffers are merged now, while the
        // tensor core works on the next tiles, so that buffers rarely fill up (and force a merge) mid-tile
        if (KREG == 0 && __any_sync(kFull, st.cnt >= kBufSlots / 2)) {
          const long long c0 = clock64();
          warp_compact(st, kBufSlots / 2, p.k, buf_warp, lists_warp, tau_warp, stats, (p.flags & 8) ? 2 : 32);
          cyc_compact += clock64() - c0;
        }
        if (KREG > 0) {   // publish an improved k-th best for the other CTAs working on these queries (once per tile)
          const uint32_t w = (uint32_t)(st.thr_key >> 32);
          // (thr_key may also hold a bound learnt from others; re-publishing that is harmless: atomicMax)
          if (w > tau_published && q_ok && !(p.flags & 32)) { atomicMax(tau_warp + lane, w); tau_published = w; }
        }
      }
      __syncwarp();
      if (!qt_ok) continue;     // phantom tile: nothing to publish
      if (KREG > 0) {
        // the register list becomes the item's list (positions >= KREG stay zero from the initialisation above)
        uint64_t* mine = lists_warp + (size_t)lane * kListCap;
#pragma unroll
        for (int i = 0; i < KREG; ++i) mine[i] = top[i];
      } else {
        // no flush merges: the (few) keys still buffered go to the item's overflow slots; merge_lists_kernel, which
        // has the whole GPU to itself, folds them in.  Lists stay valid lower bounds for finished_splits_bound.
        uint64_t* ov = p.overflow + ((size_t)item * BM + row_local) * kBufSlots;
        for (int i = 0; i < kBufSlots; ++i) ov[i] = i < st.cnt ? buf_warp[lane * kBufPitch + i] : 0ull;
      }
      // publish: this quarter of the item's lists is final (release after every lane's list writes)
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        uint32_t* flag = p.done + (size_t)item * 4 + e;
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(1u) : "memory");
      }
    }
    if (p.stats != nullptr) {
      const unsigned a = __reduce_add_sync(kFull, stats.appends);
      if (lane == 0) {
        atomicAdd(p.stats + 0, (unsigned long long)a);
        atomicAdd(p.stats + 1, (unsigned long long)stats.compactions);
        if (!(p.flags & 128)) {
          atomicAdd(p.stats + 2, (unsigned long long)stats.rare_groups);
          atomicAdd(p.stats + 3, (unsigned long long)stats.rare_blocks);
        }
        atomicAdd(p.stats + 4, (unsigned long long)cyc_wait);
        atomicAdd(p.stats + 5, (unsigned long long)cyc_compact);
        atomicAdd(p.stats + 6, (unsigned long long)(clock64() - cyc_begin));
        if (!(p.flags & 128)) atomicAdd(p.stats + 7, (unsigned long long)cyc_rare);
      }
    }
  }

  tc_fence_before();
  if (CTAS == 1) __syncthreads(); else cluster_sync();   // pair: neither CTA leaves while the other may still touch its memory
  if (warp == w_mma) {
    __syncwarp();
    tc_fence_after();
    if (CTAS == 1) tmem_dealloc(tmem_base, 512); else tmem_dealloc_pair(tmem_base, 512);
  }
}

// =============================================================================== CUDA-core kernel
namespace simt {
constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256;
constexpr int AP = BM + 4;                        // pitch (floats) of the k-major staging tiles; multiple of 4: LDS.128
constexpr int STAGE_FLOATS = 2 * BK * AP;         // one stage: A tile [BK][AP] + B tile [BK][AP]
constexpr int SS = BM * (BN + 1);                 // score tile of the scan, aliases the two staging stages
constexpr int TILE_FLOATS = (2 * STAGE_FLOATS) > SS ? (2 * STAGE_FLOATS) : SS;
constexpr int SMEM_BYTES = TILE_FLOATS * 4 + BM * kBufPitch * 8;
static_assert(BM == BN, "one loader serves both tiles");
}  // namespace simt

// 8 consecutive elements of a row starting at column k0 as floats; `fill` where the row or the column does not exist.
// VEC: rows are 16-byte aligned and d % 8 == 0, so one (16-bit types) or two (fp32) 128-bit loads do it.
template <typename T, bool VEC>
__device__ __forceinline__ void load8(const T* __restrict__ base, int64_t ld, int64_t row, int64_t rows, int k0, int d, float fill,
                                      float (&out)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) out[j] = fill;
  if (row >= rows) return;
  const T* src = base + row * ld + k0;
  if (VEC) {
    if (k0 < d) {   // d % 8 == 0 and k0 % 8 == 0: all eight columns exist
      if (sizeof(T) == 4) {
        const uint4 u0 = ldg_stream(reinterpret_cast<const uint4*>(src));
        const uint4 u1 = ldg_stream(reinterpret_cast<const uint4*>(src) + 1);
        out[0] = __uint_as_float(u0.x); out[1] = __uint_as_float(u0.y); out[2] = __uint_as_float(u0.z); out[3] = __uint_as_float(u0.w);
        out[4] = __uint_as_float(u1.x); out[5] = __uint_as_float(u1.y); out[6] = __uint_as_float(u1.z); out[7] = __uint_as_float(u1.w);
      } else {
        unpack<T>(ldg_stream(reinterpret_cast<const uint4*>(src)), out);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (k0 + j < d) out[j] = to_float<T>(src[j]);
  }
}

// CUDA-core all-pairs kernel (l1 / l2, and any measure in fp32): 128 x 128 scores per CTA tile, 8 x 8 per thread, K in
// blocks of 16 through two shared-memory stages (the next block's global loads are in flight while the current one is
// multiplied), k-major tiles so that a thread's eight query values and eight catalog values are two LDS.128 each.
template <typename T, int MEASURE, bool VEC>
__global__ void __launch_bounds__(simt::THREADS, 2)
retrieve_simt_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ cat, int64_t ldc, const RetrParams p) {
  using namespace simt;
  constexpr bool DESC = (MEASURE == IA_INNER || MEASURE == IA_COSINE);
  constexpr bool DIST = !DESC;
  extern __shared__ uint8_t smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);   // 2 stages x (A[BK][AP] | B[BK][AP])
  float* Ss = reinterpret_cast<float*>(smem_raw);      // [BM][BN+1] (aliases the staging stages)
  uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw + TILE_FLOATS * 4);
  TopKStats stats{0u, 0u, 0u, 0u};
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ty = tid >> 4, tx = tid & 15;
  const int n_items = p.n_qt * p.n_splits;
  const float eps = p.dist_eps;
  // loader role: thread -> (row of the tile, half of the K block)
  const int l_row = tid & (BM - 1), l_k = (tid >> 7) * 8;
  const int kblocks = (p.d + BK - 1) / BK;

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int qt = item % p.n_qt, split = item / p.n_qt;
    const int t0 = split * p.tiles_per_split;
    const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
    const int q0 = qt * BM;
    // top-k state: threads 0..127 own one query each (warps 0-3, like the four lane quarters of the tensor-core kernel)
    TopKThread st{0ull, 0};
    uint64_t* lists_warp = p.lists + ((size_t)item * BM + (warp & 3) * 32) * kListCap;
    uint64_t* buf_warp = buf + (size_t)((warp & 3) * 32) * kBufPitch;
    uint32_t* tau_warp = p.tau_global + qt * BM + (warp & 3) * 32;
    float qinv = 1.f;
    if (warp < 4) {
      for (int i = lane; i < 32 * kListCap; i += 32) lists_warp[i] = 0ull;
      if (MEASURE == IA_COSINE) qinv = (q0 + tid < p.q_rows) ? __ldg(p.qinv + q0 + tid) : 0.f;
      __syncwarp();
      st.thr_key = finished_splits_bound(p, qt, split, warp, tid, BM);
    }

    for (int t = t0; t < t1; ++t) {
      const int64_t j0 = (int64_t)t * BN;
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

      float ra[8], rb[8];
      // padded k / rows of the catalog: c = eps so that (0 - eps) + eps == 0 adds nothing to an l1 / l2 distance
      load8<T, VEC>(q, ldq, q0 + l_row, p.q_rows, l_k, p.d, 0.f, ra);
      load8<T, VEC>(cat, ldc, j0 + l_row, p.c_rows, l_k, p.d, DIST ? eps : 0.f, rb);
      __syncthreads();   // previous tile's scan has finished with Ss (aliases the stages)
      for (int kb = 0; kb < kblocks; ++kb) {
        float* As = tiles + (kb & 1) * STAGE_FLOATS;
        float* Bs = As + BK * AP;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          As[(l_k + j) * AP + l_row] = ra[j];
          Bs[(l_k + j) * AP + l_row] = rb[j];
        }
        __syncthreads();   // stage kb & 1 is complete; the other stage is free (its readers passed the previous barrier)
        if (kb + 1 < kblocks) {   // next K block: global loads in flight during the multiply below
          load8<T, VEC>(q, ldq, q0 + l_row, p.q_rows, (kb + 1) * BK + l_k, p.d, 0.f, ra);
          load8<T, VEC>(cat, ldc, j0 + l_row, p.c_rows, (kb + 1) * BK + l_k, p.d, DIST ? eps : 0.f, rb);
        }
#pragma unroll 4   // 4 x 200 instructions: the fully unrolled block (51 KB of code) would not stay in the instruction cache
        for (int kk = 0; kk < BK; ++kk) {
          const float4 a0 = *reinterpret_cast<const float4*>(&As[kk * AP + 4 * ty]);
          const float4 a1 = *reinterpret_cast<const float4*>(&As[kk * AP + 64 + 4 * ty]);
          const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk * AP + 4 * tx]);
          const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk * AP + 64 + 4 * tx]);
          const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (DESC) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
              else {
                const float dd = a[i] - b[j] + eps;   // nn.PairwiseDistance: eps added to the difference
                if (MEASURE == IA_L1) acc[i][j] += fabsf(dd);
                else acc[i][j] = fmaf(dd, dd, acc[i][j]);
              }
            }
        }
      }
      __syncthreads();   // all warps done with the stages before Ss (alias) is written
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = (i < 4 ? 0 : 64) + 4 * ty + (i & 3);
#pragma unroll
        for (int j = 0; j < 8; ++j) Ss[r * (BN + 1) + (j < 4 ? 0 : 64) + 4 * tx + (j & 3)] = acc[i][j];
      }
      __syncthreads();

      if (warp < 4) {
        const bool q_ok = q0 + tid < p.q_rows;
        {
          const uint64_t gk = (uint64_t)(*reinterpret_cast<volatile uint32_t*>(tau_warp + lane)) << 32;
          if (gk > st.thr_key) st.thr_key = gk;
        }
        float thr_f = st.thr_key ? score_of_goodness<DESC>((uint32_t)(st.thr_key >> 32)) : (DESC ? -INFINITY : INFINITY);
#pragma unroll 1
        for (int c = 0; c < BN; ++c) {
          float s = Ss[tid * (BN + 1) + c];
          const int64_t j = j0 + c;
          if (MEASURE == IA_L2 && !p.dist_squared) s = sqrtf(s);
          if (MEASURE == IA_COSINE) s = (s * ((j < p.c_rows) ? __ldg(p.cinv + j) : 0.f)) * qinv;
          const bool pass = DESC ? (s >= thr_f) : (s <= thr_f);
          if (q_ok && j < p.c_rows && pass) {
            const uint64_t key = make_key<DESC>(s, p.row_base + (uint32_t)j);
            if (key > st.thr_key) {
              buf_warp[lane * kBufPitch + st.cnt] = key;
              st.cnt++;
            }
          }
          if (__any_sync(kFull, st.cnt == kBufSlots)) {
            __syncwarp();
            warp_compact(st, kBufSlots / 2, p.k, buf_warp, lists_warp, tau_warp, stats);
            thr_f = st.thr_key ? score_of_goodness<DESC>((uint32_t)(st.thr_key >> 32)) : (DESC ? -INFINITY : INFINITY);
          }
        }
      }
    }
    if (warp < 4) {
      __syncwarp();
      warp_compact(st, 1, p.k, buf_warp, lists_warp, tau_warp, stats);
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        uint32_t* flag = p.done + (size_t)item * 4 + warp;
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(1u) : "memory");
      }
    }
  }
}

// =============================================================================== merge / unpack
// One warp per query: merge `parts` sorted key lists into the best k.
//   internal layout (tile_rows > 0): list of (part, q) at src + ((part*n_qt + q/tile_rows)*tile_rows + q%tile_rows)*kListCap
//   flat layout     (tile_rows == 0): src + (part*q_rows + q)*src_len
__global__ void __launch_bounds__(256) merge_lists_kernel(const uint64_t* __restrict__ src,
                                                          const uint64_t* __restrict__ overflow, int parts, int64_t q_rows,
                                                          int n_qt, int tile_rows, int src_len, int k,
                                                          uint64_t* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t q = (int64_t)blockIdx.x * 8 + warp; q < q_rows; q += (int64_t)gridDim.x * 8) {
    uint64_t A[4] = {0, 0, 0, 0};
    for (int part = 0; part < parts; ++part) {
      const uint64_t* s = tile_rows > 0
          ? src + (((size_t)part * n_qt + (size_t)(q / tile_rows)) * tile_rows + (size_t)(q % tile_rows)) * kListCap
          : src + ((size_t)part * q_rows + (size_t)q) * src_len;
      uint64_t B[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) B[r] = (lane + 32 * r < src_len) ? __ldg(s + lane + 32 * r) : 0ull;
      if (part == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) A[r] = B[r];
      } else {
        warp_merge_lists(A, B);      // one bitonic merge per part: ~200 shuffle/compare instructions
      }
      if (overflow != nullptr) {     // keys that were still in the item's append buffer (unsorted, <= kBufSlots)
        const uint64_t* ov = overflow + (((size_t)part * n_qt + (size_t)(q / tile_rows)) * tile_rows + (size_t)(q % tile_rows)) * kBufSlots;
        const uint64_t nk = lane < kBufSlots ? __ldg(ov + lane) : 0ull;
        if (__any_sync(kFull, nk != 0)) warp_list_insert(A, nk);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (lane + 32 * r < k) out[(size_t)q * k + lane + 32 * r] = A[r];
  }
}

__global__ void __launch_bounds__(256) unpack_keys_kernel(const uint64_t* __restrict__ keys, int64_t count, int descending,
                                                          float* __restrict__ scores, int64_t* __restrict__ rows) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = keys[i];
    if (key == 0) {   // fewer than k candidates
      if (scores) scores[i] = descending ? -INFINITY : INFINITY;
      if (rows) rows[i] = -1;
      continue;
    }
    const uint32_t g = (uint32_t)(key >> 32);
    if (scores) scores[i] = descending ? score_of_goodness<true>(g) : score_of_goodness<false>(g);
    if (rows) rows[i] = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu));
  }
}

// Lower bound from a probe pass: `groups` disjoint row groups were scanned with top-kp each (lists in the internal layout);
// bound[q] = the smallest of their kp-th best keys' upper words (0 = some group holds fewer than kp rows: no bound).
// groups * kp >= k rows are >= that key, so nothing below it can be in the top-k of the whole catalog.
__global__ void __launch_bounds__(256) probe_bound_kernel(const uint64_t* __restrict__ lists, int groups, int n_qt, int tile_rows, int kp,
                                                          int64_t q_rows, uint32_t* __restrict__ bound_u32, int64_t* __restrict__ bound_i64) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < q_rows; q += (int64_t)gridDim.x * blockDim.x) {
    uint32_t b = 0xFFFFFFFFu;
    for (int g = 0; g < groups; ++g) {
      const uint64_t key = __ldg(lists + (((size_t)g * n_qt + (size_t)(q / tile_rows)) * tile_rows + (size_t)(q % tile_rows)) * kListCap + (kp - 1));
      const uint32_t w = (uint32_t)(key >> 32);
      b = w < b ? w : b;
    }
    if (bound_u32) bound_u32[q] = b;
    if (bound_i64) bound_i64[q] = (int64_t)b;
  }
}

// tau_global seeded from device-resident u32 words (the library's own probe)
__global__ void __launch_bounds__(256) seed_tau_u32_kernel(const uint32_t* __restrict__ init, uint32_t* __restrict__ tau, int64_t q_rows, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    tau[i] = i < q_rows ? init[i] : 0u;
}

// tau_global seeded from caller-provided lower bounds (goodness words held in int64), 0 = no bound
__global__ void __launch_bounds__(256) seed_tau_kernel(const int64_t* __restrict__ init, uint32_t* __restrict__ tau, int64_t q_rows,
                                                       int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    tau[i] = i < q_rows ? (uint32_t)(uint64_t)init[i] : 0u;
}

template <typename T>
__global__ void __launch_bounds__(256) inv_norm_rows_kernel(const T* x, int64_t n, int d, int64_t ldx, float eps, float* out) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += (int64_t)gridDim.x * 8) {
    const T* xr = x + row * ldx;
    float a = 0.f;
    for (int j = lane; j < d; j += 32) { const float v = to_float<T>(xr[j]); a = fmaf(v, v, a); }
    a = warp_sum(a);
    if (lane == 0) out[row] = 1.0f / fmaxf(sqrtf(a), eps);
  }
}

// =============================================================================== host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// 2-D row-major [rows, d] 16-bit matrix, box = 64 columns (128 B, swizzle 128B) x box_rows rows
int make_tmap(CUtensorMap* map, int dtype, const void* base, int64_t rows, int64_t d, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) { set_error("cuTensorMapEncodeTiled not available from the driver"); return IA_ERR_CUDA; }
  const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == IA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return IA_ERR_CUDA; }
  return IA_OK;
}

}  // namespace ia

using namespace ia;

struct ia_catalog {
  int dtype;
  const void* data;
  int64_t c, d, ld;
  uint32_t row_base;
  int device;
  float* cinv;          // [c]
  bool tc_ok;           // tensor-core path usable (16-bit dtype, TMA-compatible layout)
  CUtensorMap tmap_c;        // box = 256 catalog rows x 64 columns
  CUtensorMap tmap_c_half;   // box = 128 rows: what each CTA of a pair stages (cta_group::2 form of the kernel)
  bool pair_ok;
  // lazily grown scratch
  uint64_t* lists; size_t lists_bytes;
  uint32_t* tau; size_t tau_bytes;      // [tau | done flags]
  float* qinv; size_t qinv_bytes;
  uint32_t* bound; size_t bound_bytes;  // [q] probe bound words of the last call
  unsigned long long* stats;            // [8]
  int last_splits, last_tiles_per_split;
};

static int grow(void** ptr, size_t* have, size_t need) {
  if (*have >= need) return IA_OK;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr; *have = 0;
  IA_CUDA_CHECK(cudaMalloc(ptr, need));
  *have = need;
  return IA_OK;
}

// Choose the number of catalog splits.  Every (query tile, split) item pays a cold start -- its top-k lists fill
// and its thresholds tighten from scratch -- worth roughly `penalty` tiles of tensor-core time (it grows with k),
// so the plan minimises  waves * (tiles_per_split + penalty): fill the SMs in whole waves with as FEW splits as
// that allows.  (A plan that only maximised wave efficiency picked 40+ splits and spent its time on cold starts.)
static void plan_splits(int n_qt, int n_tiles, int ctas, int min_tiles, double penalty, int* n_splits, int* tiles_per_split) {
  int best_s = 1;
  double best_cost = 1e300;
  const int max_s = n_tiles / min_tiles > 1 ? n_tiles / min_tiles : 1;
  for (int s = 1; s <= max_s && s <= 32; ++s) {      // <= 32: the finished-split bound tracks 32 splits
    const int tps = (n_tiles + s - 1) / s;
    const int real_s = (n_tiles + tps - 1) / tps;
    const int64_t items = (int64_t)n_qt * real_s;
    const int64_t waves = (items + ctas - 1) / ctas;
    const double cost = (double)waves * ((double)tps + penalty);
    if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best_s = real_s; }
  }
  *n_splits = best_s;
  *tiles_per_split = (n_tiles + best_s - 1) / best_s;
  *n_splits = (n_tiles + *tiles_per_split - 1) / *tiles_per_split;
}

extern "C" {

int ia_catalog_create(ia_catalog** out, int dtype, const void* catalog, int64_t c, int64_t d, int64_t ld,
                      int64_t row_base, ia_stream_t stream) {
  if (out == nullptr || catalog == nullptr || c <= 0 || d <= 0 || ld < d || row_base < 0 ||
      row_base + c > 0xFFFFFFFFll) {
    set_error("bad catalog arguments (rows must fit 32-bit global ids)");
    return IA_ERR_INVALID;
  }
  if (dtype < IA_F32 || dtype > IA_F16) { set_error("unsupported dtype %d", dtype); return IA_ERR_UNSUPPORTED; }
  ia_catalog* cat = new (std::nothrow) ia_catalog();
  if (!cat) { set_error("out of host memory"); return IA_ERR_CUDA; }
  cat->dtype = dtype; cat->data = catalog; cat->c = c; cat->d = d; cat->ld = ld; cat->row_base = (uint32_t)row_base;
  cat->lists = nullptr; cat->lists_bytes = 0; cat->tau = nullptr; cat->tau_bytes = 0; cat->qinv = nullptr; cat->qinv_bytes = 0;
  cat->cinv = nullptr; cat->tc_ok = false; cat->pair_ok = false; cat->stats = nullptr; cat->bound = nullptr; cat->bound_bytes = 0;
  cudaGetDevice(&cat->device);
  cudaStream_t s = (cudaStream_t)stream;
  // inverse norms padded with zeros to whole 256-row tiles: the kernel bulk-copies one tile's worth at a time
  const size_t cinv_n = ((size_t)c + tc::BN - 1) / tc::BN * tc::BN;
  cudaError_t e = cudaMalloc(&cat->cinv, sizeof(float) * cinv_n);
  if (e == cudaSuccess) e = cudaMalloc(&cat->stats, sizeof(unsigned long long) * 8);
  if (e == cudaSuccess) e = cudaMemsetAsync(cat->cinv, 0, sizeof(float) * cinv_n, s);
  if (e != cudaSuccess) { set_error("catalog allocation failed: %s", cudaGetErrorString(e)); if (cat->cinv) cudaFree(cat->cinv); delete cat; return IA_ERR_CUDA; }
  const int64_t want = (c + 7) / 8;
  const int grid = (int)(want < 8 * sm_count() ? want : 8 * sm_count());
  if (dtype == IA_F32) inv_norm_rows_kernel<float><<<grid, 256, 0, s>>>((const float*)catalog, c, (int)d, ld, kCosEps, cat->cinv);
  else if (dtype == IA_BF16) inv_norm_rows_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)catalog, c, (int)d, ld, kCosEps, cat->cinv);
  else inv_norm_rows_kernel<__half><<<grid, 256, 0, s>>>((const __half*)catalog, c, (int)d, ld, kCosEps, cat->cinv);
  g_launches.fetch_add(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("inverse-norm launch failed: %s", cudaGetErrorString(e)); cudaFree(cat->cinv); delete cat; return IA_ERR_CUDA; }
  if (dtype != IA_F32 && d % 8 == 0 && ld % 8 == 0 && reinterpret_cast<uintptr_t>(catalog) % 16 == 0) {
    if (make_tmap(&cat->tmap_c, dtype, catalog, c, d, ld, tc::BN) == IA_OK) cat->tc_ok = true;
    cat->pair_ok = cat->tc_ok && make_tmap(&cat->tmap_c_half, dtype, catalog, c, d, ld, tc::BN / 2) == IA_OK;
  }
  *out = cat;
  return IA_OK;
}

void ia_catalog_destroy(ia_catalog* cat) {
  if (!cat) return;
  if (cat->cinv) cudaFree(cat->cinv);
  if (cat->lists) cudaFree(cat->lists);
  if (cat->tau) cudaFree(cat->tau);
  if (cat->qinv) cudaFree(cat->qinv);
  if (cat->bound) cudaFree(cat->bound);
  if (cat->stats) cudaFree(cat->stats);
  delete cat;
}

int ia_catalog_topk(ia_catalog* cat, int measure, const void* queries, int64_t q, int64_t ldq, int k,
                    uint64_t* keys_out, ia_stream_t stream) {
  return ia_catalog_topk_seeded(cat, measure, queries, q, ldq, k, nullptr, keys_out, stream);
}

// CTA-pair (cta_group::2) form of the tensor-core kernel: IA_RETR_PAIR=1 / 0 forces it on / off.  A pair shares every catalog
// tile between two query tiles, so it needs at least two of them; an odd count wastes one tile of MMA work.
static bool use_pair(const ia_catalog* cat, int n_qt) {
  const char* e = getenv("IA_RETR_PAIR");     // read per call: tests and A/B runs switch it between calls
  const int env = e ? atoi(e) : -1;
  if (!cat->pair_ok || n_qt < 2 || (sm_count() & 1)) return false;
  if (env >= 0) return env != 0;
  return kPairDefault && (n_qt % 2 == 0 || n_qt >= 32);
}

// one launch of the tensor-core kernel for the decomposition in p; the grid is one CTA per SM or as many as there is work
static int launch_tc(ia_catalog* cat, int measure, bool kreg, bool pair, const CUtensorMap& tmap_q, const RetrParams& p, cudaStream_t s) {
  const int fmt_bf16 = cat->dtype == IA_BF16;
  const int sms = sm_count();
  auto launch = [&](auto kernel, auto ctas_tag) -> int {
    constexpr int CTAS = decltype(ctas_tag)::value;
    constexpr int smem = tc::Cfg<CTAS>::SMEM_BYTES;
    IA_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t n_work = (int64_t)((p.n_qt + CTAS - 1) / CTAS) * p.n_splits;
    const int units = (int)(n_work < sms / CTAS ? n_work : sms / CTAS);
    if (CTAS == 1) {
      kernel<<<units, tc::THREADS, smem, s>>>(tmap_q, cat->tmap_c, p);
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(units * CTAS)); cfg.blockDim = dim3(tc::THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      IA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, tmap_q, cat->tmap_c_half, p));
    }
    IA_LAUNCH_CHECK();
    return IA_OK;
  };
  using One = std::integral_constant<int, 1>;
  using Two = std::integral_constant<int, 2>;
#define IA_TC_DISPATCH(COS, KR)                                                                                               \
  do {                                                                                                                        \
    if (pair) return fmt_bf16 ? launch(retrieve_tc_kernel<COS, 1, KR, 2>, Two{}) : launch(retrieve_tc_kernel<COS, 0, KR, 2>, Two{}); \
    return fmt_bf16 ? launch(retrieve_tc_kernel<COS, 1, KR, 1>, One{}) : launch(retrieve_tc_kernel<COS, 0, KR, 1>, One{});    \
  } while (0)
  if (kreg) {
    if (measure == IA_COSINE) IA_TC_DISPATCH(true, 16);
    IA_TC_DISPATCH(false, 16);
  }
  if (measure == IA_COSINE) IA_TC_DISPATCH(true, 0);
  IA_TC_DISPATCH(false, 0);
#undef IA_TC_DISPATCH
}

static int grow_scratch(ia_catalog* cat, int64_t n_items, int n_qt, int BM) {
  int rc;
  const size_t lists_n = (size_t)n_items * BM * kListCap, ov_n = (size_t)n_items * BM * kBufSlots;
  if ((rc = grow((void**)&cat->lists, &cat->lists_bytes, sizeof(uint64_t) * (lists_n + ov_n))) != IA_OK) return rc;
  const size_t tau_n = (size_t)n_qt * BM, done_n = (size_t)n_items * 4;
  return grow((void**)&cat->tau, &cat->tau_bytes, sizeof(uint32_t) * (tau_n + done_n));
}

// Probe pass (tensor-core path): `groups` disjoint row groups of `tiles_per_group` 256-row tiles at the head of the catalog are
// scanned with the register top-k (kp <= 16), every group on its own (no threshold sharing); cat->bound[q] = the smallest of
// the groups' kp-th best key words.  The caller guarantees groups * kp >= the k it will ask for.
static int probe_pass(ia_catalog* cat, int measure, const CUtensorMap& tmap_q, RetrParams p, int groups, int tiles_per_group, int kp,
                      int64_t q, int64_t* bound_i64, cudaStream_t s) {
  int rc;
  p.k = kp;
  p.n_splits = groups; p.tiles_per_split = tiles_per_group; p.n_tiles = groups * tiles_per_group;
  p.flags |= 32;
  const int64_t n_items = (int64_t)p.n_qt * groups;
  if ((rc = grow_scratch(cat, n_items, p.n_qt, tc::BM)) != IA_OK) return rc;
  if ((rc = grow((void**)&cat->bound, &cat->bound_bytes, sizeof(uint32_t) * (size_t)q)) != IA_OK) return rc;
  const size_t tau_n = (size_t)p.n_qt * tc::BM, done_n = (size_t)n_items * 4;
  IA_CUDA_CHECK(cudaMemsetAsync(cat->tau, 0, sizeof(uint32_t) * (tau_n + done_n), s));
  p.lists = cat->lists; p.overflow = nullptr; p.tau_global = cat->tau; p.done = cat->tau + tau_n; p.stats = nullptr;
  const int sms = sm_count();
  if ((rc = launch_tc(cat, measure, true, use_pair(cat, p.n_qt), tmap_q, p, s)) != IA_OK) return rc;
  const int64_t want = (q + 255) / 256;
  probe_bound_kernel<<<(int)(want < 4 * sms ? want : 4 * sms), 256, 0, s>>>(cat->lists, groups, p.n_qt, tc::BM, kp, q, cat->bound, bound_i64);
  IA_LAUNCH_CHECK();
  return IA_OK;
}

static int catalog_topk_impl(ia_catalog* cat, int measure, float dist_eps, int dist_squared, const void* queries, int64_t q,
                             int64_t ldq, int k, const int64_t* tau_init, uint64_t* keys_out, ia_stream_t stream) {
  if (cat == nullptr || queries == nullptr || keys_out == nullptr || q < 0 || ldq < cat->d) { set_error("bad arguments"); return IA_ERR_INVALID; }
  if (measure < IA_INNER || measure > IA_L2) { set_error("Unsupported similarty measure: %d", measure); return IA_ERR_INVALID; }
  if (k < 1 || k > IA_MAX_K) { set_error("k must be in [1, %d]", IA_MAX_K); return IA_ERR_INVALID; }
  if (q == 0) return IA_OK;
  if (q > 0x7FFFFFFF) { set_error("too many queries in one call"); return IA_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  const bool desc = (measure == IA_INNER || measure == IA_COSINE);
  bool use_tc = desc && cat->tc_ok && ldq % 8 == 0 && reinterpret_cast<uintptr_t>(queries) % 16 == 0;
  CUtensorMap tmap_q;
  if (use_tc && make_tmap(&tmap_q, cat->dtype, queries, q, cat->d, ldq, tc::BM) != IA_OK) use_tc = false;   // -> CUDA-core path
  const int BM = use_tc ? tc::BM : simt::BM, BN = use_tc ? tc::BN : simt::BN;

  RetrParams p{};
  p.q_rows = (int)q; p.c_rows = cat->c; p.d = (int)cat->d; p.k = k;
  p.n_qt = (int)((q + BM - 1) / BM);
  p.n_tiles = (int)((cat->c + BN - 1) / BN);
  p.kblocks = (int)((cat->d + tc::BK - 1) / tc::BK);
  p.row_base = cat->row_base;
  p.dist_eps = dist_eps; p.dist_squared = dist_squared;
  p.flags = 14;
  if (const char* f = getenv("IA_RETR_FLAGS")) p.flags = atoi(f);
  p.cinv = cat->cinv;
  int rc;
  if (measure == IA_COSINE) {
    if ((rc = grow((void**)&cat->qinv, &cat->qinv_bytes, sizeof(float) * (size_t)q)) != IA_OK) return rc;
    if ((rc = ia_row_inv_norm(cat->dtype, queries, q, cat->d, ldq, kCosEps, cat->qinv, stream)) != IA_OK) return rc;
    p.qinv = cat->qinv;
  }
  const int sms = sm_count();
  int ctas = sms;
  if (!use_tc) ctas = sms * 2;
  const bool use_kreg = use_tc && k <= 16;

  // Cold start of a k > 16 scan: every (query tile, split) fills its 128-key lists from nothing, ~k (1 + ln(rows / k)) insertions
  // per query (C4: 1 975 appends and 105 list merges per query).  A probe over 1/32 of the catalog on the register top-k path
  // (ceil(k/16) row groups, top-ceil(k/groups) each) costs ~3 % of the scan and hands every item a threshold that >= k rows meet,
  // so the main pass only collects what can still matter.  Same result by construction of the bound.  IA_RETR_PROBE=0: off.
  static const int probe_on = [] { const char* e = getenv("IA_RETR_PROBE"); return e ? atoi(e) : 1; }();
  bool seeded_by_probe = false;
  if (use_tc && !use_kreg && tau_init == nullptr && probe_on) {
    const int groups = (k + 15) / 16;
    const int kp = (k + groups - 1) / groups;
    int tpg = p.n_tiles / 32 / groups;
    if (tpg < 4) tpg = 4;
    if ((int64_t)groups * tpg * 4 <= p.n_tiles && (int64_t)groups * tpg * BN <= cat->c) {
      if ((rc = probe_pass(cat, measure, tmap_q, p, groups, tpg, kp, q, nullptr, s)) != IA_OK) return rc;
      seeded_by_probe = true;
    }
  }

  // cold-start cost per item in tiles: small when the thresholds are seeded or the list lives in registers
  double penalty = (use_tc ? 0.3 : 0.15) * k;
  if (use_tc && (tau_init != nullptr || use_kreg || seeded_by_probe)) penalty = 2.0 + 0.03 * k;
  if (const char* f = getenv("IA_RETR_PENALTY")) penalty = atof(f) * k;
  // a CTA pair is one scheduling unit that takes two query tiles through a split
  const bool pair = use_tc && use_pair(cat, p.n_qt);
  if (pair) plan_splits((p.n_qt + 1) / 2, p.n_tiles, ctas / 2, 8, penalty, &p.n_splits, &p.tiles_per_split);
  else plan_splits(p.n_qt, p.n_tiles, ctas, use_tc ? 8 : 4, penalty, &p.n_splits, &p.tiles_per_split);
  const int64_t n_items = (int64_t)p.n_qt * p.n_splits;

  if ((rc = grow_scratch(cat, n_items, p.n_qt, BM)) != IA_OK) return rc;
  const size_t lists_n = (size_t)n_items * BM * kListCap;
  const size_t tau_n = (size_t)p.n_qt * BM, done_n = (size_t)n_items * 4;
  IA_CUDA_CHECK(cudaMemsetAsync(cat->tau, 0, sizeof(uint32_t) * (tau_n + done_n), s));
  if (tau_init != nullptr) {
    seed_tau_kernel<<<(int)((tau_n + 255) / 256), 256, 0, s>>>(tau_init, cat->tau, q, (int64_t)tau_n);
    IA_LAUNCH_CHECK();
  } else if (seeded_by_probe) {
    seed_tau_u32_kernel<<<(int)((tau_n + 255) / 256), 256, 0, s>>>(cat->bound, cat->tau, q, (int64_t)tau_n);
    IA_LAUNCH_CHECK();
  }
  IA_CUDA_CHECK(cudaMemsetAsync(cat->stats, 0, sizeof(unsigned long long) * 8, s));
  cat->last_splits = p.n_splits; cat->last_tiles_per_split = p.tiles_per_split;
  p.lists = cat->lists; p.overflow = cat->lists + lists_n; p.tau_global = cat->tau; p.done = cat->tau + tau_n; p.stats = cat->stats;
  const int grid = (int)(n_items < ctas ? n_items : ctas);

  if (use_tc) {
    if ((rc = launch_tc(cat, measure, use_kreg, pair, tmap_q, p, s)) != IA_OK) return rc;
  } else {
    auto launch = [&](auto kernel, auto* tq) -> int {
      using TP = decltype(tq);
      IA_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, simt::SMEM_BYTES));
      kernel<<<grid, simt::THREADS, simt::SMEM_BYTES, s>>>((TP)queries, ldq, (TP)cat->data, cat->ld, p);
      IA_LAUNCH_CHECK();
      return IA_OK;
    };
    // 128-bit row loads when every 8-column group of every row is 16-byte aligned (32-byte for fp32 pairs is not needed)
    const size_t esz = cat->dtype == IA_F32 ? 4 : 2;
    const bool vec = cat->d % 8 == 0 && (ldq * esz) % 16 == 0 && (cat->ld * esz) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(queries) % 16 == 0 && reinterpret_cast<uintptr_t>(cat->data) % 16 == 0;
#define IA_SIMT_DISPATCH_V(T, V)                                                                       \
    switch (measure) {                                                                                 \
      case IA_INNER: rc = launch(retrieve_simt_kernel<T, IA_INNER, V>, (const T*)nullptr); break;      \
      case IA_COSINE: rc = launch(retrieve_simt_kernel<T, IA_COSINE, V>, (const T*)nullptr); break;    \
      case IA_L1: rc = launch(retrieve_simt_kernel<T, IA_L1, V>, (const T*)nullptr); break;            \
      default: rc = launch(retrieve_simt_kernel<T, IA_L2, V>, (const T*)nullptr); break;               \
    }
#define IA_SIMT_DISPATCH(T) if (vec) { IA_SIMT_DISPATCH_V(T, true) } else { IA_SIMT_DISPATCH_V(T, false) }
    if (cat->dtype == IA_F32) { IA_SIMT_DISPATCH(float) }
    else if (cat->dtype == IA_BF16) { IA_SIMT_DISPATCH(__nv_bfloat16) }
    else { IA_SIMT_DISPATCH(__half) }
#undef IA_SIMT_DISPATCH_V
#undef IA_SIMT_DISPATCH
    if (rc != IA_OK) return rc;
  }
  const int64_t mwant = (q + 7) / 8;
  const int mgrid = (int)(mwant < 4 * sms ? mwant : 4 * sms);
  merge_lists_kernel<<<mgrid, 256, 0, s>>>(cat->lists, use_kreg || !use_tc ? nullptr : p.overflow, p.n_splits, q, p.n_qt, BM, kListCap, k,
                                           keys_out);
  IA_LAUNCH_CHECK();
  return IA_OK;
}

int ia_catalog_topk_seeded(ia_catalog* cat, int measure, const void* queries, int64_t q, int64_t ldq, int k,
                           const int64_t* tau_init, uint64_t* keys_out, ia_stream_t stream) {
  return catalog_topk_impl(cat, measure, kPdistEps, 0, queries, q, ldq, k, tau_init, keys_out, stream);
}

int ia_catalog_probe_bound(ia_catalog* cat, int measure, const void* queries, int64_t q, int64_t ldq, int kp, int groups,
                           int64_t rows_per_group, int64_t* bound_out, ia_stream_t stream) {
  if (cat == nullptr || queries == nullptr || bound_out == nullptr || q < 0 || ldq < cat->d) { set_error("bad arguments"); return IA_ERR_INVALID; }
  if (measure != IA_INNER && measure != IA_COSINE) { set_error("probe: inner product / cosine only"); return IA_ERR_UNSUPPORTED; }
  if (kp < 1 || kp > 16 || groups < 1 || rows_per_group < tc::BN || rows_per_group % tc::BN != 0 ||
      (int64_t)groups * rows_per_group > cat->c) {
    set_error("probe: kp in [1, 16], rows_per_group a multiple of %d, groups * rows_per_group <= catalog rows", tc::BN);
    return IA_ERR_INVALID;
  }
  if (q == 0) return IA_OK;
  if (q > 0x7FFFFFFF) { set_error("too many queries in one call"); return IA_ERR_INVALID; }
  if (!(cat->tc_ok && ldq % 8 == 0 && reinterpret_cast<uintptr_t>(queries) % 16 == 0)) { set_error("probe: needs the tensor-core path (16-bit catalog, aligned rows)"); return IA_ERR_UNSUPPORTED; }
  CUtensorMap tmap_q;
  int rc;
  if ((rc = make_tmap(&tmap_q, cat->dtype, queries, q, cat->d, ldq, tc::BM)) != IA_OK) return rc;
  RetrParams p{};
  p.q_rows = (int)q; p.c_rows = cat->c; p.d = (int)cat->d;
  p.n_qt = (int)((q + tc::BM - 1) / tc::BM);
  p.kblocks = (int)((cat->d + tc::BK - 1) / tc::BK);
  p.row_base = cat->row_base;
  p.flags = 14;
  p.cinv = cat->cinv;
  if (measure == IA_COSINE) {
    if ((rc = grow((void**)&cat->qinv, &cat->qinv_bytes, sizeof(float) * (size_t)q)) != IA_OK) return rc;
    if ((rc = ia_row_inv_norm(cat->dtype, queries, q, cat->d, ldq, kCosEps, cat->qinv, stream)) != IA_OK) return rc;
    p.qinv = cat->qinv;
  }
  return probe_pass(cat, measure, tmap_q, p, groups, (int)(rows_per_group / tc::BN), kp, q, bound_out, (cudaStream_t)stream);
}

int ia_catalog_topk_dissimilarity(ia_catalog* cat, int p_norm, float eps, int squared, const void* queries, int64_t q,
                                  int64_t ldq, int k, uint64_t* keys_out, ia_stream_t stream) {
  if (p_norm != 1 && p_norm != 2) { set_error("dissimilarity: p must be 1 or 2"); return IA_ERR_INVALID; }
  if (!(eps >= 0.f)) { set_error("dissimilarity: eps must be >= 0"); return IA_ERR_INVALID; }
  return catalog_topk_impl(cat, p_norm == 1 ? IA_L1 : IA_L2, eps, (p_norm == 2 && squared) ? 1 : 0, queries, q, ldq, k, nullptr,
                           keys_out, stream);
}

int ia_catalog_last_plan(ia_catalog* cat, int* splits, int* tiles_per_split) {
  if (cat == nullptr) { set_error("bad arguments"); return IA_ERR_INVALID; }
  if (splits) *splits = cat->last_splits;
  if (tiles_per_split) *tiles_per_split = cat->last_tiles_per_split;
  return IA_OK;
}

int ia_catalog_last_stats(ia_catalog* cat, uint64_t* out8) {
  if (cat == nullptr || out8 == nullptr) { set_error("bad arguments"); return IA_ERR_INVALID; }
  IA_CUDA_CHECK(cudaMemcpy(out8, cat->stats, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost));   // synchronising
  return IA_OK;
}

int ia_topk_merge(const uint64_t* keys_in, int parts, int64_t q, int k, uint64_t* keys_out, ia_stream_t stream) {
  if (keys_in == nullptr || keys_out == nullptr || parts < 1 || q < 0 || k < 1 || k > IA_MAX_K) { set_error("bad arguments"); return IA_ERR_INVALID; }
  if (q == 0) return IA_OK;
  const int64_t mwant = (q + 7) / 8;
  const int mgrid = (int)(mwant < 4 * sm_count() ? mwant : 4 * sm_count());
  merge_lists_kernel<<<mgrid, 256, 0, (cudaStream_t)stream>>>(keys_in, nullptr, parts, q, 0, 0, k, k, keys_out);
  IA_LAUNCH_CHECK();
  return IA_OK;
}

int ia_unpack_keys(const uint64_t* keys, int64_t count, int descending, float* scores, int64_t* rows, ia_stream_t stream) {
  if (keys == nullptr || count < 0) { set_error("bad arguments"); return IA_ERR_INVALID; }
  if (count == 0) return IA_OK;
  const int64_t want = (count + 255) / 256;
  const int grid = (int)(want < 8 * sm_count() ? want : 8 * sm_count());
  unpack_keys_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(keys, count, descending, scores, rows);
  IA_LAUNCH_CHECK();
  return IA_OK;
}

}  // extern "C"
