// Building blocks of catalog retrieval: sm_100a PTX wrappers (mbarrier, TMA, tcgen05/TMEM) and the
// per-query streaming top-k machinery shared by the tensor-core and the CUDA-core all-pairs kernels.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace ia {

// ------------------------------------------------------------------------------------ keys
// key = (goodness(score) << 32) | (0xFFFFFFFF - global_row); larger key == better candidate, ties on the
// score go to the lower row.  goodness = orderable(score) for similarities (larger is better) and its
// complement for distances.  key 0 is the "empty" sentinel (orderable() of a finite float is never 0).
__device__ __forceinline__ uint32_t orderable_u32(float f) {
  const uint32_t u = __float_as_uint(f + 0.0f);  // -0.0 -> +0.0 so equal scores compare equal
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable_u32(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}
template <bool DESC> __device__ __forceinline__ uint32_t goodness(float s) {
  const uint32_t o = orderable_u32(s);
  return DESC ? o : ~o;
}
template <bool DESC> __device__ __forceinline__ float score_of_goodness(uint32_t g) {
  return from_orderable_u32(DESC ? g : ~g);
}
template <bool DESC> __device__ __forceinline__ uint64_t make_key(float s, uint32_t row) {
  return ((uint64_t)goodness<DESC>(s) << 32) | (uint64_t)(0xFFFFFFFFu - row);
}

constexpr unsigned kFull = 0xffffffffu;
constexpr int kListCap = 128;   // IA_MAX_K: list entries per query (4 per lane)
constexpr int kBufSlots = 24;   // append-buffer slots per query
constexpr int kBufPitch = kBufSlots + 1;  // u64 words per query in shared memory (padded)

// ---------------------------------------------------------------------------------- warp-level list merges
// A per-query list is 128 keys sorted descending, held by a warp 4 per lane in STRIPED order (register r of
// lane l is position l + 32*r; empty = 0 sinks to the end).  Merging is done with bitonic networks on
// registers + shuffles: no shared memory, no data-dependent addressing, 4 independent chains per stage.
__device__ __forceinline__ uint64_t u64max(uint64_t a, uint64_t b) { return a > b ? a : b; }
__device__ __forceinline__ uint64_t u64min(uint64_t a, uint64_t b) { return a < b ? a : b; }

// Sort a bitonic 128-sequence (striped over the warp) into descending order: strides 64, 32 live inside a
// thread, strides 16..1 are shuffles.
__device__ __forceinline__ void warp_bitonic_finish_desc(uint64_t (&A)[4]) {
  const int lane = threadIdx.x & 31;
  uint64_t hi, lo;
  hi = u64max(A[0], A[2]); lo = u64min(A[0], A[2]); A[0] = hi; A[2] = lo;      // stride 64
  hi = u64max(A[1], A[3]); lo = u64min(A[1], A[3]); A[1] = hi; A[3] = lo;
  hi = u64max(A[0], A[1]); lo = u64min(A[0], A[1]); A[0] = hi; A[1] = lo;      // stride 32
  hi = u64max(A[2], A[3]); lo = u64min(A[2], A[3]); A[2] = hi; A[3] = lo;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const bool lower = (lane & s) == 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint64_t o = __shfl_xor_sync(kFull, A[r], s);
      A[r] = lower ? u64max(A[r], o) : u64min(A[r], o);
    }
  }
}

// A <- top-128 of (A  U  B), both sorted descending, striped.  C[i] = max(A[i], B[127-i]) is bitonic and holds
// the 128 largest of the union; then finish the sort.
__device__ __forceinline__ void warp_merge_lists(uint64_t (&A)[4], const uint64_t (&B)[4]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < 4; ++r) A[r] = u64max(A[r], __shfl_sync(kFull, B[3 - r], 31 - lane));
  warp_bitonic_finish_desc(A);
}

// A <- top-128 of (A  U  {one new key per lane, unsorted, 0 = none}).
__device__ __forceinline__ void warp_list_insert(uint64_t (&A)[4], uint64_t nk) {
  const int lane = threadIdx.x & 31;
  // bitonic sort of the 32 new keys, descending by lane
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint64_t o = __shfl_xor_sync(kFull, nk, j);
      const bool desc = (k == 32) || ((lane & k) == 0);
      const bool keep_max = (((lane & j) == 0) == desc);
      nk = keep_max ? u64max(nk, o) : u64min(nk, o);
    }
  }
  // new keys occupy positions 0..31 of a (zero padded) sorted list B: B[127-p] is non-empty only for p >= 96
  A[3] = u64max(A[3], __shfl_sync(kFull, nk, 31 - lane));
  warp_bitonic_finish_desc(A);
}

__device__ __forceinline__ void warp_load_list(uint64_t (&Lr)[4], const uint64_t* list) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < 4; ++r) Lr[r] = list[lane + 32 * r];
}
__device__ __forceinline__ void warp_store_list(uint64_t* list, const uint64_t (&Lr)[4]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < 4; ++r) list[lane + 32 * r] = Lr[r];
}
// key at position pos (0..127) of a striped list, broadcast to all lanes
__device__ __forceinline__ uint64_t warp_list_at(const uint64_t (&Lr)[4], int pos) {
  const int r = pos >> 5;
  const uint64_t v = r == 0 ? Lr[0] : (r == 1 ? Lr[1] : (r == 2 ? Lr[2] : Lr[3]));
  return __shfl_sync(kFull, v, pos & 31);
}

// Streaming top-k state of ONE query, owned by one thread; 32 queries (one warp) are compacted together.
//   thr_key : candidates must beat this key (max of the local k-th best key and the bounds other CTAs
//             published for this query)
//   cnt     : filled slots of this query's append buffer (shared memory, kBufPitch u64 per query)
struct TopKThread {
  uint64_t thr_key;
  int cnt;
};

struct TopKStats {
  unsigned appends, compactions, rare_groups, rare_blocks;
};

// Compact the append buffers of every lane whose buffer holds >= min_cnt keys into its global list.
// buf_warp: this warp's 32 append buffers; lists_warp: this warp's 32 lists (kListCap u64 each).
// Must be called by all 32 lanes.  tau_global may be null.
// One copy in the binary (noinline): it is called from several places of hot loops whose code must stay
// inside the instruction cache.  The next lane's list is prefetched from L2 while the current one is merged.
// State travels by value (registers): references would force the caller's hot state through local memory.
struct CompactResult {
  uint64_t thr_key;
  int cnt;
  unsigned merges;
};
// One copy in the binary (noinline): it is called from several places of hot loops whose code must stay inside
// the instruction cache.  Fullest buffers first, at most max_lanes of them per call (bounded work keeps the
// epilogue's pace even); the next lane's list is prefetched from L2 while the current one is merged.
__device__ __noinline__ CompactResult warp_compact_impl(uint64_t thr_key, int cnt, int min_cnt, int k, uint64_t* buf_warp,
                                                        uint64_t* lists_warp, uint32_t* tau_global_warp, int max_lanes) {
  TopKThread st{thr_key, cnt};
  unsigned merges = 0;
  const int lane = threadIdx.x & 31;
  int mx = __reduce_max_sync(kFull, st.cnt);
  if (mx < min_cnt || mx == 0) return CompactResult{st.thr_key, st.cnt, 0u};
  int ql = __ffs(__ballot_sync(kFull, st.cnt == mx)) - 1;
  uint64_t Lnext[4];
  warp_load_list(Lnext, lists_warp + (size_t)ql * kListCap);
  while (true) {
    const int c = mx;
    uint64_t* list = lists_warp + (size_t)ql * kListCap;
    const uint64_t bk = lane < c ? buf_warp[ql * kBufPitch + lane] : 0ull;
    uint64_t Lr[4] = {Lnext[0], Lnext[1], Lnext[2], Lnext[3]};
    if (lane == ql) st.cnt = 0;
    ++merges;
    mx = __reduce_max_sync(kFull, st.cnt);
    const bool more = (int)merges < max_lanes && mx >= min_cnt && mx > 0;
    const int qn = more ? __ffs(__ballot_sync(kFull, st.cnt == mx)) - 1 : 0;
    if (more) warp_load_list(Lnext, lists_warp + (size_t)qn * kListCap);
    warp_list_insert(Lr, bk);
    warp_store_list(list, Lr);
    const uint64_t kth = warp_list_at(Lr, k - 1);
    if (lane == ql) {
      if (kth > st.thr_key) st.thr_key = kth;
      if (tau_global_warp != nullptr && kth != 0) atomicMax(tau_global_warp + ql, (uint32_t)(kth >> 32));
    }
    if (!more) break;
    ql = qn;
  }
  __syncwarp();
  return CompactResult{st.thr_key, st.cnt, merges};
}
__device__ __forceinline__ void warp_compact(TopKThread& st, int min_cnt, int k, uint64_t* buf_warp, uint64_t* lists_warp,
                                             uint32_t* tau_global_warp, TopKStats& stats, int max_lanes = 32) {
  const CompactResult r = warp_compact_impl(st.thr_key, st.cnt, min_cnt, k, buf_warp, lists_warp, tau_global_warp, max_lanes);
  st.thr_key = r.thr_key;
  st.cnt = r.cnt;
  stats.compactions += r.merges;
}

}  // namespace ia
