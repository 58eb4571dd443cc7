// Building blocks of catalog retrieval: sm_100a PTX wrappers (mbarrier, TMA, tcgen05/TMEM) and the
// per-query streaming top-k machinery shared by the tensor-core and the CUDA-core all-pairs kernels.
#pragma once
#include "common.cuh"

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

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (an error the host sees), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 200000000u) {
      printf("ia_b200: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// TMA: 2-D tiled bulk tensor load global -> shared, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 1-D bulk copy global -> shared with mbarrier completion (SASS: UBLKCP); 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// TMEM allocation (one warp), 512 columns = the whole tensor memory of the SM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]^T, bf16/fp16 inputs, fp32 accumulate (SASS: UTCHMMA)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread (SASS: LDTM)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld_1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major operand tile staged by TMA with 128-byte swizzle:
// rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO), version 1 (Blackwell), layout SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;                   // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version, bits [46,48)
  d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
// Instruction descriptor: fp32 accumulate, A and B K-major, dense; ab_format 0 = fp16, 1 = bf16.
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n, int ab_format) {
  return (1u << 4) | ((uint32_t)ab_format << 7) | ((uint32_t)ab_format << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

}  // namespace ia
