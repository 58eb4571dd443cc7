// Best-F1 threshold search on the device: reference finetune_bert.py:72-106
//
//   rows = sorted(zip(scores, labels), key=score, reverse=high_score_more_similar)        (Python's sort: STABLE)
//   for i in range(len(rows) - 1):  nextract = i + 1; ncorrect += (label == 1) ...
//       precision = ncorrect / nextract; recall = ncorrect / total; f1 = 2 * precision * recall / (precision + recall)
//       if f1 > best_f1: best = (f1, precision, recall, acc); threshold = (rows[i].score + rows[i + 1].score) / 2
//
// as hand-written kernels: an LSD radix sort (8-bit digits, stable by construction, so equal scores keep their input order
// exactly like Python's sort) of (orderable score key, label bit), a two-level prefix count of the positives, the F1 of
// every cut point in float64 with the same IEEE operations in the same order as the Python loop, and a first-maximum
// reduction (the loop's strict '>' keeps the earliest cut).  No float atomics anywhere: bit-reproducible.
//
// HBM-bound integer/byte work: a pass moves key + label of every element once in, once out (fp32 scores: 4 passes x 10 B,
// fp64: 8 passes x 18 B).  Tiles of 8192 elements per CTA; ranking inside a tile uses warp match + per-warp digit counts.
#include "common.cuh"

namespace ia {

namespace f1 {
constexpr int THREADS = 256, WARPS = 8, ITEMS = 32, TILE = THREADS * ITEMS;   // elements per CTA tile
constexpr int RADIX = 256;
}  // namespace f1

__device__ __forceinline__ uint32_t orderable_key(float f) {
  const uint32_t u = __float_as_uint(f + 0.0f);          // -0.0 -> +0.0: equal scores compare equal
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t orderable_key(double f) {
  const uint64_t u = (uint64_t)__double_as_longlong(f + 0.0);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_score(uint32_t k) { return (double)__uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); }
__device__ __forceinline__ double key_score(uint64_t k) {
  return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k));
}

// keys ascending == rows in the reference's order: descending scores (high_score_more_similar) use the complemented key
template <typename S, typename K>
__global__ void __launch_bounds__(256) f1_prepare_kernel(const S* __restrict__ scores, const int64_t* __restrict__ labels, int64_t n,
                                                         int descending, K* __restrict__ keys, uint8_t* __restrict__ lab,
                                                         unsigned long long* __restrict__ total_pos) {
  unsigned pos = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const K k = orderable_key(scores[i]);
    keys[i] = descending ? (K)~k : k;
    const uint8_t l = labels[i] == 1;
    lab[i] = l;
    pos += l;
  }
  pos = __reduce_add_sync(0xffffffffu, pos);
  if ((threadIdx.x & 31) == 0 && pos) atomicAdd(total_pos, (unsigned long long)pos);   // integer: order-independent
}

// per-tile digit histogram -> hist[digit][tile]
template <typename K>
__global__ void __launch_bounds__(f1::THREADS) radix_hist_kernel(const K* __restrict__ keys, int64_t n, int shift, int n_tiles,
                                                                  uint32_t* __restrict__ hist) {
  using namespace f1;
  __shared__ uint32_t h[RADIX];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * TILE;
#pragma unroll 4
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + j * THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// one CTA per digit: exclusive scan of the digit's row over the tiles (in place) + the digit's total
__global__ void __launch_bounds__(1024) radix_row_scan_kernel(uint32_t* __restrict__ hist, int n_tiles, uint32_t* __restrict__ digit_total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  uint32_t* row = hist + (size_t)blockIdx.x * n_tiles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < n_tiles ? row[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t s = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
      warp_sums[lane] = s;
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    const uint32_t excl = carry + (warp > 0 ? warp_sums[warp - 1] : 0u) + (x - v);
    if (i < n_tiles) row[i] = excl;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) digit_total[blockIdx.x] = carry_s;
}

// stable scatter of one tile: position = digit_base[d] + (elements with digit d in earlier tiles) + (earlier ones in this tile)
template <typename K>
__global__ void __launch_bounds__(f1::THREADS) radix_scatter_kernel(const K* __restrict__ keys_in, const uint8_t* __restrict__ lab_in,
                                                                     int64_t n, int shift, int n_tiles, const uint32_t* __restrict__ hist,
                                                                     const uint32_t* __restrict__ digit_total, K* __restrict__ keys_out,
                                                                     uint8_t* __restrict__ lab_out) {
  using namespace f1;
  __shared__ uint32_t base_s[RADIX];            // next free global slot of every digit for this tile
  __shared__ uint32_t warp_cnt[WARPS][RADIX];   // digit counts of the current chunk per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    // exclusive scan of the 256 digit totals (one per thread) + this tile's offset inside the digit
    const uint32_t v = digit_total[threadIdx.x];
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    __shared__ uint32_t ws[WARPS];
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    uint32_t off = 0;
    for (int w2 = 0; w2 < warp; ++w2) off += ws[w2];
    base_s[threadIdx.x] = off + (x - v) + hist[(size_t)threadIdx.x * n_tiles + blockIdx.x];
  }
  const int64_t tile0 = (int64_t)blockIdx.x * TILE;
  for (int j = 0; j < ITEMS; ++j) {              // chunks of 256 consecutive elements, in order
    for (int d = threadIdx.x; d < WARPS * RADIX; d += THREADS) (&warp_cnt[0][0])[d] = 0;
    __syncthreads();
    const int64_t i = tile0 + (int64_t)j * THREADS + threadIdx.x;
    const bool ok = i < n;
    K key = 0;
    uint8_t l = 0;
    uint32_t d = 0;
    if (ok) { key = keys_in[i]; l = lab_in[i]; d = (uint32_t)(key >> shift) & 255u; }
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    unsigned peers = 0;
    if (ok) peers = __match_any_sync(act, d);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (ok && rank == 0) warp_cnt[warp][d] = __popc(peers);       // the first lane of every digit group
    __syncthreads();
    if (ok) {
      uint32_t before = 0;
      for (int w2 = 0; w2 < warp; ++w2) before += warp_cnt[w2][d];
      const uint32_t pos = base_s[d] + before + rank;
      keys_out[pos] = key;
      lab_out[pos] = l;
    }
    __syncthreads();
    {
      uint32_t s = 0;
#pragma unroll
      for (int w2 = 0; w2 < WARPS; ++w2) s += warp_cnt[w2][threadIdx.x];
      base_s[threadIdx.x] += s;
    }
    __syncthreads();
  }
}

// positives per tile of the sorted labels
__global__ void __launch_bounds__(f1::THREADS) f1_tile_count_kernel(const uint8_t* __restrict__ lab, int64_t n, uint32_t* __restrict__ tile_pos) {
  using namespace f1;
  __shared__ uint32_t ws[WARPS];
  const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
  uint32_t c = 0;
#pragma unroll 4
  for (int j = 0; j < ITEMS; ++j) c += (base + j < n) ? lab[base + j] : 0u;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
    for (int w = 0; w < WARPS; ++w) s += ws[w];
    tile_pos[blockIdx.x] = s;
  }
}

struct F1Best {
  double f1;
  unsigned long long idx;
  unsigned long long ncorrect;   // positives among rows 0..idx
};
__device__ __forceinline__ bool f1_better(double fa, unsigned long long ia, double fb, unsigned long long ib) {
  return fa > fb || (fa == fb && ia < ib);      // the loop's strict '>' keeps the EARLIEST maximum
}

// F1 of every cut point i in [0, n-2] of this tile (thread t owns ITEMS consecutive rows) -> the tile's best (f1, i).
// tile_pos has been scanned exclusively by radix_row_scan_kernel (one row).
__global__ void __launch_bounds__(f1::THREADS) f1_eval_kernel(const uint8_t* __restrict__ lab, int64_t n, const uint32_t* __restrict__ tile_excl,
                                                              const unsigned long long* __restrict__ total_pos, F1Best* __restrict__ tile_best) {
  using namespace f1;
  __shared__ uint32_t ws[WARPS];
  __shared__ double bf[WARPS];
  __shared__ unsigned long long bi[WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
  uint8_t l[ITEMS];
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) { l[j] = (base + j < n) ? lab[base + j] : 0; c += l[j]; }
  // exclusive count of the positives before this thread's rows
  uint32_t x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  uint32_t before = tile_excl[blockIdx.x] + (x - c);
  for (int w2 = 0; w2 < warp; ++w2) before += ws[w2];
  const double total = (double)*total_pos;
  double best_f = 0.0;
  unsigned long long best_i = ~0ull, best_n = 0;
  unsigned long long ncorrect = before;
#pragma unroll 4
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + j;
    ncorrect += l[j];
    if (i < n - 1 && ncorrect > 0) {
      const double precision = (double)ncorrect / (double)(i + 1);
      const double recall = (double)ncorrect / total;
      const double f = 2.0 * precision * recall / (precision + recall);
      if (f > best_f) { best_f = f; best_i = (unsigned long long)i; best_n = ncorrect; }   // strict: earliest maximum of this thread
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double of = __shfl_xor_sync(0xffffffffu, best_f, o);
    const unsigned long long oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    const unsigned long long on = __shfl_xor_sync(0xffffffffu, best_n, o);
    if (f1_better(of, oi, best_f, best_i)) { best_f = of; best_i = oi; best_n = on; }
  }
  __shared__ unsigned long long bn[WARPS];
  if (lane == 0) { bf[warp] = best_f; bi[warp] = best_i; bn[warp] = best_n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < WARPS; ++w)
      if (f1_better(bf[w], bi[w], best_f, best_i)) { best_f = bf[w]; best_i = bi[w]; best_n = bn[w]; }
    tile_best[blockIdx.x] = F1Best{best_f, best_i, best_n};
  }
}

// out5 = (acc, f1, precision, recall, threshold) of the best cut; all zero when no cut has f1 > 0 (or n < 2)
template <typename K>
__global__ void __launch_bounds__(1024) f1_final_kernel(const F1Best* __restrict__ tile_best, int n_tiles, const K* __restrict__ keys,
                                                         int64_t n, int descending, const unsigned long long* __restrict__ total_pos,
                                                         double* __restrict__ out5) {
  __shared__ double bf[32];
  __shared__ unsigned long long bi[32], bn[32];
  double best_f = 0.0;
  unsigned long long best_i = ~0ull, best_n = 0;
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
    const F1Best b = tile_best[t];
    if (f1_better(b.f1, b.idx, best_f, best_i)) { best_f = b.f1; best_i = b.idx; best_n = b.ncorrect; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double of = __shfl_xor_sync(0xffffffffu, best_f, o);
    const unsigned long long oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    const unsigned long long on = __shfl_xor_sync(0xffffffffu, best_n, o);
    if (f1_better(of, oi, best_f, best_i)) { best_f = of; best_i = oi; best_n = on; }
  }
  if (lane == 0) { bf[warp] = best_f; bi[warp] = best_i; bn[warp] = best_n; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
    if (f1_better(bf[w], bi[w], best_f, best_i)) { best_f = bf[w]; best_i = bi[w]; best_n = bn[w]; }
  if (!(best_f > 0.0) || best_i == ~0ull) {
    for (int k = 0; k < 5; ++k) out5[k] = 0.0;
    return;
  }
  const int64_t i = (int64_t)best_i;
  const unsigned long long ncorrect = best_n;
  const double total = (double)*total_pos;
  const unsigned long long nextract = (unsigned long long)i + 1, fneg = nextract - ncorrect;
  const double neg_total = (double)n - total;
  const double precision = (double)ncorrect / (double)nextract;
  const double recall = (double)ncorrect / total;
  K k0 = keys[i], k1 = keys[i + 1];
  if (descending) { k0 = (K)~k0; k1 = (K)~k1; }
  out5[0] = ((double)ncorrect + neg_total - (double)fneg) / (double)n;
  out5[1] = 2.0 * precision * recall / (precision + recall);
  out5[2] = precision;
  out5[3] = recall;
  out5[4] = (key_score(k0) + key_score(k1)) / 2.0;
}

struct F1Plan {
  int n_tiles;
  size_t key_bytes, off_keys_b, off_lab_a, off_lab_b, off_hist, off_dtot, off_tpos, off_best, off_total, bytes;
};
static F1Plan f1_plan(int64_t n, size_t key_size) {
  F1Plan p;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  p.n_tiles = (int)((n + f1::TILE - 1) / f1::TILE);
  if (p.n_tiles < 1) p.n_tiles = 1;
  p.key_bytes = al((size_t)n * key_size);
  size_t o = p.key_bytes;             // keys A at 0
  p.off_keys_b = o; o += p.key_bytes;
  p.off_lab_a = o; o += al((size_t)n);
  p.off_lab_b = o; o += al((size_t)n);
  p.off_hist = o; o += al(sizeof(uint32_t) * f1::RADIX * (size_t)p.n_tiles);
  p.off_dtot = o; o += al(sizeof(uint32_t) * f1::RADIX);
  p.off_tpos = o; o += al(sizeof(uint32_t) * (size_t)p.n_tiles);
  p.off_best = o; o += al(sizeof(F1Best) * (size_t)p.n_tiles);
  p.off_total = o; o += 256;
  p.bytes = o;
  return p;
}

template <typename S, typename K>
static int best_f1_impl(const S* scores, const int64_t* labels, int64_t n, int descending, double* out5, char* ws, const F1Plan& pl,
                        cudaStream_t st) {
  using namespace f1;
  K* keys_a = reinterpret_cast<K*>(ws);
  K* keys_b = reinterpret_cast<K*>(ws + pl.off_keys_b);
  uint8_t* lab_a = reinterpret_cast<uint8_t*>(ws + pl.off_lab_a);
  uint8_t* lab_b = reinterpret_cast<uint8_t*>(ws + pl.off_lab_b);
  uint32_t* hist = reinterpret_cast<uint32_t*>(ws + pl.off_hist);
  uint32_t* dtot = reinterpret_cast<uint32_t*>(ws + pl.off_dtot);
  uint32_t* tpos = reinterpret_cast<uint32_t*>(ws + pl.off_tpos);
  F1Best* best = reinterpret_cast<F1Best*>(ws + pl.off_best);
  unsigned long long* total = reinterpret_cast<unsigned long long*>(ws + pl.off_total);
  IA_CUDA_CHECK(cudaMemsetAsync(total, 0, sizeof(unsigned long long), st));
  const int64_t want = (n + 255) / 256;
  const int pgrid = (int)(want < 8 * sm_count() ? want : 8 * sm_count());
  f1_prepare_kernel<S, K><<<pgrid, 256, 0, st>>>(scores, labels, n, descending, keys_a, lab_a, total);
  IA_LAUNCH_CHECK();
  for (int pass = 0; pass < (int)sizeof(K); ++pass) {
    const int shift = 8 * pass;
    radix_hist_kernel<K><<<pl.n_tiles, THREADS, 0, st>>>(keys_a, n, shift, pl.n_tiles, hist);
    IA_LAUNCH_CHECK();
    radix_row_scan_kernel<<<RADIX, 1024, 0, st>>>(hist, pl.n_tiles, dtot);
    IA_LAUNCH_CHECK();
    radix_scatter_kernel<K><<<pl.n_tiles, THREADS, 0, st>>>(keys_a, lab_a, n, shift, pl.n_tiles, hist, dtot, keys_b, lab_b);
    IA_LAUNCH_CHECK();
    K* tk = keys_a; keys_a = keys_b; keys_b = tk;
    uint8_t* tl = lab_a; lab_a = lab_b; lab_b = tl;
  }
  f1_tile_count_kernel<<<pl.n_tiles, THREADS, 0, st>>>(lab_a, n, tpos);
  IA_LAUNCH_CHECK();
  radix_row_scan_kernel<<<1, 1024, 0, st>>>(tpos, pl.n_tiles, dtot);      // exclusive scan of the tile counts (one row)
  IA_LAUNCH_CHECK();
  f1_eval_kernel<<<pl.n_tiles, THREADS, 0, st>>>(lab_a, n, tpos, total, best);
  IA_LAUNCH_CHECK();
  f1_final_kernel<K><<<1, 1024, 0, st>>>(best, pl.n_tiles, keys_a, n, descending, total, out5);
  IA_LAUNCH_CHECK();
  return IA_OK;
}

}  // namespace ia

using namespace ia;

extern "C" {

size_t ia_best_f1_workspace_bytes(int64_t n, int score_dtype) {
  if (n < 0) return 0;
  return f1_plan(n, score_dtype == IA_F64 ? 8 : 4).bytes;
}

int ia_best_f1_threshold(int score_dtype, const void* scores, const int64_t* labels, int64_t n, int high_score_more_similar,
                         double* out5, void* workspace, size_t workspace_bytes, ia_stream_t stream) {
  if (score_dtype != IA_F32 && score_dtype != IA_F64) { set_error("best-F1 search: scores must be fp32 or fp64"); return IA_ERR_UNSUPPORTED; }
  if (n < 0 || out5 == nullptr || (n > 0 && (scores == nullptr || labels == nullptr))) { set_error("bad arguments"); return IA_ERR_INVALID; }
  if (n >= (1ll << 31)) { set_error("best-F1 search: n must be below 2^31"); return IA_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  if (n < 2) {   // the reference loop body never runs: (0, 0, 0, 0, 0)
    IA_CUDA_CHECK(cudaMemsetAsync(out5, 0, 5 * sizeof(double), st));
    return IA_OK;
  }
  const F1Plan pl = f1_plan(n, score_dtype == IA_F64 ? 8 : 4);
  if (workspace == nullptr || workspace_bytes < pl.bytes || (reinterpret_cast<uintptr_t>(workspace) & 255)) {
    set_error("best-F1 search: workspace of %zu bytes (256-byte aligned) required", pl.bytes);
    return IA_ERR_WORKSPACE;
  }
  char* ws = static_cast<char*>(workspace);
  if (score_dtype == IA_F64)
    return best_f1_impl<double, uint64_t>(static_cast<const double*>(scores), labels, n, high_score_more_similar != 0, out5, ws, pl, st);
  return best_f1_impl<float, uint32_t>(static_cast<const float*>(scores), labels, n, high_score_more_similar != 0, out5, ws, pl, st);
}

}  // extern "C"
