// Pair scoring kernels: one warp owns one pair (row of x, row of y).
//
// HBM-bound design: each pair's x and y rows are read ONCE with 128-bit streaming loads
// (all of a row's loads are issued before the first use so every lane keeps 2*VPL requests in flight),
// kept in registers as raw vectors, reduced with warp shuffles (sum xy, sum x^2, sum y^2, sum |d| / d^2),
// the per-pair scalar math (score -> probability -> loss -> dL/ds) runs redundantly in every lane, and
// dx, dy are produced from the same registers and written once with 128-bit stores.
// Algorithmic traffic: forward 2*D*e + 8 B/pair, fused forward+backward 4*D*e + 16 B/pair.
#pragma once
#ifndef IA_LOAD_SHAPE
#define IA_LOAD_SHAPE 1   // 1: loads pinned in address order (see the load loop); 0: compiler's order (A/B builds)
#endif
#include <cstdlib>

#include "common.cuh"
#include "ptx_sm100.cuh"   // mbarrier / bulk-copy PTX wrappers

namespace ia {

enum PairMode { kModeFwd = 0, kModeFused = 1, kModeBwd = 2 };

struct PairParams {
  const void* x;
  const void* y;
  int64_t ldx, ldy;            // elements
  const int64_t* xi;           // optional gather: pair r reads row xi[r] of x (and yi[r] of y); null = row r
  const int64_t* yi;
  const int64_t* labels;       // fused
  const float* gsim;           // bwd
  int64_t n;
  int d;
  float* sim;
  float* probs;
  uint8_t* labels_out;         // fwd
  double threshold;            // fwd
  float* loss_out;             // fused
  void* dx;
  void* dy;
  int64_t lddx, lddy;
  int loss;
  float margin;
  int reduction;
  float grad_scale;            // upstream * (1/n for mean)
  double loss_scale;           // 1/n for mean, 1 otherwise
  void* workspace;
  int load_mode;               // bit 1: ld.global.cs instead of the no-allocate loads (0 in production; see the load loop)
  int64_t rows_x, rows_y;      // gather: rows of the embedding matrices; an index outside [0, rows) poisons that pair with NaN
  const float* upstream;       // fused: optional DEVICE scalar d(total)/d(loss) folded into the gradients (autograd backward)
  int upstream_skip_one;       // with upstream: leave at once when *upstream == 1 (the gradients already in dx, dy are exact)
  float act_bwd;               // 0: dx, dy are gradients w.r.t. x, y.  > 0: x, y are tanh outputs after nn.Dropout with keep scale
                               // act_bwd = 1/(1-p) (1 = no dropout) and the kernel writes d_pre = g * keep * scale * (1 - t^2),
                               // the gradient w.r.t. the dense layer's output (head projection backward, base.py:67-75)
};

// Per-pair sums gathered in one sweep over the registers.
struct RowSums {
  float xy, xx, yy, dist;  // dist = sum |d| (l1) or sum d^2 (l2), d = x - y + eps
};

template <int MEASURE, bool NEED_COS>
__device__ __forceinline__ void accumulate(const float* fx, const float* fy, int n, RowSums& s) {
#pragma unroll
  for (int j = 0; j < n; ++j) {
    if (MEASURE == IA_INNER || MEASURE == IA_COSINE || NEED_COS) s.xy = fmaf(fx[j], fy[j], s.xy);
    if (MEASURE == IA_COSINE || NEED_COS) {
      s.xx = fmaf(fx[j], fx[j], s.xx);
      s.yy = fmaf(fy[j], fy[j], s.yy);
    }
    if (MEASURE == IA_L1) s.dist += fabsf(fx[j] - fy[j] + kPdistEps);
    if (MEASURE == IA_L2) {
      const float dd = fx[j] - fy[j] + kPdistEps;
      s.dist = fmaf(dd, dd, s.dist);
    }
  }
}

// Coefficients of the gradient:   inner / cosine / cosine-embedding:  dx = A*y - Bx*x,  dy = A*x - By*y
//                                  l1: dx = A*sign(d), dy = -dx ;  l2: dx = A*d, dy = -dx
struct GradCoef {
  float A, Bx, By;
};

template <int MEASURE>
__device__ __forceinline__ float score_from_sums(const RowSums& s, float& nx, float& ny) {
  if (MEASURE == IA_INNER) return s.xy;
  if (MEASURE == IA_COSINE) {
    // nn.CosineSimilarity (base.py:58): each side divided by max(norm, eps)
    nx = fmaxf(sqrtf(s.xx), kCosEps);
    ny = fmaxf(sqrtf(s.yy), kCosEps);
    return s.xy / (nx * ny);
  }
  if (MEASURE == IA_L1) return s.dist;
  return sqrtf(s.dist);
}

template <int MEASURE>
__device__ __forceinline__ GradCoef score_grad_coef(const RowSums& s, float sim, float nx, float ny, float g) {
  GradCoef c{0.f, 0.f, 0.f};
  if (MEASURE == IA_INNER) {
    c.A = g;
  } else if (MEASURE == IA_COSINE) {
    // ATen: ds/dx = (yh - s * x/|x|) / max(|x|,eps), with x/|x| := 0 at |x| = 0
    const float rx = sqrtf(s.xx), ry = sqrtf(s.yy);
    c.A = g / (nx * ny);
    c.Bx = rx > 0.f ? g * sim / (nx * rx) : 0.f;
    c.By = ry > 0.f ? g * sim / (ny * ry) : 0.f;
  } else if (MEASURE == IA_L1) {
    c.A = g;
  } else {
    c.A = sim > 0.f ? g / sim : 0.f;
  }
  return c;
}

// BULK: rows arrive through a per-warp shared-memory ring filled by cp.async.bulk (TMA engine, one instruction
// per row, completion on an mbarrier): kBulkStages rows per warp are in flight at all times, independent of the
// warp's compute / store phase and of its register budget.  Otherwise: plain 128-bit streaming loads.
constexpr int kBulkStages = 3;

// ROWS: pairs a warp processes per iteration (their loads are all issued up front).  2 KB rows (16-bit, D=1024)
// use ROWS=2 so that a warp has 8 KB in flight per iteration like a 4 KB fp32 row does.
// ACT: the backward of the head's tanh (+ dropout) is folded into the gradient store (PairParams::act_bwd).  A template
// parameter, not a run-time test: the test alone cost the benchmarked instantiation 2 % (93.4 vs 91.3 us at config 2).
template <typename T, typename G, int MEASURE, int MODE, bool COSLOSS, int VPL, bool BULK, int ROWS, bool ACT = false>
__global__ void __launch_bounds__(256) pair_kernel(const PairParams p) {
  static_assert(!BULK || ROWS == 1, "the bulk ring feeds one row per iteration");
  pdl_wait();       // launched with programmatic stream serialization (launch_pdl): nothing is read before the predecessor is done
  pdl_trigger();
  constexpr int E = VecTraits<T>::kElems;
  constexpr bool kGradCosForm = COSLOSS || MEASURE == IA_INNER || MEASURE == IA_COSINE;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int nvec = p.d / E;
  float loss_acc = 0.f;
  float gscale = p.grad_scale;
  if (MODE == kModeFused && p.upstream != nullptr) {
    // gradient recomputation of the autograd backward: the upstream scalar (GradScaler's 65536, 1/accum_steps, ...) is folded
    // in BEFORE the one rounding to the gradient type, like the reference's fp32 backward does (finetune_text.py:479-482)
    const float u = __ldg(p.upstream);
    if (p.upstream_skip_one && u == 1.0f) return;
    gscale *= u;
  }

  extern __shared__ __align__(128) uint8_t ring_raw[];
  const uint32_t row_bytes = (uint32_t)p.d * (uint32_t)sizeof(T);
  uint8_t* ring = ring_raw + (size_t)warp_in_block * kBulkStages * 2 * row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_raw + (size_t)8 * kBulkStages * 2 * row_bytes) + warp_in_block * kBulkStages;
  const int64_t row0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_in_block;
  auto arm = [&](int stage, int64_t row) {   // lane 0 only
    uint8_t* dst = ring + (size_t)stage * 2 * row_bytes;
    const int64_t rx = row, ry = row;   // the bulk ring is never used for gathered rows (launch_pair_vpl)
    mbar_arrive_expect_tx(&bars[stage], 2 * row_bytes);
    bulk_load_1d(dst, static_cast<const T*>(p.x) + rx * p.ldx, row_bytes, &bars[stage]);
    bulk_load_1d(dst + row_bytes, static_cast<const T*>(p.y) + ry * p.ldy, row_bytes, &bars[stage]);
  };
  if (BULK) {
    if (lane == 0) {
      for (int s = 0; s < kBulkStages; ++s) mbar_init(&bars[s], 1);
      fence_mbar_init();
      for (int s = 0; s < kBulkStages; ++s)
        if (row0 + s * warps_total < p.n) arm(s, row0 + s * warps_total);
    }
    __syncwarp();
  }

  int it = 0;
  // a warp owns ROWS ADJACENT pairs per iteration (their rows are contiguous in memory when ld == d)
  const int64_t n_super = (p.n + ROWS - 1) / ROWS;
  for (int64_t rbase = row0; rbase < n_super; rbase += warps_total, ++it) {
    uint4 xv[ROWS][VPL], yv[ROWS][VPL];
    bool live[ROWS], bad[ROWS];
    int64_t rows[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      rows[k] = rbase * ROWS + k;
      live[k] = rows[k] < p.n;
      bad[k] = false;
    }
    if (BULK) {
      const int stage = it % kBulkStages;
      mbar_wait(&bars[stage], (uint32_t)(it / kBulkStages) & 1u);
      const uint4* xs = reinterpret_cast<const uint4*>(ring + (size_t)stage * 2 * row_bytes);
      const uint4* ys = reinterpret_cast<const uint4*>(ring + (size_t)stage * 2 * row_bytes + row_bytes);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) { xv[0][i] = xs[v]; yv[0][i] = ys[v]; }
        else { xv[0][i] = make_uint4(0, 0, 0, 0); yv[0][i] = make_uint4(0, 0, 0, 0); }
      }
      __syncwarp();                       // every lane has its vectors in registers: the slot can be refilled
      if (lane == 0) {
        const int64_t next = rbase + (int64_t)kBulkStages * warps_total;
        if (next < p.n) { fence_proxy_async(); arm(stage, next); }
      }
    } else {
#pragma unroll
      for (int k = 0; k < ROWS; ++k) {
        int64_t rx = rows[k], ry = rows[k];
        if (p.xi != nullptr && live[k]) {
          // gathered pair: an index outside the matrix reads row 0 instead and the pair's results become NaN (memory-safe;
          // torch indexing in the reference loop, graph.py:87-117, would raise)
          rx = __ldg(p.xi + rows[k]);
          ry = __ldg(p.yi + rows[k]);
          if ((uint64_t)rx >= (uint64_t)p.rows_x) { rx = 0; bad[k] = true; }
          if ((uint64_t)ry >= (uint64_t)p.rows_y) { ry = 0; bad[k] = true; }
        }
        const uint4* xr = reinterpret_cast<const uint4*>(static_cast<const T*>(p.x) + rx * p.ldx);
        const uint4* yr = reinterpret_cast<const uint4*>(static_cast<const T*>(p.y) + ry * p.ldy);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          if (live[k] && v < nvec) {
#if IA_LOAD_SHAPE == 1
            // The loads of a row group must ISSUE in address order (x0 y0 x1 y1 ... row by row): left to itself ptxas
            // shuffles the independent loads of one basic block (e.g. +0x200, +0x200', +0, +0x400, ...) and C2 drops from
            // 89.5 % to 85.7 % of the copy peak (profiles/r01/pair_load_order_ab.log).  A runtime-selectable cache operator
            // (load_mode bit 1, never set in production) puts every load pair in its own basic block, which pins the order.
            if (p.load_mode & 2) { xv[k][i] = ldg_cs(xr + v); yv[k][i] = ldg_cs(yr + v); }
            else { xv[k][i] = ldg_stream(xr + v); yv[k][i] = ldg_stream(yr + v); }
#else
            xv[k][i] = ldg_stream(xr + v);
            yv[k][i] = ldg_stream(yr + v);
#endif
          } else {
            xv[k][i] = make_uint4(0, 0, 0, 0);
            yv[k][i] = make_uint4(0, 0, 0, 0);
          }
        }
      }
    }
    int label[ROWS];
    float gup[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      label[k] = 0;
      gup[k] = 0.f;
      if (MODE == kModeFused && live[k]) label[k] = (int)(__ldg(p.labels + rows[k]) != 0);
      if (MODE == kModeBwd && live[k]) gup[k] = __ldg(p.gsim + rows[k]);
    }

    RowSums s[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      s[k] = RowSums{0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (lane + 32 * i < nvec) {   // masked lanes must not add the l1/l2 eps
          float fx[E], fy[E];
          unpack<T>(xv[k][i], fx);
          unpack<T>(yv[k][i], fy);
          accumulate<MEASURE, COSLOSS>(fx, fy, E, s[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      if (MEASURE == IA_INNER || MEASURE == IA_COSINE || COSLOSS) s[k].xy = warp_sum(s[k].xy);
      if (MEASURE == IA_COSINE || COSLOSS) { s[k].xx = warp_sum(s[k].xx); s[k].yy = warp_sum(s[k].yy); }
      if (MEASURE == IA_L1 || MEASURE == IA_L2) s[k].dist = warp_sum(s[k].dist);
    }

#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      if (!live[k]) continue;
      const int64_t row = rows[k];
      float nx = 1.f, ny = 1.f;
      float sim = score_from_sums<MEASURE>(s[k], nx, ny);
      if (bad[k]) { sim = __int_as_float(0x7fc00000); s[k].xy = sim; }

      if (MODE != kModeBwd && lane == 0) {
        if (p.sim) p.sim[row] = sim;
        const float pr = prob_of<MEASURE>(sim);
        if (p.probs) p.probs[row] = pr;
        if (MODE == kModeFwd && p.labels_out) p.labels_out[row] = (uint8_t)((double)pr >= p.threshold);
      }
      if (MODE == kModeFwd) continue;

      GradCoef c;
      if (MODE == kModeFused) {
        float li, g;
        if (COSLOSS) {
          // nn.CosineEmbeddingLoss (text.py:1401,1471): c = xy / sqrt((xx+eps)(yy+eps))
          const float a = s[k].xx + kCosEmbEps, b = s[k].yy + kCosEmbEps;
          const float den = sqrtf(a * b);
          const float cs = s[k].xy / den;
          float gc;
          if (label[k]) { li = 1.f - cs; gc = -1.f; }
          else { li = fmaxf(0.f, cs - p.margin); gc = (cs - p.margin) > 0.f ? 1.f : 0.f; }
          gc *= gscale;
          c.A = gc / den; c.Bx = gc * cs / a; c.By = gc * cs / b;
        } else {
          li = scalar_loss(p.loss, sim, label[k], p.margin, g);
          c = score_grad_coef<MEASURE>(s[k], sim, nx, ny, g * gscale);
        }
        if (bad[k]) li = sim;   // NaN (fmaxf in the hinge / cosine-embedding losses would swallow it)
        if (p.reduction == IA_RED_NONE) { if (lane == 0) p.loss_out[row] = li; }
        else loss_acc += li;
      } else {
        c = score_grad_coef<MEASURE>(s[k], sim, nx, ny, gup[k]);
      }

      if (p.dx != nullptr) {
        G* dxr = static_cast<G*>(p.dx) + row * p.lddx;
        G* dyr = static_cast<G*>(p.dy) + row * p.lddy;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          if (v < nvec) {
            float fx[E], fy[E], gx[E], gy[E];
            unpack<T>(xv[k][i], fx);
            unpack<T>(yv[k][i], fy);
#pragma unroll
            for (int j = 0; j < E; ++j) {
              if (MEASURE == IA_INNER && !COSLOSS) {
                gx[j] = __fmul_rn(c.A, fy[j]);          // Bx = By = 0
                gy[j] = __fmul_rn(c.A, fx[j]);
              } else if (kGradCosForm) {
                // explicit rounding points: every instantiation of this kernel (ROWS = 1 / 2, gathered, generic) must produce
                // the same bits, whatever the compiler would choose to contract
                gx[j] = __fmaf_rn(c.A, fy[j], -__fmul_rn(c.Bx, fx[j]));
                gy[j] = __fmaf_rn(c.A, fx[j], -__fmul_rn(c.By, fy[j]));
              } else if (MEASURE == IA_L1) {
                const float dd = fx[j] - fy[j] + kPdistEps;
                gx[j] = dd > 0.f ? c.A : (dd < 0.f ? -c.A : 0.f);
                gy[j] = -gx[j];
              } else {
                const float dd = __fadd_rn(__fsub_rn(fx[j], fy[j]), kPdistEps);
                gx[j] = __fmul_rn(c.A, dd);
                gy[j] = -gx[j];
              }
            }
            if (ACT) {
              // tanh (and dropout) backward fused into the store: t = out / scale where kept; with dropout active a dropped
              // element is recognised by its exact zero (a kept tanh value is 0 only for a pre-activation of exactly 0)
              const float sc = p.act_bwd, inv = 1.0f / p.act_bwd;
              const bool drop = p.act_bwd != 1.0f;
#pragma unroll
              for (int j = 0; j < E; ++j) {
                const float tx = fx[j] * inv, ty = fy[j] * inv;
                gx[j] = (drop && fx[j] == 0.f) ? 0.f : __fmul_rn(__fmul_rn(gx[j], sc), __fmaf_rn(-tx, tx, 1.0f));
                gy[j] = (drop && fy[j] == 0.f) ? 0.f : __fmul_rn(__fmul_rn(gy[j], sc), __fmaf_rn(-ty, ty, 1.0f));
              }
            }
            Packer<G, E>::store(dxr + (int64_t)v * E, gx);
            Packer<G, E>::store(dyr + (int64_t)v * E, gy);
          }
        }
      }
    }
  }

  if (MODE == kModeFused) {
    if (p.reduction != IA_RED_NONE) {
      __shared__ float warp_loss[8];
      if (lane == 0) warp_loss[warp_in_block] = loss_acc;
      __syncthreads();
      double blk = 0.0;
      if (threadIdx.x == 0) {
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) blk += (double)warp_loss[w];
      }
      grid_sum_finish(blk, p.workspace, p.loss_out, p.loss_scale);
    }
  }
}

// Any D, any alignment: scalar loads, two sweeps over the row (the second one hits L1/L2).
template <typename T, typename G, int MEASURE, int MODE, bool COSLOSS, bool ACT = false>
__global__ void __launch_bounds__(256) pair_kernel_generic(const PairParams p) {
  pdl_wait();
  pdl_trigger();
  constexpr bool kGradCosForm = COSLOSS || MEASURE == IA_INNER || MEASURE == IA_COSINE;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  float loss_acc = 0.f;
  float gscale = p.grad_scale;
  if (MODE == kModeFused && p.upstream != nullptr) {
    const float u = __ldg(p.upstream);
    if (p.upstream_skip_one && u == 1.0f) return;
    gscale *= u;
  }
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_in_block; row < p.n; row += warps_total) {
    int64_t rx = row, ry = row;
    bool bad = false;
    if (p.xi != nullptr) {
      rx = __ldg(p.xi + row);
      ry = __ldg(p.yi + row);
      if ((uint64_t)rx >= (uint64_t)p.rows_x) { rx = 0; bad = true; }
      if ((uint64_t)ry >= (uint64_t)p.rows_y) { ry = 0; bad = true; }
    }
    const T* xr = static_cast<const T*>(p.x) + rx * p.ldx;
    const T* yr = static_cast<const T*>(p.y) + ry * p.ldy;
    RowSums s{0.f, 0.f, 0.f, 0.f};
    for (int j = lane; j < p.d; j += 32) {
      const float fx = to_float<T>(xr[j]), fy = to_float<T>(yr[j]);
      accumulate<MEASURE, COSLOSS>(&fx, &fy, 1, s);
    }
    s.xy = warp_sum(s.xy); s.xx = warp_sum(s.xx); s.yy = warp_sum(s.yy); s.dist = warp_sum(s.dist);
    float nx = 1.f, ny = 1.f;
    float sim = score_from_sums<MEASURE>(s, nx, ny);
    if (bad) { sim = __int_as_float(0x7fc00000); s.xy = sim; }
    if (MODE != kModeBwd && lane == 0) {
      if (p.sim) p.sim[row] = sim;
      const float pr = prob_of<MEASURE>(sim);
      if (p.probs) p.probs[row] = pr;
      if (MODE == kModeFwd && p.labels_out) p.labels_out[row] = (uint8_t)((double)pr >= p.threshold);
    }
    if (MODE == kModeFwd) continue;
    GradCoef c;
    if (MODE == kModeFused) {
      const int label = (int)(__ldg(p.labels + row) != 0);
      float li, g;
      if (COSLOSS) {
        const float a = s.xx + kCosEmbEps, b = s.yy + kCosEmbEps;
        const float den = sqrtf(a * b);
        const float cs = s.xy / den;
        float gc;
        if (label) { li = 1.f - cs; gc = -1.f; }
        else { li = fmaxf(0.f, cs - p.margin); gc = (cs - p.margin) > 0.f ? 1.f : 0.f; }
        gc *= gscale;
        c.A = gc / den; c.Bx = gc * cs / a; c.By = gc * cs / b;
      } else {
        li = scalar_loss(p.loss, sim, label, p.margin, g);
        c = score_grad_coef<MEASURE>(s, sim, nx, ny, g * gscale);
      }
      if (bad) li = sim;
      if (p.reduction == IA_RED_NONE) { if (lane == 0) p.loss_out[row] = li; }
      else loss_acc += li;
    } else {
      c = score_grad_coef<MEASURE>(s, sim, nx, ny, __ldg(p.gsim + row));
    }
    if (p.dx != nullptr) {
      G* dxr = static_cast<G*>(p.dx) + row * p.lddx;
      G* dyr = static_cast<G*>(p.dy) + row * p.lddy;
      for (int j = lane; j < p.d; j += 32) {
        const float fx = to_float<T>(xr[j]), fy = to_float<T>(yr[j]);
        float gx, gy;
        if (MEASURE == IA_INNER && !COSLOSS) {
          gx = __fmul_rn(c.A, fy);
          gy = __fmul_rn(c.A, fx);
        } else if (kGradCosForm) {
          gx = __fmaf_rn(c.A, fy, -__fmul_rn(c.Bx, fx));
          gy = __fmaf_rn(c.A, fx, -__fmul_rn(c.By, fy));
        } else if (MEASURE == IA_L1) {
          const float dd = fx - fy + kPdistEps;
          gx = dd > 0.f ? c.A : (dd < 0.f ? -c.A : 0.f);
          gy = -gx;
        } else {
          gx = __fmul_rn(c.A, __fadd_rn(__fsub_rn(fx, fy), kPdistEps));
          gy = -gx;
        }
        if (ACT) {
          const float sc = p.act_bwd, inv = 1.0f / p.act_bwd;
          const bool drop = p.act_bwd != 1.0f;
          const float tx = fx * inv, ty = fy * inv;
          gx = (drop && fx == 0.f) ? 0.f : __fmul_rn(__fmul_rn(gx, sc), __fmaf_rn(-tx, tx, 1.0f));
          gy = (drop && fy == 0.f) ? 0.f : __fmul_rn(__fmul_rn(gy, sc), __fmaf_rn(-ty, ty, 1.0f));
        }
        dxr[j] = from_float<G>(gx);
        dyr[j] = from_float<G>(gy);
      }
    }
  }
  if (MODE == kModeFused && p.reduction != IA_RED_NONE) {
    __shared__ float warp_loss[8];
    if (lane == 0) warp_loss[warp_in_block] = loss_acc;
    __syncthreads();
    double blk = 0.0;
    if (threadIdx.x == 0)
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) blk += (double)warp_loss[w];
    grid_sum_finish(blk, p.workspace, p.loss_out, p.loss_scale);
  }
}

// ---------------------------------------------------------------------------------- launch
template <void (*kernel)(const PairParams)>
int launch_rows(const PairParams& p, cudaStream_t stream) {
  constexpr int kThreads = 256, kWarps = 8;
  static int cached_bps_dev[kMaxDevices] = {};   // one instantiation per kernel -> per-kernel, per-device occupancy cache
  int& cached_bps = cached_bps_dev[device_slot()];
  if (cached_bps == 0) cached_bps = blocks_per_sm(kernel, kThreads);
  int64_t want = (p.n + kWarps - 1) / kWarps;
  int64_t cap = (int64_t)sm_count() * cached_bps;
  if (cap > kMaxPartials) cap = kMaxPartials;
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  IA_PDL_LAUNCH_CHECK(launch_pdl(kernel, grid, kThreads, 0, stream, p));
  return IA_OK;
}

// bulk-ring launch: dynamic shared memory = 8 warps x kBulkStages x (x row + y row) + barriers
template <void (*kernel)(const PairParams)>
int launch_rows_bulk(const PairParams& p, size_t elem_size, cudaStream_t stream) {
  constexpr int kThreads = 256, kWarps = 8;
  const size_t smem = (size_t)kWarps * kBulkStages * 2 * (size_t)p.d * elem_size + kWarps * kBulkStages * 8;
  static size_t configured_dev[kMaxDevices] = {};   // function attributes are per context: one slot per device
  static int cached_bps_dev[kMaxDevices] = {};
  size_t& configured = configured_dev[device_slot()];
  int& cached_bps = cached_bps_dev[device_slot()];
  if (smem > configured) {
    IA_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
    cached_bps = 0;
  }
  if (cached_bps == 0) cached_bps = blocks_per_sm(kernel, kThreads, smem);
  int64_t want = (p.n + kWarps - 1) / kWarps;
  int64_t cap = (int64_t)sm_count() * cached_bps;
  if (cap > kMaxPartials) cap = kMaxPartials;
  int grid = (int)(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  IA_PDL_LAUNCH_CHECK(launch_pdl(kernel, grid, kThreads, smem, stream, p));
  return IA_OK;
}

inline bool pair_bulk_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("IA_PAIR_BULK");
    on = e ? atoi(e) : 0;   // measured (profiles/r01/pair_configs_bulk_ab.log): direct loads win for 2 KB rows, tie for 4 KB
  }
  return on != 0;
}

// IA_PAIR_ROWS: 1 = one pair per warp iteration; default = group adjacent rows (see launch_pair_vpl)
inline int pair_rows_pref() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("IA_PAIR_ROWS");
    v = e ? atoi(e) : 2;
  }
  return v;
}

// activation-backward variants exist for the fused mode of 16-bit tensors with gradients in the same type (what the head's
// training path produces); anything else is refused loudly
template <typename T, typename G, int MEASURE, int MODE, bool COSLOSS>
int launch_pair_act(const PairParams& p, bool vec_ok, cudaStream_t stream) {
  if constexpr (MODE == kModeFused && sizeof(T) == 2 && sizeof(G) == 2) {
    constexpr int E = VecTraits<T>::kElems;
    const int nvec = p.d / E;
    if (!vec_ok || nvec > 32 * 8) return launch_rows<pair_kernel_generic<T, G, MEASURE, MODE, COSLOSS, true>>(p, stream);
    const size_t row_bytes = (size_t)p.d * sizeof(T);
    const bool group = pair_rows_pref() > 1 && p.n >= 16384 && row_bytes >= 1536 && row_bytes <= 3072;
    if (nvec <= 64) return launch_rows<pair_kernel<T, G, MEASURE, MODE, COSLOSS, 2, false, 1, true>>(p, stream);
    if (nvec <= 128) {
      if (group) return launch_rows<pair_kernel<T, G, MEASURE, MODE, COSLOSS, 4, false, 2, true>>(p, stream);
      return launch_rows<pair_kernel<T, G, MEASURE, MODE, COSLOSS, 4, false, 1, true>>(p, stream);
    }
    return launch_rows<pair_kernel<T, G, MEASURE, MODE, COSLOSS, 8, false, 1, true>>(p, stream);
  } else {
    set_error("activation backward (act_bwd_scale > 0) is offered for the fused launch on bf16 / fp16 tensors with gradients in the same type");
    return IA_ERR_UNSUPPORTED;
  }
}

template <typename T, typename G, int MEASURE, int MODE, bool COSLOSS>
int launch_pair_vpl(const PairParams& p, bool vec_ok, cudaStream_t stream) {
  constexpr int E = VecTraits<T>::kElems;
  const int nvec = p.d / E;
  if (p.act_bwd > 0.f) return launch_pair_act<T, G, MEASURE, MODE, COSLOSS>(p, vec_ok, stream);
  if (!vec_ok || nvec > 32 * 8) return launch_rows<pair_kernel_generic<T, G, MEASURE, MODE, COSLOSS>>(p, stream);
  // big rows (>= 1 KB) of the forward / fused kernels go through the bulk-copy ring
  if (MODE != kModeBwd && pair_bulk_enabled() && (size_t)p.d * sizeof(T) >= 1024 && p.n >= 1024 && p.xi == nullptr) {
    if (nvec <= 64) return launch_rows_bulk<pair_kernel<T, G, MEASURE, MODE, COSLOSS, 2, true, 1>>(p, sizeof(T), stream);
    if (nvec <= 128) return launch_rows_bulk<pair_kernel<T, G, MEASURE, MODE, COSLOSS, 4, true, 1>>(p, sizeof(T), stream);
    return launch_rows_bulk<pair_kernel<T, G, MEASURE, MODE, COSLOSS, 8, true, 1>>(p, sizeof(T), stream);
  }
  // ROWS adjacent pairs per warp iteration: one contiguous chunk of ROWS * row_bytes per tensor.  Measured on B200
  // (profiles/r01/pair_rows_ab.log), kernels that also write gradients: 2 KB rows (bf16 D=1024) gain 7 % when paired
  // (84.6 -> 91 % of the copy peak), 3 KB rows (fp32 D=768) gain 6 %; 1 KB rows and 4 KB rows are best alone, as is
  // the forward-only kernel (already at the read-only peak).  IA_PAIR_ROWS=1 turns the grouping off.
  const size_t row_bytes = (size_t)p.d * sizeof(T);
  const bool group = MODE != kModeFwd && pair_rows_pref() > 1 && p.n >= 16384 && row_bytes >= 1536 && row_bytes <= 3072 &&
                     p.xi == nullptr && p.yi == nullptr;   // gathered rows are not adjacent in memory
  if (nvec <= 64) return launch_rows<pair_kernel<T, G, MEASURE, MODE, COSLOSS, 2, false, 1>>(p, stream);
  if (nvec <= 128) {
    if (group) return launch_rows<pair_kernel<T, G, MEASURE, MODE, COSLOSS, 4, false, 2>>(p, stream);
    return launch_rows<pair_kernel<T, G, MEASURE, MODE, COSLOSS, 4, false, 1>>(p, stream);
  }
  if (group) return launch_rows<pair_kernel<T, G, MEASURE, MODE, COSLOSS, 8, false, 2>>(p, stream);
  return launch_rows<pair_kernel<T, G, MEASURE, MODE, COSLOSS, 8, false, 1>>(p, stream);
}

template <typename T, typename G, int MODE, bool COSLOSS>
int launch_pair_measure(int measure, const PairParams& p, bool vec_ok, cudaStream_t stream) {
  switch (measure) {
    case IA_INNER: return launch_pair_vpl<T, G, IA_INNER, MODE, COSLOSS>(p, vec_ok, stream);
    case IA_COSINE: return launch_pair_vpl<T, G, IA_COSINE, MODE, COSLOSS>(p, vec_ok, stream);
    case IA_L1: return launch_pair_vpl<T, G, IA_L1, MODE, COSLOSS>(p, vec_ok, stream);
    case IA_L2: return launch_pair_vpl<T, G, IA_L2, MODE, COSLOSS>(p, vec_ok, stream);
  }
  set_error("Unsupported similarty measure: %d", measure);
  return IA_ERR_INVALID;
}

// one entry per (T, G): implemented in pair_<dtype>.cu
template <typename T, typename G>
int launch_pair(int mode, bool cosloss, int measure, const PairParams& p, bool vec_ok, cudaStream_t stream) {
  if (mode == kModeFwd) return launch_pair_measure<T, G, kModeFwd, false>(measure, p, vec_ok, stream);
  if (mode == kModeBwd) return launch_pair_measure<T, G, kModeBwd, false>(measure, p, vec_ok, stream);
  if (cosloss) return launch_pair_measure<T, G, kModeFused, true>(measure, p, vec_ok, stream);
  return launch_pair_measure<T, G, kModeFused, false>(measure, p, vec_ok, stream);
}

int launch_pair_f32_f32(int mode, bool cosloss, int measure, const PairParams& p, bool vec_ok, cudaStream_t s);
int launch_pair_bf16_bf16(int mode, bool cosloss, int measure, const PairParams& p, bool vec_ok, cudaStream_t s);
int launch_pair_bf16_f32(int mode, bool cosloss, int measure, const PairParams& p, bool vec_ok, cudaStream_t s);
int launch_pair_f16_f16(int mode, bool cosloss, int measure, const PairParams& p, bool vec_ok, cudaStream_t s);
int launch_pair_f16_f32(int mode, bool cosloss, int measure, const PairParams& p, bool vec_ok, cudaStream_t s);

}  // namespace ia
