// "softmax" similarity measure: TwoTowerClassificationHead (reference src/models/base.py:103-117)
// fused with CrossEntropyLoss forward + backward (reference src/models/text.py:1408-1409,1473).
//
//   logits = [x ; y] . W^T + b     W: [2, 2h] fp32        probs = softmax(logits)
//
// HBM-bound like the pair kernels: one warp owns one pair, x and y rows are read once into registers
// (the torch.cat copy of the reference never exists), W lives in shared memory, and dx, dy are written
// from the same registers.  dW is accumulated in registers per lane over all rows a warp visits and
// reduced block -> workspace -> finalize kernel in a fixed order (bit-reproducible run to run).
// Two classes: delta = p1 - label is formed without cancellation (label 1 -> -p0), dlogit1 = delta,
// dlogit0 = -delta, so dW[0] = -dW[1] and db[0] = -db[1].
#include <cstdlib>

#include "softmax_head.cuh"

namespace ia {


// Shared-memory layout of one weight-row half (h floats): "lane-interleaved chunks" -- the float4 a lane needs for
// input vector v = lane + 32*i, chunk q sits at float4 index (i*(E/4)+q)*32 + lane, so every LDS.128 of a warp
// covers 512 contiguous bytes (conflict free).
template <int E>
__device__ __forceinline__ int swz(int e) {
  const int v = e / E, j = e % E;
  return (((v >> 5) * (E / 4) + (j >> 2)) * 32 + (v & 31)) * 4 + (j & 3);
}

constexpr int kHeadMaxStages = 4;   // rows in flight per warp in the bulk-copy ring (TRAIN kernel), fewer if smem is short

// The training kernel is bound by SHARED-MEMORY bandwidth if every row re-reads the whole W tile (6 float4 arrays,
// 24 KB per row per warp at h = 1024): it therefore processes RR rows per iteration -- each W chunk is loaded once and
// applied to RR rows held in registers -- and keeps 2*VPL*E dW accumulators per lane.  With that many registers only
// 8 warps fit per SM, so the rows arrive through a per-warp shared-memory ring filled with cp.async (every lane copies
// exactly the vectors it will consume: no cross-lane synchronisation), p.stages row groups in flight per warp.
// (Measured first: a cp.async.bulk ring and a 16-warp variant both sat at 55 % -- the W re-reads were the limit.)
// The forward-only kernel is light: plain streaming loads, RR = 1.
// FULL: h fills every lane's VPL vectors exactly (nvec == 32 * VPL): the per-vector bounds tests disappear.  Rows past the end
// of the batch (last group only) are CLAMPED to the last row for loading and masked at the stores (delta = 0), so the
// hot loop has no data-dependent branches.
// NW / REREAD: the 8-warp training kernel keeps the RR rows of an iteration in registers next to the dW accumulators (~250
// registers per thread).  REREAD trades registers for shared-memory reads: rows stay in the ring and are read where they are
// used (forward pass, then again in the gradient pass), which fits 12 warps per SM (3 per scheduler instead of 2) -- the
// kernel is bound by dependent-issue latency, not by bandwidth.
template <typename T, typename G, bool TRAIN, int VPL, int RR, bool FULL, int NW, bool REREAD>
__global__ void __launch_bounds__(NW * 32, TRAIN ? 1 : 2) softmax_head_kernel(const HeadParams p) {
  pdl_wait();       // launched with programmatic stream serialization (launch_pdl)
  constexpr int E = VecTraits<T>::kElems;
  constexpr int C = E / 4;            // float4 chunks per 128-bit input vector
  constexpr int P4 = VPL * 32 * C;    // float4 per (padded) weight-row half
  extern __shared__ float4 smem4[];
  float gscale = p.grad_scale;
  if (TRAIN && p.upstream != nullptr) {   // gradient recomputation with the upstream scalar folded in before the one rounding
    const float u = __ldg(p.upstream);
    if (p.upstream_skip_one && u == 1.0f) return;
    gscale *= u;
  }
  // [w0x | w0y | w1x | w1y | (TRAIN) wdx | wdy | sacc[2h] | ring]
  const float4* w0x = smem4;
  const float4* w0y = smem4 + P4;
  const float4* w1x = smem4 + 2 * P4;
  const float4* w1y = smem4 + 3 * P4;
  const float4* wdx = smem4 + 4 * P4;
  const float4* wdy = smem4 + 5 * P4;
  float* sw = reinterpret_cast<float*>(smem4);
  float* sacc = sw + 24 * P4;
  const int h = p.h, h2 = 2 * p.h;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warps_total = (int64_t)gridDim.x * NW;
  const int nvec = h / E;
  const uint32_t row_bytes = (uint32_t)h * (uint32_t)sizeof(T);
  uint8_t* ring_base = reinterpret_cast<uint8_t*>(sacc + ((h2 + 3) & ~3));
  const size_t stage_bytes = (size_t)RR * 2 * row_bytes;   // RR x [x row | y row]
  uint8_t* ring = ring_base + (size_t)wib * p.stages * stage_bytes;
  const int64_t n_groups = (p.n + RR - 1) / RR;           // a warp owns RR adjacent rows per iteration
  const int64_t grp0 = (int64_t)blockIdx.x * NW + wib;
  auto arm = [&](int stage, int64_t grp) {
    if (grp < n_groups) {
#pragma unroll
      for (int k = 0; k < RR; ++k) {
        const int64_t row = min(grp * RR + k, p.n - 1);
        {
          const uint32_t dst = smem_u32(ring + (size_t)stage * stage_bytes + (size_t)k * 2 * row_bytes);
          const uint4* xr = reinterpret_cast<const uint4*>(static_cast<const T*>(p.x) + row * p.ldx);
          const uint4* yr = reinterpret_cast<const uint4*>(static_cast<const T*>(p.y) + row * p.ldy);
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            const int v = lane + 32 * i;
            if (FULL || v < nvec) {
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)v * 16u), "l"(xr + v) : "memory");
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + row_bytes + (uint32_t)v * 16u), "l"(yr + v) : "memory");
            }
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // always commit: group counting stays uniform
  };
  if (TRAIN) {
    for (int s = 0; s < p.stages; ++s) arm(s, grp0 + s * warps_total);
  }
  // the W tiles are built while the first rows are already streaming in
  for (int i0 = threadIdx.x; i0 < h2; i0 += 4 * blockDim.x) {   // four independent loads in flight per thread
    float a[4], c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      a[u] = i < h2 ? __ldg(p.w + i) : 0.f;
      c[u] = i < h2 ? __ldg(p.w + h2 + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < h2) {
        const int half = i >= h, e = half ? i - h : i;
        const int pos = half * (4 * P4) + swz<E>(e);
        sw[pos] = a[u];
        sw[8 * P4 + pos] = c[u];
        if (TRAIN) { sw[16 * P4 + pos] = c[u] - a[u]; sacc[i] = 0.f; }
      }
    }
  }
  __syncthreads();
  const float b0 = p.b[0], b1 = p.b[1];

  float accx[TRAIN ? VPL * E : 1], accy[TRAIN ? VPL * E : 1];
  if (TRAIN) {
#pragma unroll
    for (int i = 0; i < VPL * E; ++i) { accx[i] = 0.f; accy[i] = 0.f; }
  }
  float loss_acc = 0.f, db_acc = 0.f;

  long long label_next[RR];
#pragma unroll
  for (int k = 0; k < RR; ++k) label_next[k] = (TRAIN && grp0 * RR + k < p.n) ? __ldg(p.labels + grp0 * RR + k) : 0;
  int it = 0;
  for (int64_t grp = grp0; grp < n_groups; grp += warps_total, ++it) {
    const int stage = TRAIN ? it % p.stages : 0;
    uint4 xv[RR][VPL], yv[RR][VPL];
    bool live[RR];
    int64_t rows[RR];
#pragma unroll
    for (int k = 0; k < RR; ++k) { rows[k] = grp * RR + k; live[k] = rows[k] < p.n; }
    long long label[RR];   // this group's labels were loaded one iteration ago (software pipelined: L2 latency off the path)
#pragma unroll
    for (int k = 0; k < RR; ++k) label[k] = label_next[k];
    if (TRAIN) {
      const int64_t gn = grp + warps_total;
#pragma unroll
      for (int k = 0; k < RR; ++k) label_next[k] = (gn * RR + k < p.n) ? __ldg(p.labels + gn * RR + k) : 0;
    }
    if (TRAIN) {
      if (p.stages == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
      else if (p.stages == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
      else if (p.stages == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (!REREAD) {
#pragma unroll
        for (int k = 0; k < RR; ++k) {
          const uint4* xs = reinterpret_cast<const uint4*>(ring + (size_t)stage * stage_bytes + (size_t)k * 2 * row_bytes);
          const uint4* ys = reinterpret_cast<const uint4*>(ring + (size_t)stage * stage_bytes + (size_t)k * 2 * row_bytes + row_bytes);
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            const int v = lane + 32 * i;
            if (FULL || v < nvec) { xv[k][i] = xs[v]; yv[k][i] = ys[v]; }
            else { xv[k][i] = make_uint4(0, 0, 0, 0); yv[k][i] = make_uint4(0, 0, 0, 0); }
          }
        }
        arm(stage, grp + (int64_t)p.stages * warps_total);   // the slot is in registers now: refill it
      }
    } else {
#pragma unroll
      for (int k = 0; k < RR; ++k) {
        const int64_t rl = min(rows[k], p.n - 1);
        const uint4* xr = reinterpret_cast<const uint4*>(static_cast<const T*>(p.x) + rl * p.ldx);
        const uint4* yr = reinterpret_cast<const uint4*>(static_cast<const T*>(p.y) + rl * p.ldy);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          if (FULL || v < nvec) {
            // loads pinned in address order: one basic block per pair (see pair_kernels.cuh, load loop)
            if (p.load_mode & 2) { xv[k][i] = ldg_cs(xr + v); yv[k][i] = ldg_cs(yr + v); }
            else { xv[k][i] = ldg_stream(xr + v); yv[k][i] = ldg_stream(yr + v); }
          }
          else { xv[k][i] = make_uint4(0, 0, 0, 0); yv[k][i] = make_uint4(0, 0, 0, 0); }
        }
      }
    }
    // ---- logits: every W chunk is read once and applied to all RR rows
    float l0[RR], l1[RR], l0b[RR], l1b[RR];
#pragma unroll
    for (int k = 0; k < RR; ++k) { l0[k] = 0.f; l1[k] = 0.f; l0b[k] = 0.f; l1b[k] = 0.f; }
    // row vector (k, i) of this lane: from registers, or (REREAD) from the ring slot it was copied into
    const uint4* ring_stage = reinterpret_cast<const uint4*>(ring + (size_t)stage * stage_bytes);
    const int row_vecs = (int)(row_bytes / 16);
    auto x_vec = [&](int k, int i) -> uint4 { return (TRAIN && REREAD) ? ring_stage[(2 * k) * row_vecs + lane + 32 * i] : xv[k][i]; };
    auto y_vec = [&](int k, int i) -> uint4 { return (TRAIN && REREAD) ? ring_stage[(2 * k + 1) * row_vecs + lane + 32 * i] : yv[k][i]; };
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (FULL || lane + 32 * i < nvec) {
        float4 a0[C], a1[C], c0[C], c1[C];
#pragma unroll
        for (int q = 0; q < C; ++q) {
          const int f4 = (i * C + q) * 32 + lane;
          a0[q] = w0x[f4]; a1[q] = w0y[f4]; c0[q] = w1x[f4]; c1[q] = w1y[f4];
        }
#pragma unroll
        for (int k = 0; k < RR; ++k) {
          float fx[E], fy[E];
          unpack<T>(x_vec(k, i), fx);
          unpack<T>(y_vec(k, i), fy);
#pragma unroll
          for (int q = 0; q < C; ++q) {
            l0[k] = fmaf(fx[4 * q + 0], a0[q].x, l0[k]); l0[k] = fmaf(fx[4 * q + 1], a0[q].y, l0[k]);
            l0[k] = fmaf(fx[4 * q + 2], a0[q].z, l0[k]); l0[k] = fmaf(fx[4 * q + 3], a0[q].w, l0[k]);
            l0b[k] = fmaf(fy[4 * q + 0], a1[q].x, l0b[k]); l0b[k] = fmaf(fy[4 * q + 1], a1[q].y, l0b[k]);
            l0b[k] = fmaf(fy[4 * q + 2], a1[q].z, l0b[k]); l0b[k] = fmaf(fy[4 * q + 3], a1[q].w, l0b[k]);
            l1[k] = fmaf(fx[4 * q + 0], c0[q].x, l1[k]); l1[k] = fmaf(fx[4 * q + 1], c0[q].y, l1[k]);
            l1[k] = fmaf(fx[4 * q + 2], c0[q].z, l1[k]); l1[k] = fmaf(fx[4 * q + 3], c0[q].w, l1[k]);
            l1b[k] = fmaf(fy[4 * q + 0], c1[q].x, l1b[k]); l1b[k] = fmaf(fy[4 * q + 1], c1[q].y, l1b[k]);
            l1b[k] = fmaf(fy[4 * q + 2], c1[q].z, l1b[k]); l1b[k] = fmaf(fy[4 * q + 3], c1[q].w, l1b[k]);
          }
        }
      }
    }
    float delta[RR];
#pragma unroll
    for (int k = 0; k < RR; ++k) {
      const float z0 = warp_sum(l0[k] + l0b[k]) + b0;
      const float z1 = warp_sum(l1[k] + l1b[k]) + b1;
      const float m = fmaxf(z0, z1);
      const float e0 = expf(z0 - m), e1 = expf(z1 - m);
      const float den = e0 + e1;
      const float p0 = e0 / den, p1 = e1 / den;
      delta[k] = 0.f;
      if (live[k]) {
        if (lane == 0) {
          if (p.logits) { p.logits[2 * rows[k]] = z0; p.logits[2 * rows[k] + 1] = z1; }
          if (p.probs) { p.probs[2 * rows[k]] = p0; p.probs[2 * rows[k] + 1] = p1; }
        }
        if (TRAIN) {
          loss_acc += logf(den) - ((label[k] != 0 ? z1 : z0) - m);      // -log softmax[label]
          if ((unsigned long long)label[k] > 1ull) loss_acc = __int_as_float(0x7fc00000);   // label outside {0,1}: the loss is NaN, loudly
          delta[k] = (label[k] != 0 ? -p0 : p1) * gscale;          // d loss / d logit1 ( = -d loss / d logit0 )
          db_acc += delta[k];
        }
      }
    }
    if (!TRAIN) continue;
    // ---- gradients: dx = delta * (W1x - W0x), dW accumulators += delta * x; again one W read for RR rows
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      if (FULL || v < nvec) {
        float4 d0[C], d1[C];
#pragma unroll
        for (int q = 0; q < C; ++q) { d0[q] = wdx[(i * C + q) * 32 + lane]; d1[q] = wdy[(i * C + q) * 32 + lane]; }
#pragma unroll
        for (int k = 0; k < RR; ++k) {
          float fx[E], fy[E], gx[E], gy[E];   // a clamped (dead) row has delta = 0: it adds nothing and is not stored
          unpack<T>(x_vec(k, i), fx);
          unpack<T>(y_vec(k, i), fy);
#pragma unroll
          for (int q = 0; q < C; ++q) {
            gx[4 * q + 0] = delta[k] * d0[q].x; gx[4 * q + 1] = delta[k] * d0[q].y; gx[4 * q + 2] = delta[k] * d0[q].z; gx[4 * q + 3] = delta[k] * d0[q].w;
            gy[4 * q + 0] = delta[k] * d1[q].x; gy[4 * q + 1] = delta[k] * d1[q].y; gy[4 * q + 2] = delta[k] * d1[q].z; gy[4 * q + 3] = delta[k] * d1[q].w;
          }
#pragma unroll
          for (int j = 0; j < E; ++j) {
            accx[i * E + j] = fmaf(delta[k], fx[j], accx[i * E + j]);
            accy[i * E + j] = fmaf(delta[k], fy[j], accy[i * E + j]);
          }
          if (p.dx && live[k]) {
            Packer<G, E>::store(static_cast<G*>(p.dx) + rows[k] * p.lddx + (int64_t)v * E, gx);
            Packer<G, E>::store(static_cast<G*>(p.dy) + rows[k] * p.lddy + (int64_t)v * E, gy);
          }
        }
      }
    }
    if (REREAD) arm(stage, grp + (int64_t)p.stages * warps_total);   // both passes have read the slot: refill it
  }

  if (TRAIN) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // block reduction of the per-lane dW accumulators in a FIXED order (warp 0 + warp 1 + ...), done in parallel:
    // every warp parks its accumulators in the (now idle) ring memory, then each thread adds the NW values of its
    // columns.  Needs NW * 2h floats of ring space (always true for >= 1 stage of 16-bit rows x 2 or fp32 rows).
    __syncthreads();
    float* park = reinterpret_cast<float*>(ring_base);
    const bool park_fits = (size_t)NW * h2 * sizeof(float) <= (size_t)NW * p.stages * stage_bytes;
    if (park_fits) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) {
#pragma unroll
          for (int j = 0; j < E; j += 4) {
            *reinterpret_cast<float4*>(park + (size_t)wib * h2 + v * E + j) = make_float4(accx[i * E + j], accx[i * E + j + 1], accx[i * E + j + 2], accx[i * E + j + 3]);
            *reinterpret_cast<float4*>(park + (size_t)wib * h2 + h + v * E + j) = make_float4(accy[i * E + j], accy[i * E + j + 1], accy[i * E + j + 2], accy[i * E + j + 3]);
          }
        }
      }
      __syncthreads();
      for (int c = threadIdx.x; c < h2; c += blockDim.x) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) a += park[(size_t)w * h2 + c];
        sacc[c] = a;
      }
      __syncthreads();
    } else {
      for (int w = 0; w < NW; ++w) {
        if (wib == w) {
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            const int v = lane + 32 * i;
            if (v < nvec) {
#pragma unroll
              for (int j = 0; j < E; ++j) {
                sacc[v * E + j] += accx[i * E + j];
                sacc[h + v * E + j] += accy[i * E + j];
              }
            }
          }
        }
        __syncthreads();
      }
    }
    __shared__ float wl[NW], wdb[NW];
    if (lane == 0) { wl[wib] = loss_acc; wdb[wib] = db_acc; }
    __syncthreads();
    float* part = reinterpret_cast<float*>(static_cast<char*>(p.workspace) + kWorkspaceBytes) +
                  (size_t)blockIdx.x * (h2 + 2);
    for (int i = threadIdx.x; i < h2; i += blockDim.x) part[i] = sacc[i];
    double blk = 0.0;
    if (threadIdx.x == 0) {
      float dbs = 0.f;
      for (int w = 0; w < NW; ++w) { blk += (double)wl[w]; dbs += wdb[w]; }
      part[h2] = dbs;
    }
    // the dW finalize launch may be scheduled from here on (it waits for this grid to complete before it reads the partials).
    // Not earlier: its CTAs would sit next to this kernel's for the whole run (measured: 72 -> 75 us at h = 768).
    pdl_trigger();
    grid_sum_finish(blk, p.workspace, p.loss_out, p.loss_scale);
  }
}

// dW[1][j] = sum over blocks of partial[b][j] (fixed order: lane-strided partial sums, then a shuffle tree);
// dW[0] = -dW[1]; same for db.  One warp per column.
__global__ void __launch_bounds__(256) softmax_head_finalize(const float* partials, int nblocks, int h2, float* dw,
                                                             float* db, const float* upstream, int skip_one) {
  pdl_wait();       // scheduled while the main kernel drains; the partials are read only once it has completed
  pdl_trigger();
  if (upstream != nullptr && skip_one && __ldg(upstream) == 1.0f) return;   // the main kernel left at once: no new partials
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j > h2) return;
  float acc = 0.f;
  for (int b = lane; b < nblocks; b += 32) acc += partials[(size_t)b * (h2 + 2) + j];
  acc = warp_sum(acc);
  if (lane == 0) {
    if (j < h2) {
      if (dw) { dw[h2 + j] = acc; dw[j] = -acc; }
    } else if (db) {
      db[1] = acc; db[0] = -acc;
    }
  }
}

template <typename T, typename G, bool TRAIN, int VPL, int RR, int NW = 8, bool REREAD = false>
static int launch_head_rr(const HeadParams& p_in, int stages, size_t smem, cudaStream_t stream, float* dw, float* db) {
  const bool full = p_in.h / VecTraits<T>::kElems == 32 * VPL;
  auto kernel = full ? softmax_head_kernel<T, G, TRAIN, VPL, RR, true, NW, REREAD>
                     : softmax_head_kernel<T, G, TRAIN, VPL, RR, false, NW, REREAD>;
  HeadParams p = p_in;
  p.stages = stages;
  p.group = RR;
  static size_t configured_smem[kMaxDevices][2] = {};   // function attributes are per context: one slot per device
  const int slot = device_slot();
  if (smem > configured_smem[slot][full]) {
    IA_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_smem[slot][full] = smem;
  }
  const int64_t groups = (p.n + RR - 1) / RR;
  int64_t want = (groups + NW - 1) / NW;
  int64_t cap = (int64_t)sm_count() * (TRAIN ? 1 : blocks_per_sm(kernel, NW * 32, smem));
  if (cap > kHeadMaxGrid) cap = kHeadMaxGrid;
  const int grid = (int)(want < cap ? want : cap);
  IA_PDL_LAUNCH_CHECK(launch_pdl(kernel, grid, NW * 32, smem, stream, p));
  if (TRAIN && (dw || db)) {
    const float* partials = reinterpret_cast<const float*>(static_cast<const char*>(p.workspace) + kWorkspaceBytes);
    const int h2 = 2 * p.h;
    IA_PDL_LAUNCH_CHECK(launch_pdl(softmax_head_finalize, (h2 + 1 + 7) / 8, 256, 0, stream, partials, grid, h2, dw, db, p.upstream, p.upstream_skip_one));
  }
  return IA_OK;
}

template <typename T, typename G, bool TRAIN, int VPL>
static int launch_head_one(const HeadParams& p, cudaStream_t stream, float* dw, float* db) {
  constexpr int P4 = VPL * 32 * (VecTraits<T>::kElems / 4);
  const size_t w_bytes = (size_t)(TRAIN ? 6 : 4) * P4 * 16;
  if (!TRAIN) {   // forward only: two adjacent rows per W read when both fit in registers (halves the shared-memory traffic)
    if (VPL <= 4 && p.n >= 4096) return launch_head_rr<T, G, false, (VPL <= 4 ? VPL : 4), 2>(p, 1, w_bytes, stream, dw, db);
    return launch_head_rr<T, G, false, VPL, 1>(p, 1, w_bytes, stream, dw, db);
  }
  const size_t fixed = w_bytes + (size_t)(((2 * p.h + 3) & ~3)) * 4;
  const size_t row_bytes = (size_t)p.h * sizeof(T);
  static const int max_stages = [] { const char* e = getenv("IA_HEAD_STAGES"); const int v = e ? atoi(e) : kHeadMaxStages; return v < 1 ? 1 : (v > kHeadMaxStages ? kHeadMaxStages : v); }();
  // Ring depth: measured, MORE row data in flight is not better -- beyond ~128 KB per SM the memory system queues up and
  // the kernel slows down (12 warps: 1 stage 103.2 us, 2 stages 108.3 us; 8 warps: 2 stages 109.5 us, 4 stages 111.5 us at
  // 65 536 x 1024 bf16), so the depth is capped by bytes in flight, then by what fits next to the W tiles.
  auto fit = [&](int rr, int nw = 8) {
    static const size_t inflight = [] { const char* e = getenv("IA_HEAD_INFLIGHT_KB"); return (size_t)(e ? atoi(e) : 128) * 1024; }();
    int st = (int)(inflight / ((size_t)nw * rr * 2 * row_bytes));
    st = st < 1 ? 1 : (st > max_stages ? max_stages : st);
    while (st > 1 && fixed + (size_t)nw * st * rr * 2 * row_bytes > 227 * 1024) --st;
    return (fixed + (size_t)nw * st * rr * 2 * row_bytes <= 227 * 1024) ? st : 0;
  };
  // 12 warps, rows re-read from the ring (see the kernel comment): 16-bit rows up to 2 KB, two stages of two rows per warp
  static const int mode12 = [] { const char* e = getenv("IA_HEAD_12W"); return e ? atoi(e) : 1; }();
  if (mode12 && VPL <= 4 && sizeof(T) == 2 && p.n >= 4096) {
    const int st12 = fit(2, 12);
    // park space of the block reduction: 12 warps x 2h floats must fit into the ring
    if (st12 >= 1 && (size_t)12 * 2 * p.h * 4 <= (size_t)12 * st12 * 2 * 2 * row_bytes)
      return launch_head_rr<T, G, true, (VPL <= 4 ? VPL : 4), 2, 12, true>(p, st12, fixed + (size_t)12 * st12 * 2 * 2 * row_bytes, stream, dw, db);
  }
  // fp32 rows are twice the bytes per element: one row per W read gives the same W traffic per byte as two 16-bit rows
  if (mode12 && sizeof(T) == 4 && p.n >= 4096) {
    const int st12 = fit(1, 12);
    if (st12 >= 1 && (size_t)12 * 2 * p.h * 4 <= (size_t)12 * st12 * 2 * row_bytes)
      return launch_head_rr<T, G, true, VPL, 1, 12, true>(p, st12, fixed + (size_t)12 * st12 * 2 * row_bytes, stream, dw, db);
  }
  // two rows per W read when the rows are small enough to keep both in registers (VPL <= 4) and the ring has >= 2 stages
  if (VPL <= 4 && p.n >= 4096) {
    const int st2 = fit(2);
    if (st2 >= 2) return launch_head_rr<T, G, true, (VPL <= 4 ? VPL : 4), 2>(p, st2, fixed + (size_t)8 * st2 * 2 * 2 * row_bytes, stream, dw, db);
  }
  const int st1 = fit(1);
  if (st1 >= 1) return launch_head_rr<T, G, true, VPL, 1>(p, st1, fixed + (size_t)8 * st1 * 2 * row_bytes, stream, dw, db);
  set_error("h too large for the shared-memory W tile + row ring");
  return IA_ERR_UNSUPPORTED;
}

template <typename T, typename G>
static int launch_head(const HeadParams& p, bool train, cudaStream_t stream, float* dw, float* db) {
  const int nvec = p.h / VecTraits<T>::kElems;
  if (train) {
    if (nvec <= 64) return launch_head_one<T, G, true, 2>(p, stream, dw, db);
    if (nvec <= 96) return launch_head_one<T, G, true, 3>(p, stream, dw, db);     // h = 768 in 16-bit types
    if (nvec <= 128) return launch_head_one<T, G, true, 4>(p, stream, dw, db);
    return launch_head_one<T, G, true, 8>(p, stream, dw, db);
  }
  if (nvec <= 64) return launch_head_one<T, G, false, 2>(p, stream, dw, db);
  if (nvec <= 96) return launch_head_one<T, G, false, 3>(p, stream, dw, db);
  if (nvec <= 128) return launch_head_one<T, G, false, 4>(p, stream, dw, db);
  return launch_head_one<T, G, false, 8>(p, stream, dw, db);
}

}  // namespace ia

using namespace ia;

extern "C" {

size_t ia_softmax_head_workspace_bytes(int64_t h) {
  return kWorkspaceBytes + sizeof(float) * (size_t)kHeadMaxGrid * (size_t)(2 * h + 2);
}

int ia_softmax_head_fwd_bwd(int dtype, int grad_dtype, const void* x, const void* y, int64_t ldx, int64_t ldy,
                            const float* w, const float* b, const int64_t* labels, int64_t n, int64_t h,
                            float* logits, float* probs, float* loss_out, void* dx, void* dy, int64_t lddx,
                            int64_t lddy, float* dw, float* db, float grad_scale, const float* upstream_dev,
                            int skip_if_one, void* workspace, size_t workspace_bytes, ia_stream_t stream) {
  if (n < 0 || h <= 0 || ldx < h || ldy < h || w == nullptr || b == nullptr || (n > 0 && (x == nullptr || y == nullptr))) {
    set_error("bad arguments");
    return IA_ERR_INVALID;
  }
  const bool train = labels != nullptr;
  if (train && loss_out == nullptr && upstream_dev == nullptr) { set_error("loss_out must not be NULL when labels are given"); return IA_ERR_INVALID; }
  if ((dx == nullptr) != (dy == nullptr)) { set_error("dx and dy must both be given or both be NULL"); return IA_ERR_INVALID; }
  if (grad_dtype != dtype && grad_dtype != IA_F32) { set_error("grad_dtype must equal dtype or be fp32"); return IA_ERR_UNSUPPORTED; }
  if (train && (workspace == nullptr || workspace_bytes < ia_softmax_head_workspace_bytes(h))) {
    set_error("workspace too small: need %zu bytes", ia_softmax_head_workspace_bytes(h));
    return IA_ERR_WORKSPACE;
  }
  if (train && loss_out == nullptr)   // gradient recomputation: the loss value is not wanted, it lands in the workspace header
    loss_out = reinterpret_cast<float*>(static_cast<char*>(workspace) + kWorkspaceScratchOffset);
  const int elems = dtype == IA_F32 ? 4 : 8;
  const size_t es = dtype == IA_F32 ? 4 : 2, gs = grad_dtype == IA_F32 ? 4 : 2;
  auto al = [](const void* ptr, int64_t ld, size_t e) { return reinterpret_cast<uintptr_t>(ptr) % 16 == 0 && (ld * (int64_t)e) % 16 == 0; };
  if (h % elems != 0 || h / elems > 256 || !al(x, ldx, es) || !al(y, ldy, es) ||
      (dx && (!al(dx, lddx, gs) || !al(dy, lddy, gs) || lddx < h || lddy < h))) {
    set_error("softmax head kernel needs 16-byte aligned rows, h %% %d == 0 and h <= %d", elems, 256 * elems);
    return IA_ERR_UNSUPPORTED;
  }
  if (n == 0) {
    if (train) {
      static const float kNan = __builtin_nanf("");   // static storage: the async copy may outlive this frame
      IA_CUDA_CHECK(cudaMemcpyAsync(loss_out, &kNan, sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
      if (dw) IA_CUDA_CHECK(cudaMemsetAsync(dw, 0, sizeof(float) * 4 * h, (cudaStream_t)stream));
      if (db) IA_CUDA_CHECK(cudaMemsetAsync(db, 0, sizeof(float) * 2, (cudaStream_t)stream));
    }
    return IA_OK;
  }
  HeadParams p{};
  p.x = x; p.y = y; p.ldx = ldx; p.ldy = ldy; p.w = w; p.b = b; p.labels = labels; p.n = n; p.h = (int)h;
  p.logits = logits; p.probs = probs; p.loss_out = loss_out; p.dx = dx; p.dy = dy; p.lddx = lddx; p.lddy = lddy;
  p.grad_scale = (float)((double)grad_scale / (double)n);
  p.loss_scale = 1.0 / (double)n;
  p.workspace = workspace;
  p.upstream = upstream_dev; p.upstream_skip_one = skip_if_one;
  cudaStream_t s = (cudaStream_t)stream;
  // 16-bit rows, training: the tensor-core kernel (softmax_head_mma.cu); IA_HEAD_MMA=0 keeps the CUDA-core kernel (A/B runs)
  static const int use_mma = [] { const char* e = getenv("IA_HEAD_MMA"); return e ? atoi(e) : 1; }();
  if (use_mma && train && softmax_head_mma_eligible(dtype, p)) return launch_softmax_head_mma(dtype, grad_dtype, p, s, dw, db);
  if (dtype == IA_F32) return launch_head<float, float>(p, train, s, dw, db);
  if (dtype == IA_BF16 && grad_dtype == IA_BF16) return launch_head<__nv_bfloat16, __nv_bfloat16>(p, train, s, dw, db);
  if (dtype == IA_BF16) return launch_head<__nv_bfloat16, float>(p, train, s, dw, db);
  if (dtype == IA_F16 && grad_dtype == IA_F16) return launch_head<__half, __half>(p, train, s, dw, db);
  if (dtype == IA_F16) return launch_head<__half, float>(p, train, s, dw, db);
  set_error("unsupported dtype %d", dtype);
  return IA_ERR_UNSUPPORTED;
}

}  // extern "C"
