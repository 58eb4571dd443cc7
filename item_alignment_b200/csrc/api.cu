// extern "C" entry points of libia_b200.so for pair scoring / losses (see include/ia_b200.h).
#include <cstring>
#include <mutex>

#include "pair_kernels.cuh"

namespace ia {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64];
  static std::once_flag once;
  std::call_once(once, [] { memset(cached, 0, sizeof(cached)); });
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev % kMaxDevices;
}

static size_t dtype_size(int dtype) { return dtype == IA_F32 ? 4 : 2; }

static bool aligned16(const void* p, int64_t ld_elems, size_t esize) {
  return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && ((ld_elems * (int64_t)esize) % 16 == 0);
}

static int dispatch_pair(int mode, bool cosloss, int measure, int dtype, int grad_dtype, const PairParams& p,
                         cudaStream_t stream) {
  const size_t es = dtype_size(dtype), gs = dtype_size(grad_dtype);
  const int elems = dtype == IA_F32 ? 4 : 8;
  bool vec_ok = (p.d % elems == 0) && aligned16(p.x, p.ldx, es) && aligned16(p.y, p.ldy, es);
  if (p.dx != nullptr) vec_ok = vec_ok && aligned16(p.dx, p.lddx, gs) && aligned16(p.dy, p.lddy, gs);
  if (dtype == IA_F32 && grad_dtype == IA_F32) return launch_pair_f32_f32(mode, cosloss, measure, p, vec_ok, stream);
  if (dtype == IA_BF16 && grad_dtype == IA_BF16) return launch_pair_bf16_bf16(mode, cosloss, measure, p, vec_ok, stream);
  if (dtype == IA_BF16 && grad_dtype == IA_F32) return launch_pair_bf16_f32(mode, cosloss, measure, p, vec_ok, stream);
  if (dtype == IA_F16 && grad_dtype == IA_F16) return launch_pair_f16_f16(mode, cosloss, measure, p, vec_ok, stream);
  if (dtype == IA_F16 && grad_dtype == IA_F32) return launch_pair_f16_f32(mode, cosloss, measure, p, vec_ok, stream);
  set_error("unsupported dtype combination: inputs %d, gradients %d", dtype, grad_dtype);
  return IA_ERR_UNSUPPORTED;
}

static int check_common(int measure, int dtype, const void* x, const void* y, int64_t n, int64_t d, int64_t ldx,
                        int64_t ldy) {
  if (measure < IA_INNER || measure > IA_L2) {
    set_error("Unsupported similarty measure: %d", measure);  // wording of reference base.py:64
    return IA_ERR_INVALID;
  }
  if (dtype < IA_F32 || dtype > IA_F16) {
    set_error("unsupported dtype %d", dtype);
    return IA_ERR_UNSUPPORTED;
  }
  if (n < 0 || d <= 0 || d > (1 << 24) || ldx < d || ldy < d) {
    set_error("bad shape: n=%lld d=%lld ldx=%lld ldy=%lld", (long long)n, (long long)d, (long long)ldx, (long long)ldy);
    return IA_ERR_INVALID;
  }
  if (n > 0 && (x == nullptr || y == nullptr)) {
    set_error("null input pointer");
    return IA_ERR_INVALID;
  }
  return IA_OK;
}

// ------------------------------------------------------------------ small elementwise kernels
__global__ void __launch_bounds__(256) score_loss_kernel(int loss, float margin, int reduction, const float* sim,
                                                         const int64_t* target, int64_t n, float* loss_out,
                                                         float* gsim, float grad_scale, double loss_scale,
                                                         void* workspace) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = target[i];
    const int label = (loss == IA_LOSS_BCE) ? (t != 0) : (t > 0);
    float g;
    const float li = scalar_loss(loss, sim[i], label, margin, g);
    if (gsim) gsim[i] = g * grad_scale;
    if (reduction == IA_RED_NONE) loss_out[i] = li;
    else acc += li;
  }
  if (reduction != IA_RED_NONE) {
    __shared__ double warp_acc[8];
    double v = (double)acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_acc[threadIdx.x >> 5] = v;
    __syncthreads();
    double blk = 0.0;
    if (threadIdx.x == 0)
      for (int w = 0; w < 8; ++w) blk += warp_acc[w];
    grid_sum_finish(blk, workspace, loss_out, loss_scale);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scale_inplace_kernel(T* a, T* b, int64_t count, const float* g) {
  const float s = __ldg(g);
  if (s == 1.0f) return;  // plain loss.backward(): nothing to do, no traffic
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    a[i] = from_float<T>(to_float<T>(a[i]) * s);
    if (b) b[i] = from_float<T>(to_float<T>(b[i]) * s);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) row_inv_norm_kernel(const T* x, int64_t n, int d, int64_t ldx, float eps,
                                                           float* out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * 8;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += warps_total) {
    const T* xr = x + row * ldx;
    float acc = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float v = to_float<T>(xr[j]);
      acc = fmaf(v, v, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row] = 1.0f / fmaxf(sqrtf(acc), eps);
  }
}

// Confusion counts of `probs >= thr` against {0,1} labels for up to 32 thresholds at once (the sweep of
// reference finetune_text.py:576-580).  counts: [T][4] = tp, fp, fn, tn (u64, integer atomics: deterministic).
__global__ void __launch_bounds__(256) threshold_sweep_kernel(const float* __restrict__ probs, const int64_t* __restrict__ labels,
                                                              int64_t n, const double* __restrict__ thr, int nthr,
                                                              unsigned long long* __restrict__ counts) {
  __shared__ unsigned long long sc[32][4];
  for (int i = threadIdx.x; i < 32 * 4; i += blockDim.x) (&sc[0][0])[i] = 0ull;
  __syncthreads();
  double t[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) t[k] = k < nthr ? thr[k] : 0.0;
  unsigned tp[32], pp[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) { tp[k] = 0; pp[k] = 0; }
  unsigned pos = 0, seen = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double p = (double)probs[i];
    const unsigned l = labels[i] != 0;
    pos += l; ++seen;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const unsigned pred = (k < nthr) && (p >= t[k]);
      pp[k] += pred;
      tp[k] += pred & l;
    }
  }
  for (int k = 0; k < nthr; ++k) {
    const unsigned a = __reduce_add_sync(0xffffffffu, tp[k]);
    const unsigned b = __reduce_add_sync(0xffffffffu, pp[k]);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sc[k][0], (unsigned long long)a); atomicAdd(&sc[k][1], (unsigned long long)b); }
  }
  const unsigned ps = __reduce_add_sync(0xffffffffu, pos), ss = __reduce_add_sync(0xffffffffu, seen);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&sc[0][2], (unsigned long long)ps); atomicAdd(&sc[0][3], (unsigned long long)ss); }
  __syncthreads();
  if (threadIdx.x < nthr) {
    const int k = threadIdx.x;
    const unsigned long long tpk = sc[k][0], ppk = sc[k][1], posb = sc[0][2], seenb = sc[0][3];
    atomicAdd(&counts[k * 4 + 0], tpk);                       // tp
    atomicAdd(&counts[k * 4 + 1], ppk - tpk);                 // fp
    atomicAdd(&counts[k * 4 + 2], posb - tpk);                // fn
    atomicAdd(&counts[k * 4 + 3], (seenb - posb) - (ppk - tpk));   // tn
  }
}

}  // namespace ia

using namespace ia;

extern "C" {

const char* ia_version(void) { return "ia_b200 0.1.0 (sm_100a)"; }
const char* ia_last_error(void) { return g_err; }
size_t ia_workspace_bytes(void) { return kWorkspaceBytes; }
int64_t ia_launch_count(void) { return g_launches.load(); }

int ia_pair_score_fwd(int measure, int dtype, const void* x, const void* y, int64_t n, int64_t d, int64_t ldx,
                      int64_t ldy, float* sim, float* probs, double threshold, uint8_t* labels_out,
                      ia_stream_t stream) {
  int rc = check_common(measure, dtype, x, y, n, d, ldx, ldy);
  if (rc != IA_OK) return rc;
  if (n == 0) return IA_OK;
  PairParams p{};
  p.x = x; p.y = y; p.ldx = ldx; p.ldy = ldy; p.n = n; p.d = (int)d;
  p.sim = sim; p.probs = probs; p.labels_out = labels_out; p.threshold = threshold;
  return dispatch_pair(kModeFwd, false, measure, dtype, dtype, p, (cudaStream_t)stream);
}

int ia_pair_score_loss_fwd_bwd(int measure, int loss, float margin, int reduction, int dtype, int grad_dtype,
                               const void* x, const void* y, int64_t ldx, int64_t ldy, const int64_t* labels,
                               int64_t n, int64_t d, float* sim, float* probs, float* loss_out, void* dx, void* dy,
                               int64_t lddx, int64_t lddy, float grad_scale, const float* upstream_dev,
                               int skip_if_one, void* workspace, size_t workspace_bytes, ia_stream_t stream) {
  return ia_pair_score_loss_act_bwd(measure, loss, margin, reduction, dtype, grad_dtype, x, y, ldx, ldy, labels, n, d, sim, probs, loss_out,
                                    dx, dy, lddx, lddy, grad_scale, upstream_dev, skip_if_one, 0.0f, workspace, workspace_bytes, stream);
}

int ia_pair_score_loss_act_bwd(int measure, int loss, float margin, int reduction, int dtype, int grad_dtype,
                               const void* x, const void* y, int64_t ldx, int64_t ldy, const int64_t* labels,
                               int64_t n, int64_t d, float* sim, float* probs, float* loss_out, void* dx, void* dy,
                               int64_t lddx, int64_t lddy, float grad_scale, const float* upstream_dev,
                               int skip_if_one, float act_bwd_scale, void* workspace, size_t workspace_bytes,
                               ia_stream_t stream) {
  if (!(act_bwd_scale >= 0.f)) { set_error("act_bwd_scale must be 0 (off) or the dropout keep scale 1/(1-p) >= 1"); return IA_ERR_INVALID; }
  int rc = check_common(measure, dtype, x, y, n, d, ldx, ldy);
  if (rc != IA_OK) return rc;
  if (loss < IA_LOSS_BCE || loss > IA_LOSS_COSINE) { set_error("unsupported loss_type %d", loss); return IA_ERR_INVALID; }
  if (reduction < IA_RED_NONE || reduction > IA_RED_SUM) { set_error("bad reduction %d", reduction); return IA_ERR_INVALID; }
  if (grad_dtype != dtype && grad_dtype != IA_F32) { set_error("grad_dtype must equal dtype or be fp32"); return IA_ERR_UNSUPPORTED; }
  if ((dx == nullptr) != (dy == nullptr)) { set_error("dx and dy must both be given or both be NULL"); return IA_ERR_INVALID; }
  if (dx != nullptr && (lddx < d || lddy < d)) { set_error("bad gradient leading dimension"); return IA_ERR_INVALID; }
  if (n > 0 && labels == nullptr) { set_error("labels must not be NULL"); return IA_ERR_INVALID; }
  if (workspace == nullptr || workspace_bytes < kWorkspaceBytes) {
    if (reduction != IA_RED_NONE || loss_out == nullptr) { set_error("workspace too small: need %zu bytes", kWorkspaceBytes); return IA_ERR_WORKSPACE; }
  }
  if (loss_out == nullptr) {
    // gradient recomputation (autograd backward): the loss value is not wanted, it lands in the workspace header
    if (upstream_dev == nullptr || reduction == IA_RED_NONE) { set_error("loss_out may only be NULL with upstream_dev and mean / sum reduction"); return IA_ERR_INVALID; }
    loss_out = reinterpret_cast<float*>(static_cast<char*>(workspace) + kWorkspaceScratchOffset);
  }
  if (n == 0) {
    // torch: mean over an empty batch is nan, sum is 0
    if (reduction != IA_RED_NONE) {
      static const float kEmpty[2] = {__builtin_nanf(""), 0.f};   // static storage: the async copy may outlive this frame
      IA_CUDA_CHECK(cudaMemcpyAsync(loss_out, &kEmpty[reduction == IA_RED_MEAN ? 0 : 1], sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    }
    return IA_OK;
  }
  PairParams p{};
  p.x = x; p.y = y; p.ldx = ldx; p.ldy = ldy; p.labels = labels; p.n = n; p.d = (int)d;
  p.sim = sim; p.probs = probs; p.loss_out = loss_out; p.dx = dx; p.dy = dy; p.lddx = lddx; p.lddy = lddy;
  p.loss = loss; p.margin = margin; p.reduction = reduction;
  p.loss_scale = reduction == IA_RED_MEAN ? 1.0 / (double)n : 1.0;
  p.grad_scale = reduction == IA_RED_MEAN ? (float)((double)grad_scale / (double)n) : grad_scale;
  p.workspace = workspace;
  p.upstream = upstream_dev; p.upstream_skip_one = skip_if_one; p.act_bwd = act_bwd_scale;
  return dispatch_pair(kModeFused, loss == IA_LOSS_COSINE, measure, dtype, grad_dtype, p, (cudaStream_t)stream);
}

int ia_pair_score_bwd(int measure, int dtype, int grad_dtype, const void* x, const void* y, int64_t ldx, int64_t ldy,
                      const float* gsim, int64_t n, int64_t d, void* dx, void* dy, int64_t lddx, int64_t lddy,
                      ia_stream_t stream) {
  int rc = check_common(measure, dtype, x, y, n, d, ldx, ldy);
  if (rc != IA_OK) return rc;
  if (grad_dtype != dtype && grad_dtype != IA_F32) { set_error("grad_dtype must equal dtype or be fp32"); return IA_ERR_UNSUPPORTED; }
  if (dx == nullptr || dy == nullptr || gsim == nullptr || lddx < d || lddy < d) { set_error("bad gradient arguments"); return IA_ERR_INVALID; }
  if (n == 0) return IA_OK;
  PairParams p{};
  p.x = x; p.y = y; p.ldx = ldx; p.ldy = ldy; p.gsim = gsim; p.n = n; p.d = (int)d;
  p.dx = dx; p.dy = dy; p.lddx = lddx; p.lddy = lddy;
  return dispatch_pair(kModeBwd, false, measure, dtype, grad_dtype, p, (cudaStream_t)stream);
}

int ia_pair_score_gather_fwd(int measure, int dtype, const void* ex, const void* ey, int64_t rows_x, int64_t rows_y,
                             int64_t ldx, int64_t ldy, const int64_t* xi, const int64_t* yi, int64_t n, int64_t d,
                             float* sim, float* probs, double threshold, uint8_t* labels_out, ia_stream_t stream) {
  int rc = check_common(measure, dtype, ex, ey, n, d, ldx, ldy);
  if (rc != IA_OK) return rc;
  if (n > 0 && (xi == nullptr || yi == nullptr || rows_x <= 0 || rows_y <= 0)) { set_error("gather needs both index arrays and the row counts"); return IA_ERR_INVALID; }
  if (n == 0) return IA_OK;
  PairParams p{};
  p.x = ex; p.y = ey; p.ldx = ldx; p.ldy = ldy; p.xi = xi; p.yi = yi; p.n = n; p.d = (int)d;
  p.rows_x = rows_x; p.rows_y = rows_y;
  p.sim = sim; p.probs = probs; p.labels_out = labels_out; p.threshold = threshold;
  return dispatch_pair(kModeFwd, false, measure, dtype, dtype, p, (cudaStream_t)stream);
}

int ia_pair_score_gather_loss_fwd_bwd(int measure, int loss, float margin, int reduction, int dtype, int grad_dtype,
                                      const void* ex, const void* ey, int64_t rows_x, int64_t rows_y, int64_t ldx,
                                      int64_t ldy, const int64_t* xi, const int64_t* yi, const int64_t* labels, int64_t n,
                                      int64_t d, float* sim,
                                      float* probs, float* loss_out, void* dx, void* dy, int64_t lddx, int64_t lddy,
                                      float grad_scale, void* workspace, size_t workspace_bytes, ia_stream_t stream) {
  int rc = check_common(measure, dtype, ex, ey, n, d, ldx, ldy);
  if (rc != IA_OK) return rc;
  if (loss < IA_LOSS_BCE || loss > IA_LOSS_COSINE) { set_error("unsupported loss_type %d", loss); return IA_ERR_INVALID; }
  if (reduction != IA_RED_MEAN && reduction != IA_RED_SUM) { set_error("gather entry point supports mean / sum reduction"); return IA_ERR_INVALID; }
  if (grad_dtype != dtype && grad_dtype != IA_F32) { set_error("grad_dtype must equal dtype or be fp32"); return IA_ERR_UNSUPPORTED; }
  if ((dx == nullptr) != (dy == nullptr)) { set_error("dx and dy must both be given or both be NULL"); return IA_ERR_INVALID; }
  if (loss_out == nullptr || (n > 0 && (labels == nullptr || xi == nullptr || yi == nullptr || rows_x <= 0 || rows_y <= 0))) {
    set_error("loss_out / labels / indices must not be NULL and the row counts must be positive");
    return IA_ERR_INVALID;
  }
  if (workspace == nullptr || workspace_bytes < kWorkspaceBytes) { set_error("workspace too small: need %zu bytes", kWorkspaceBytes); return IA_ERR_WORKSPACE; }
  if (n == 0) {
    static const float kEmpty[2] = {__builtin_nanf(""), 0.f};
    IA_CUDA_CHECK(cudaMemcpyAsync(loss_out, &kEmpty[reduction == IA_RED_MEAN ? 0 : 1], sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return IA_OK;
  }
  PairParams p{};
  p.x = ex; p.y = ey; p.ldx = ldx; p.ldy = ldy; p.xi = xi; p.yi = yi; p.labels = labels; p.n = n; p.d = (int)d;
  p.rows_x = rows_x; p.rows_y = rows_y;
  p.sim = sim; p.probs = probs; p.loss_out = loss_out; p.dx = dx; p.dy = dy; p.lddx = lddx; p.lddy = lddy;
  p.loss = loss; p.margin = margin; p.reduction = reduction;
  p.loss_scale = reduction == IA_RED_MEAN ? 1.0 / (double)n : 1.0;
  p.grad_scale = reduction == IA_RED_MEAN ? (float)((double)grad_scale / (double)n) : grad_scale;
  p.workspace = workspace;
  return dispatch_pair(kModeFused, loss == IA_LOSS_COSINE, measure, dtype, grad_dtype, p, (cudaStream_t)stream);
}

int ia_threshold_sweep(const float* probs, const int64_t* labels, int64_t n, const double* thresholds, int nthr,
                       uint64_t* counts, ia_stream_t stream) {
  if (n < 0 || nthr < 1 || nthr > 32 || counts == nullptr || thresholds == nullptr || (n > 0 && (probs == nullptr || labels == nullptr))) {
    set_error("bad arguments (1..32 thresholds)");
    return IA_ERR_INVALID;
  }
  IA_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * 4 * nthr, (cudaStream_t)stream));
  if (n == 0) return IA_OK;
  int64_t want = (n + 255) / 256;
  int grid = (int)(want < 4 * sm_count() ? want : 4 * sm_count());
  threshold_sweep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(probs, labels, n, thresholds, nthr,
                                                                 reinterpret_cast<unsigned long long*>(counts));
  IA_LAUNCH_CHECK();
  return IA_OK;
}

int ia_score_loss_fwd_bwd(int loss, float margin, int reduction, const float* sim, const int64_t* target, int64_t n,
                          float* loss_out, float* gsim, float grad_scale, void* workspace, size_t workspace_bytes,
                          ia_stream_t stream) {
  if (loss < IA_LOSS_BCE || loss > IA_LOSS_EUCLIDEAN) { set_error("unsupported loss_type %d for a score vector", loss); return IA_ERR_INVALID; }
  if (reduction < IA_RED_NONE || reduction > IA_RED_SUM) { set_error("bad reduction %d", reduction); return IA_ERR_INVALID; }
  if (n < 0 || loss_out == nullptr || (n > 0 && (sim == nullptr || target == nullptr))) { set_error("bad arguments"); return IA_ERR_INVALID; }
  if (reduction != IA_RED_NONE && (workspace == nullptr || workspace_bytes < kWorkspaceBytes)) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return IA_ERR_WORKSPACE;
  }
  if (n == 0) {
    if (reduction != IA_RED_NONE) {
      static const float kEmpty[2] = {__builtin_nanf(""), 0.f};
      IA_CUDA_CHECK(cudaMemcpyAsync(loss_out, &kEmpty[reduction == IA_RED_MEAN ? 0 : 1], sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    }
    return IA_OK;
  }
  int64_t want = (n + 255) / 256;
  int grid = (int)(want < 2 * sm_count() ? want : 2 * sm_count());
  const double ls = reduction == IA_RED_MEAN ? 1.0 / (double)n : 1.0;
  const float gsc = reduction == IA_RED_MEAN ? (float)((double)grad_scale / (double)n) : grad_scale;
  score_loss_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(loss, margin, reduction, sim, target, n, loss_out, gsim, gsc,
                                                            ls, workspace);
  IA_LAUNCH_CHECK();
  return IA_OK;
}

int ia_scale_inplace(int dtype, void* a, void* b, int64_t count, const float* g, ia_stream_t stream) {
  if (a == nullptr || g == nullptr || count < 0) { set_error("bad arguments"); return IA_ERR_INVALID; }
  if (count == 0) return IA_OK;
  int64_t want = (count + 255) / 256;
  int grid = (int)(want < 8 * sm_count() ? want : 8 * sm_count());
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == IA_F32) scale_inplace_kernel<float><<<grid, 256, 0, s>>>((float*)a, (float*)b, count, g);
  else if (dtype == IA_BF16) scale_inplace_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((__nv_bfloat16*)a, (__nv_bfloat16*)b, count, g);
  else if (dtype == IA_F16) scale_inplace_kernel<__half><<<grid, 256, 0, s>>>((__half*)a, (__half*)b, count, g);
  else { set_error("unsupported dtype %d", dtype); return IA_ERR_UNSUPPORTED; }
  IA_LAUNCH_CHECK();
  return IA_OK;
}

int ia_row_inv_norm(int dtype, const void* x, int64_t n, int64_t d, int64_t ldx, float eps, float* out,
                    ia_stream_t stream) {
  if (n < 0 || d <= 0 || ldx < d || out == nullptr || (n > 0 && x == nullptr)) { set_error("bad arguments"); return IA_ERR_INVALID; }
  if (n == 0) return IA_OK;
  int64_t want = (n + 7) / 8;
  int grid = (int)(want < 8 * sm_count() ? want : 8 * sm_count());
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == IA_F32) row_inv_norm_kernel<float><<<grid, 256, 0, s>>>((const float*)x, n, (int)d, ldx, eps, out);
  else if (dtype == IA_BF16) row_inv_norm_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, n, (int)d, ldx, eps, out);
  else if (dtype == IA_F16) row_inv_norm_kernel<__half><<<grid, 256, 0, s>>>((const __half*)x, n, (int)d, ldx, eps, out);
  else { set_error("unsupported dtype %d", dtype); return IA_ERR_UNSUPPORTED; }
  IA_LAUNCH_CHECK();
  return IA_OK;
}

}  // extern "C"
