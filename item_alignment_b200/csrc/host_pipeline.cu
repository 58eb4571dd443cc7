// Host-buffer entry points (include/ia_b200.h, "host-buffer entry points").
//
// What a CPU-side caller of the reference would bind: inputs and outputs live in host memory.  The
// batch is cut into row chunks that travel through kStreams independent CUDA streams, each doing
// [H2D chunk -> kernel -> D2H results] on its own device staging buffers, so the PCIe upload of chunk
// i+1 overlaps the kernel and the download of chunk i (PCIe is full duplex).  Pinned host memory gives
// full link speed; pageable memory works but is staged by the driver.
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace ia {

constexpr int kStreams = 3;

struct Lane {
  cudaStream_t stream = nullptr;
  void* x = nullptr; void* y = nullptr; void* dx = nullptr; void* dy = nullptr;
  float* sim = nullptr; float* probs = nullptr; uint8_t* lab = nullptr; int64_t* labels = nullptr;
  float* loss = nullptr; void* ws = nullptr;
  size_t in_bytes = 0, grad_bytes = 0, rows = 0;
};

struct HostCtx {
  std::mutex mu;
  Lane lanes[kStreams];
  float* pinned_loss = nullptr;   // per-chunk loss partials land here (pinned: the D2H must not block the host)
};
constexpr int kMaxChunks = 4096 * kStreams;

static HostCtx g_ctx[16];

static int ensure(Lane& l, size_t in_bytes, size_t grad_bytes, size_t rows) {
  if (l.stream == nullptr) IA_CUDA_CHECK(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
  if (l.ws == nullptr) {
    IA_CUDA_CHECK(cudaMalloc(&l.ws, kWorkspaceBytes));
    IA_CUDA_CHECK(cudaMemset(l.ws, 0, kWorkspaceBytes));
    IA_CUDA_CHECK(cudaMalloc(&l.loss, sizeof(float) * 4096));
  }
  if (l.in_bytes < in_bytes) {
    if (l.x) { cudaFree(l.x); cudaFree(l.y); }
    l.in_bytes = 0;
    IA_CUDA_CHECK(cudaMalloc(&l.x, in_bytes));
    IA_CUDA_CHECK(cudaMalloc(&l.y, in_bytes));
    l.in_bytes = in_bytes;
  }
  if (l.grad_bytes < grad_bytes) {
    if (l.dx) { cudaFree(l.dx); cudaFree(l.dy); }
    l.grad_bytes = 0;
    IA_CUDA_CHECK(cudaMalloc(&l.dx, grad_bytes));
    IA_CUDA_CHECK(cudaMalloc(&l.dy, grad_bytes));
    l.grad_bytes = grad_bytes;
  }
  if (l.rows < rows) {
    if (l.sim) { cudaFree(l.sim); cudaFree(l.probs); cudaFree(l.lab); cudaFree(l.labels); }
    l.rows = 0;
    IA_CUDA_CHECK(cudaMalloc(&l.sim, sizeof(float) * rows));
    IA_CUDA_CHECK(cudaMalloc(&l.probs, sizeof(float) * rows));
    IA_CUDA_CHECK(cudaMalloc(&l.lab, rows));
    IA_CUDA_CHECK(cudaMalloc(&l.labels, sizeof(int64_t) * rows));
    l.rows = rows;
  }
  return IA_OK;
}

// The entry points run on `device` and hand the caller's current device back on EVERY exit path; on an error exit the lanes
// are drained first, so no asynchronous copy is still writing into the caller's buffers after the call has returned.
struct DeviceScope {
  int prev = -1;
  HostCtx* ctx = nullptr;
  bool ok = true;
  explicit DeviceScope(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceScope() {
    if (ctx != nullptr)
      for (int i = 0; i < kStreams; ++i)
        if (ctx->lanes[i].stream != nullptr) cudaStreamSynchronize(ctx->lanes[i].stream);
    if (prev >= 0) cudaSetDevice(prev);
  }
};

static int64_t chunk_rows_for(int64_t n, int64_t row_bytes) {
  // bytes of x per chunk.  Small chunks shorten the pipeline's fill and drain (the first upload and the last download are
  // not overlapped with anything); large chunks amortise the per-chunk API calls.  IA_HOST_CHUNK_MB overrides (A/B runs).
  static const int64_t chunk_bytes = [] { const char* e = getenv("IA_HOST_CHUNK_MB"); const int v = e ? atoi(e) : 0; return (int64_t)(v > 0 ? v : 16) << 20; }();
  int64_t rows = chunk_bytes / (row_bytes > 0 ? row_bytes : 1);
  if (rows < 1024) rows = 1024;
  if (rows > n) rows = n;
  return rows;
}

}  // namespace ia

using namespace ia;

extern "C" {

// Page-locked host buffers for the host entry points.  write_combined = 1: for buffers the CPU only WRITES and the device
// reads (inputs): the device's DMA reads do not snoop the CPU caches; reading such memory from the CPU is slow.
int ia_host_alloc(void** ptr, size_t bytes, int write_combined) {
  if (ptr == nullptr || bytes == 0) { set_error("bad arguments"); return IA_ERR_INVALID; }
  IA_CUDA_CHECK(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
  return IA_OK;
}
int ia_host_free(void* ptr) {
  if (ptr != nullptr) IA_CUDA_CHECK(cudaFreeHost(ptr));
  return IA_OK;
}

int ia_pair_score_host(int measure, int dtype, const void* x, const void* y, int64_t n, int64_t d, float* sim,
                       float* probs, double threshold, uint8_t* labels_out, int device) {
  if (device < 0 || device >= 16) { set_error("bad device %d", device); return IA_ERR_INVALID; }
  if (n < 0 || d <= 0 || (n > 0 && (x == nullptr || y == nullptr))) { set_error("bad arguments"); return IA_ERR_INVALID; }
  if (n == 0) return IA_OK;
  HostCtx& ctx = g_ctx[device];
  std::lock_guard<std::mutex> lock(ctx.mu);
  DeviceScope scope(device);
  if (!scope.ok) { set_error("cudaSetDevice(%d) failed", device); return IA_ERR_CUDA; }
  scope.ctx = &ctx;
  const size_t es = dtype == IA_F32 ? 4 : 2;
  const int64_t row_bytes = d * (int64_t)es;
  const int64_t chunk = chunk_rows_for(n, row_bytes);
  int rc;
  for (int i = 0; i < kStreams; ++i)
    if ((rc = ensure(ctx.lanes[i], (size_t)(chunk * row_bytes), 0, (size_t)chunk)) != IA_OK) return rc;
  int li = 0;
  for (int64_t r0 = 0; r0 < n; r0 += chunk, li = (li + 1) % kStreams) {
    Lane& l = ctx.lanes[li];
    const int64_t rows = (n - r0) < chunk ? (n - r0) : chunk;
    const char* xs = static_cast<const char*>(x) + r0 * row_bytes;
    const char* ys = static_cast<const char*>(y) + r0 * row_bytes;
    IA_CUDA_CHECK(cudaMemcpyAsync(l.x, xs, (size_t)(rows * row_bytes), cudaMemcpyHostToDevice, l.stream));
    IA_CUDA_CHECK(cudaMemcpyAsync(l.y, ys, (size_t)(rows * row_bytes), cudaMemcpyHostToDevice, l.stream));
    rc = ia_pair_score_fwd(measure, dtype, l.x, l.y, rows, d, d, d, l.sim, probs ? l.probs : nullptr, threshold,
                           labels_out ? l.lab : nullptr, l.stream);
    if (rc != IA_OK) return rc;
    if (sim) IA_CUDA_CHECK(cudaMemcpyAsync(sim + r0, l.sim, sizeof(float) * rows, cudaMemcpyDeviceToHost, l.stream));
    if (probs) IA_CUDA_CHECK(cudaMemcpyAsync(probs + r0, l.probs, sizeof(float) * rows, cudaMemcpyDeviceToHost, l.stream));
    if (labels_out) IA_CUDA_CHECK(cudaMemcpyAsync(labels_out + r0, l.lab, (size_t)rows, cudaMemcpyDeviceToHost, l.stream));
  }
  for (int i = 0; i < kStreams; ++i) IA_CUDA_CHECK(cudaStreamSynchronize(ctx.lanes[i].stream));
  return IA_OK;
}

int ia_pair_score_loss_host(int measure, int loss, float margin, int reduction, int dtype, const void* x, const void* y,
                            const int64_t* labels, int64_t n, int64_t d, float* loss_out, void* dx, void* dy,
                            int device) {
  if (device < 0 || device >= 16) { set_error("bad device %d", device); return IA_ERR_INVALID; }
  if (n < 0 || d <= 0 || loss_out == nullptr || (n > 0 && (x == nullptr || y == nullptr || labels == nullptr))) {
    set_error("bad arguments");
    return IA_ERR_INVALID;
  }
  if (reduction != IA_RED_MEAN && reduction != IA_RED_SUM) { set_error("host entry point supports mean / sum reduction"); return IA_ERR_INVALID; }
  if ((dx == nullptr) != (dy == nullptr)) { set_error("dx and dy must both be given or both be NULL"); return IA_ERR_INVALID; }
  if (n == 0) { *loss_out = reduction == IA_RED_MEAN ? __builtin_nanf("") : 0.f; return IA_OK; }
  HostCtx& ctx = g_ctx[device];
  std::lock_guard<std::mutex> lock(ctx.mu);
  DeviceScope scope(device);
  if (!scope.ok) { set_error("cudaSetDevice(%d) failed", device); return IA_ERR_CUDA; }
  scope.ctx = &ctx;
  const size_t es = dtype == IA_F32 ? 4 : 2;
  const int64_t row_bytes = d * (int64_t)es;
  const int64_t chunk = chunk_rows_for(n, row_bytes);
  const int64_t n_chunks = (n + chunk - 1) / chunk;
  if (n_chunks > kMaxChunks) { set_error("batch too large for the host pipeline"); return IA_ERR_UNSUPPORTED; }
  if (ctx.pinned_loss == nullptr) IA_CUDA_CHECK(cudaMallocHost(&ctx.pinned_loss, sizeof(float) * kMaxChunks));
  int rc;
  for (int i = 0; i < kStreams; ++i)
    if ((rc = ensure(ctx.lanes[i], (size_t)(chunk * row_bytes), dx ? (size_t)(chunk * row_bytes) : 0, (size_t)chunk)) != IA_OK) return rc;
  float* partial = ctx.pinned_loss;
  const float gscale = reduction == IA_RED_MEAN ? (float)(1.0 / (double)n) : 1.0f;
  int li = 0;
  int64_t ci = 0;
  for (int64_t r0 = 0; r0 < n; r0 += chunk, li = (li + 1) % kStreams, ++ci) {
    Lane& l = ctx.lanes[li];
    const int64_t rows = (n - r0) < chunk ? (n - r0) : chunk;
    IA_CUDA_CHECK(cudaMemcpyAsync(l.x, static_cast<const char*>(x) + r0 * row_bytes, (size_t)(rows * row_bytes), cudaMemcpyHostToDevice, l.stream));
    IA_CUDA_CHECK(cudaMemcpyAsync(l.y, static_cast<const char*>(y) + r0 * row_bytes, (size_t)(rows * row_bytes), cudaMemcpyHostToDevice, l.stream));
    IA_CUDA_CHECK(cudaMemcpyAsync(l.labels, labels + r0, sizeof(int64_t) * rows, cudaMemcpyHostToDevice, l.stream));
    float* lslot = l.loss + (ci / kStreams) % 4096;
    rc = ia_pair_score_loss_fwd_bwd(measure, loss, margin, IA_RED_SUM, dtype, dtype, l.x, l.y, d, d, l.labels, rows, d,
                                    nullptr, nullptr, lslot, dx ? l.dx : nullptr, dx ? l.dy : nullptr, d, d, gscale,
                                    nullptr, 0, l.ws, kWorkspaceBytes, l.stream);
    if (rc != IA_OK) return rc;
    IA_CUDA_CHECK(cudaMemcpyAsync(&partial[(size_t)ci], lslot, sizeof(float), cudaMemcpyDeviceToHost, l.stream));
    if (dx) {
      IA_CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(dx) + r0 * row_bytes, l.dx, (size_t)(rows * row_bytes), cudaMemcpyDeviceToHost, l.stream));
      IA_CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(dy) + r0 * row_bytes, l.dy, (size_t)(rows * row_bytes), cudaMemcpyDeviceToHost, l.stream));
    }
  }
  for (int i = 0; i < kStreams; ++i) IA_CUDA_CHECK(cudaStreamSynchronize(ctx.lanes[i].stream));
  double total = 0.0;
  for (int64_t i = 0; i < n_chunks; ++i) total += (double)partial[(size_t)i];   // fixed order: deterministic
  *loss_out = (float)(reduction == IA_RED_MEAN ? total / (double)n : total);
  return IA_OK;
}

}  // extern "C"
