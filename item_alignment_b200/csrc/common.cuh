// Shared device/host helpers for the ia_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "../../include/ia_b200.h"

namespace ia {

// ----------------------------------------------------------------------------- host side
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
int sm_count();  // SMs of the current device (cached per device)
constexpr int kMaxDevices = 64;
int device_slot();  // index of the current device in [0, kMaxDevices): per-device caches of function attributes

#define IA_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ia::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                    __LINE__);                                                           \
      return IA_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define IA_LAUNCH_CHECK()                     \
  do {                                        \
    ia::g_launches.fetch_add(1);              \
    IA_CUDA_CHECK(cudaGetLastError());        \
  } while (0)

// Programmatic dependent launch: the grid may be SCHEDULED while its predecessor in the stream is still draining (its CTAs take
// the SM slots the predecessor's CTAs leave), and every kernel launched this way executes pdl_wait() before its first global
// memory access -- that returns when the predecessor has completed and its writes are visible, so the stream's ordering is
// unchanged; what overlaps is launch latency, CTA scheduling and the prologue.  pdl_trigger() at the top of a kernel lets ITS
// successor do the same.  IA_PDL=0 launches the plain way (A/B runs).
inline bool pdl_enabled() {
  static const int on = [] { const char* e = getenv("IA_PDL"); return e ? atoi(e) : 1; }();
  return on != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args... args) {
  if (!pdl_enabled()) {
    kernel<<<grid, block, smem, stream>>>(args...);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#define IA_PDL_LAUNCH_CHECK(expr)             \
  do {                                        \
    ia::g_launches.fetch_add(1);              \
    IA_CUDA_CHECK(expr);                      \
  } while (0)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

template <typename K>
int blocks_per_sm(K kernel, int threads, size_t smem = 0) {
  int b = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, threads, smem) != cudaSuccess || b < 1) b = 1;
  return b;
}

// workspace layout shared by the fused kernels: [counter u32 | pad to 64 B | double partials[]]
constexpr int kMaxPartials = 8192;
constexpr size_t kWorkspaceBytes = 64 + sizeof(double) * kMaxPartials;
constexpr size_t kWorkspaceScratchOffset = 16;   // one float of the header: where an unwanted loss value goes (gradient recomputation)

// ----------------------------------------------------------------------------- device side
constexpr float kCosEps = 1e-8f;      // nn.CosineSimilarity eps (reference base.py:58)
constexpr float kPdistEps = 1e-6f;    // nn.PairwiseDistance eps (reference base.py:60,62)
constexpr float kCosEmbEps = 1e-12f;  // ATen cosine_embedding_loss EPSILON (reference text.py:1401)

template <typename T> struct VecTraits;
template <> struct VecTraits<float> { static constexpr int kElems = 4; };
template <> struct VecTraits<__nv_bfloat16> { static constexpr int kElems = 8; };
template <> struct VecTraits<__half> { static constexpr int kElems = 8; };

// streaming 128-bit global load: read-only path, do not allocate in L1 (each byte is used once)
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_cs(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// unpack one 128-bit vector into kElems floats
template <typename T> __device__ __forceinline__ void unpack(const uint4& v, float* f);
template <> __device__ __forceinline__ void unpack<float>(const uint4& v, float* f) {
  f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
  f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}
template <> __device__ __forceinline__ void unpack<__nv_bfloat16>(const uint4& v, float* f) {
  // bf16 -> fp32 is a 16-bit shift
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack<__half>(const uint4& v, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}

// pack floats back (round to nearest even) and store; NF = number of floats (4 or 8)
template <typename G, int NF> struct Packer;
template <> struct Packer<float, 4> {
  static __device__ __forceinline__ void store(float* dst, const float* f) {
    stg_stream(reinterpret_cast<uint4*>(dst), make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]),
                                                         __float_as_uint(f[2]), __float_as_uint(f[3])));
  }
};
template <> struct Packer<float, 8> {
  static __device__ __forceinline__ void store(float* dst, const float* f) {
    Packer<float, 4>::store(dst, f);
    Packer<float, 4>::store(dst + 4, f + 4);
  }
};
template <> struct Packer<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void store(__nv_bfloat16* dst, const float* f) {
    uint4 v;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(f[0], f[1]); v.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(f[2], f[3]); v.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(f[4], f[5]); v.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(f[6], f[7]); v.w = *reinterpret_cast<uint32_t*>(&t);
    stg_stream(reinterpret_cast<uint4*>(dst), v);
  }
};
template <> struct Packer<__half, 8> {
  static __device__ __forceinline__ void store(__half* dst, const float* f) {
    uint4 v;
    __half2 t;
    t = __floats2half2_rn(f[0], f[1]); v.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2half2_rn(f[2], f[3]); v.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2half2_rn(f[4], f[5]); v.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2half2_rn(f[6], f[7]); v.w = *reinterpret_cast<uint32_t*>(&t);
    stg_stream(reinterpret_cast<uint4*>(dst), v);
  }
};

template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// probability map of VecSimClassificationHead.forward (reference base.py:79-86)
template <int MEASURE> __device__ __forceinline__ float prob_of(float s) {
  if (MEASURE == IA_COSINE) return (s + 1.0f) / 2.0f;
  if (MEASURE == IA_INNER) return 1.0f / (1.0f + expf(-s));
  return expf(-s);
}

// loss on a scalar score (reference text.py:1468-1477 ladder; loss.py:61-68,126-134).
// label in {0,1}; returns loss_i and writes g = d loss_i / d s.
__device__ __forceinline__ float scalar_loss(int loss, float s, int label, float margin, float& g) {
  const float t = label ? 1.0f : -1.0f;
  if (loss == IA_LOSS_BCE) {  // nn.BCEWithLogitsLoss with label.float()
    const float l = (float)label;
    g = 1.0f / (1.0f + expf(-s)) - l;
    return fmaxf(s, 0.0f) - s * l + log1pf(expf(-fabsf(s)));
  }
  if (loss == IA_LOSS_HINGE) {  // max(0, margin - s*t); torch.max splits the sub-gradient at the kink
    const float z = margin - s * t;
    g = z > 0.0f ? -t : (z == 0.0f ? -0.5f * t : 0.0f);
    return fmaxf(0.0f, z);
  }
  // euclidean: s ** t  (s for positives, 1/s for negatives)
  if (label) { g = 1.0f; return s; }
  g = -1.0f / (s * s);
  return 1.0f / s;
}

// Deterministic grid-wide sum of one double per block: block partial -> workspace; the last block to
// arrive (atomic ticket) adds all partials in index order and writes scale*sum.  The ticket is reset so
// the workspace stays zeroed for the next launch.  Call with all threads of the block.
__device__ __forceinline__ void grid_sum_finish(double block_val, void* workspace, float* out, double scale) {
  unsigned* counter = reinterpret_cast<unsigned*>(workspace);
  double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 64);
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = block_val;
    __threadfence();
    const unsigned ticket = atomicAdd(counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    double acc = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) acc += __ldcg(&partials[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) {
      *out = (float)(acc * scale);
      *counter = 0u;
    }
  }
}

}  // namespace ia
