// common.cuh -- shared host/device helpers for the vtc_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstring>

#include "../../include/vtc_b200.h"

namespace vtc {

extern std::atomic<uint64_t> g_launch_count;
// opt-in launch trace (vtc_trace_begin / vtc_trace_end): one CUDA event after every launch
extern std::atomic<int> g_trace_on;
void trace_mark(const char* file, int line);

inline int cuda_err(cudaError_t e) { return e == cudaSuccess ? VTC_OK : VTC_ERR_CUDA_BASE - (int)e; }

// Every kernel launch goes through this so that launch errors surface and launches are counted.
#define VTC_LAUNCH_CHECK()                                   \
  do {                                                       \
    ::vtc::g_launch_count.fetch_add(1);                      \
    if (::vtc::g_trace_on.load(std::memory_order_relaxed))   \
      ::vtc::trace_mark(__FILE__, __LINE__);                 \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return ::vtc::cuda_err(e__);     \
  } while (0)

#define VTC_RETURN_IF_ERROR(expr) \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != VTC_OK) return rc__; \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- programmatic dependent launch (PDL).  The paths of this library are chains of short kernels
// on one stream (CAM forward: 16, a chunked evaluation: 4 per call): launched back to back, each
// pays its launch latency and its own set-up (barrier init, tensor-memory allocation, descriptor
// prefetch) after the previous kernel has drained.  With PDL a kernel is scheduled as soon as every
// CTA of its predecessor has STARTED (griddep_launch at the top of each kernel), does its set-up,
// and blocks in griddep_wait until the predecessor has completed and flushed its writes.
// Rules kept by every kernel launched through launch_pdl: (1) all threads execute griddep_wait()
// before the first access to global memory and before any early return -- so a grid never completes
// before its predecessor, which makes the ordering transitive along the chain; (2) nothing but
// shared-memory / tensor-memory set-up and kernel-parameter reads happens above it.
// VTC_PDL=0 launches everything fully serialised (profiling / debugging).
__device__ __forceinline__ void griddep_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void griddep_launch() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled();  // api.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}
template <typename T>
__host__ __device__ inline T round_up(T a, T b) {
  return ceil_div(a, b) * b;
}

// Bump allocator over the caller's workspace (256-byte aligned slices).
struct Workspace {
  char* base;
  size_t size;
  size_t used;
  Workspace(void* p, size_t n) : base((char*)p), size(n), used(0) {
    // align the base itself
    size_t mis = ((uintptr_t)base) & 255;
    if (mis) used = 256 - mis;
  }
  template <typename T>
  T* take(size_t count) {
    size_t bytes = round_up<size_t>(count * sizeof(T), 256);
    if (base == nullptr || used + bytes > size) {
      used += bytes;  // keep counting so that the caller can size the workspace
      return nullptr;
    }
    T* p = (T*)(base + used);
    used += bytes;
    return p;
  }
  bool ok() const { return base != nullptr && used <= size; }
};

// element loads widened to double (exact for both storage types)
__device__ __forceinline__ double to_f64(float v) { return (double)v; }
__device__ __forceinline__ double to_f64(__nv_bfloat16 v) { return (double)__bfloat162float(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// 8 consecutive elements widened to double with 128-bit loads (p must be 16-byte aligned).
__device__ __forceinline__ void load8_f64(const float* p, double (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
__device__ __forceinline__ void load8_f64(const __nv_bfloat16* p, double (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = (double)__uint_as_float(w[i] << 16);
    v[2 * i + 1] = (double)__uint_as_float(w[i] & 0xffff0000u);
  }
}
template <typename T>
__device__ __forceinline__ bool aligned16(const T* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}
// 8 consecutive elements as loaded (16-byte aligned p); widened later, so that several loads can be
// in flight before the first dependent FMA
template <typename T>
struct Raw8;
template <>
struct Raw8<float> {
  float4 a, b;
};
template <>
struct Raw8<__nv_bfloat16> {
  uint4 u;
};
__device__ __forceinline__ Raw8<float> load8_raw(const float* p) {
  Raw8<float> r;
  r.a = __ldg(reinterpret_cast<const float4*>(p));
  r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  return r;
}
__device__ __forceinline__ Raw8<__nv_bfloat16> load8_raw(const __nv_bfloat16* p) {
  Raw8<__nv_bfloat16> r;
  r.u = __ldg(reinterpret_cast<const uint4*>(p));
  return r;
}
__device__ __forceinline__ void widen8(const Raw8<float>& r, double (&v)[8]) {
  v[0] = r.a.x, v[1] = r.a.y, v[2] = r.a.z, v[3] = r.a.w;
  v[4] = r.b.x, v[5] = r.b.y, v[6] = r.b.z, v[7] = r.b.w;
}
__device__ __forceinline__ void widen8(const Raw8<__nv_bfloat16>& r, double (&v)[8]) {
  const uint32_t w[4] = {r.u.x, r.u.y, r.u.z, r.u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = (double)__uint_as_float(w[i] << 16);
    v[2 * i + 1] = (double)__uint_as_float(w[i] & 0xffff0000u);
  }
}

// fp64-sequential dot product: acc = fma(q[k], x[k], acc) for k = 0..D-1, in that order.  The
// FMA chain is serial by definition; the loads are not, so the main loop keeps 128 bytes per operand
// in flight (these kernels are latency-bound: one thread walks two rows).
template <typename T>
__device__ __forceinline__ double dot_seq64(const T* __restrict__ q, const T* __restrict__ x, int D) {
  double acc = 0.0;
  int k = 0;
  if (aligned16(q) && aligned16(x)) {
    constexpr int U = sizeof(T) == 2 ? 8 : 4;  // groups of 8 elements per step: 128 bytes per operand
    for (; k + 8 * U <= D; k += 8 * U) {
      Raw8<T> ra[U], rb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ra[u] = load8_raw(q + k + 8 * u);
        rb[u] = load8_raw(x + k + 8 * u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        double a[8], b[8];
        widen8(ra[u], a);
        widen8(rb[u], b);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fma(a[i], b[i], acc);
      }
    }
    for (; k + 8 <= D; k += 8) {
      double a[8], b[8];
      load8_f64(q + k, a);
      load8_f64(x + k, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fma(a[i], b[i], acc);
    }
  }
  for (; k < D; ++k) acc = fma(to_f64(q[k]), to_f64(x[k]), acc);
  return acc;
}
template <typename T>
__device__ __forceinline__ double sq_seq64(const T* __restrict__ x, int D) {
  double acc = 0.0;
  int k = 0;
  if (aligned16(x)) {
    constexpr int U = sizeof(T) == 2 ? 4 : 2;
    for (; k + 8 * U <= D; k += 8 * U) {
      Raw8<T> ra[U];
#pragma unroll
      for (int u = 0; u < U; ++u) ra[u] = load8_raw(x + k + 8 * u);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        double a[8];
        widen8(ra[u], a);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fma(a[i], a[i], acc);
      }
    }
    for (; k + 8 <= D; k += 8) {
      double a[8];
      load8_f64(x + k, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fma(a[i], a[i], acc);
    }
  }
  for (; k < D; ++k) {
    const double v = to_f64(x[k]);
    acc = fma(v, v, acc);
  }
  return acc;
}

// ||x||^2 in fp64 by a whole warp: lane-strided partial sums, then a butterfly -- a fixed order, so
// deterministic, but NOT the fp64-sequential canonical value (differs in the last bits): for guard
// bands and reported distances, never for a comparison that decides a rank.  All lanes return it.
template <typename T>
__device__ __forceinline__ double warp_sq64(const T* __restrict__ x, int D) {
  double acc = 0.0;
  for (int k = threadIdx.x & 31; k < D; k += 32) {
    const double v = to_f64(x[k]);
    acc = fma(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace vtc
