// common.cuh -- shared host/device helpers for the vtc_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/vtc_b200.h"

namespace vtc {

extern std::atomic<uint64_t> g_launch_count;

inline int cuda_err(cudaError_t e) { return e == cudaSuccess ? VTC_OK : VTC_ERR_CUDA_BASE - (int)e; }

// Every kernel launch goes through this so that launch errors surface and launches are counted.
#define VTC_LAUNCH_CHECK()                                   \
  do {                                                       \
    ::vtc::g_launch_count.fetch_add(1);                      \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return ::vtc::cuda_err(e__);     \
  } while (0)

#define VTC_RETURN_IF_ERROR(expr) \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != VTC_OK) return rc__; \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

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
