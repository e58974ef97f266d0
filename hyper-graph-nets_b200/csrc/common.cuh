// Shared host/device helpers for libhgn_b200.so
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hgn_b200.h"

namespace hgn {

// last error string (per host thread), returned by hgn_last_error()
void set_error(const char* fmt, ...);

#define HGN_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::hgn::set_error(__VA_ARGS__);             \
      return HGN_ERR_INVALID_ARGUMENT;           \
    }                                            \
  } while (0)

#define HGN_CUDA_OK(expr)                                                                     \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::hgn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return HGN_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define HGN_LAUNCH_OK(name)                                                                   \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      ::hgn::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));              \
      return HGN_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

// RAII CUDA-event bracket around one kernel launch (active only after hgn_profile_enable(1))
struct KernelTimer {
  KernelTimer(const char* name, cudaStream_t st);
  ~KernelTimer();
  const char* name_;
  cudaStream_t st_;
  cudaEvent_t start_;
  bool on_;
};
#define HGN_TIMED(name, st) ::hgn::KernelTimer _hgn_timer_##__LINE__(name, st)

constexpr int kD = 128;   // latent width the MLP tile kernels are specialised for (flag.py:57)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// 4 consecutive feature values <-> registers, for fp32 (16 B) and bf16 (8 B) rows
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Sum over `parts` values `stride` apart in their original order (p = 0, 1, 2, ...: bitwise the sequential sum), with the loads of
// eight parts in flight at a time -- the dependent load-add chain made these reductions latency-bound (64 us for 148 parts).
__device__ __forceinline__ float ordered_sum(const float* __restrict__ src, int parts, int64_t stride) {
  float s = 0.f;
  int p = 0;
  for (; p + 8 <= parts; p += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(src + int64_t(p + k) * stride);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
  }
  for (; p < parts; ++p) s += __ldg(src + int64_t(p) * stride);
  return s;
}

}  // namespace hgn
