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

// Sum over `parts` values `stride` apart in a FIXED order (bitwise reproducible run to run): four interleaved partial sums
// (parts p = 0, 4, 8 ... | 1, 5, 9 ... | ...) with sixteen loads in flight, combined as (s0 + s1) + (s2 + s3).  One sequential chain
// over 148 per-CTA partials with eight loads in flight was latency-bound (17-19 us per reduction kernel, 1 ms per cfg5 step at every
// GPU count -- the largest fixed cost of the partitioned runs).
__device__ __forceinline__ float ordered_sum(const float* __restrict__ src, int parts, int64_t stride) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int p = 0;
  for (; p + 16 <= parts; p += 16) {
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = __ldg(src + int64_t(p + k) * stride);
#pragma unroll
    for (int k = 0; k < 16; k += 4) { s0 += v[k]; s1 += v[k + 1]; s2 += v[k + 2]; s3 += v[k + 3]; }
  }
  for (; p + 4 <= parts; p += 4) {
    s0 += __ldg(src + int64_t(p) * stride); s1 += __ldg(src + int64_t(p + 1) * stride);
    s2 += __ldg(src + int64_t(p + 2) * stride); s3 += __ldg(src + int64_t(p + 3) * stride);
  }
  if (p < parts) s0 += __ldg(src + int64_t(p) * stride);
  if (p + 1 < parts) s1 += __ldg(src + int64_t(p + 1) * stride);
  if (p + 2 < parts) s2 += __ldg(src + int64_t(p + 2) * stride);
  return (s0 + s1) + (s2 + s3);
}


// The same sum by a whole 256-thread block for 32 adjacent outputs: lane = output, warp g sums parts g, g + 8, g + 16 ... (a warp
// reads 128 contiguous bytes per part), the eight warp sums are combined through shared memory as ((0+1)+(2+3))+((4+5)+(6+7)).
// Fixed order, and eight times the loads in flight of the one-thread-per-output version (which took 31 us for the 148 per-CTA
// partials of one edge backward -- a fixed cost per layer at every GPU count).  `src` may be null for an inactive lane; the
// result is valid in warp 0 only.  `sm` = 256 floats.
__device__ __forceinline__ float ordered_sum_block8(const float* __restrict__ src, int parts, int64_t stride, float* sm) {
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (src != nullptr) {
    int p = g;
    for (; p + 24 < parts; p += 32) {
      const float x0 = __ldg(src + int64_t(p) * stride), x1 = __ldg(src + int64_t(p + 8) * stride);
      const float x2 = __ldg(src + int64_t(p + 16) * stride), x3 = __ldg(src + int64_t(p + 24) * stride);
      a0 += x0; a1 += x1; a2 += x2; a3 += x3;
    }
    if (p < parts) a0 += __ldg(src + int64_t(p) * stride);
    if (p + 8 < parts) a1 += __ldg(src + int64_t(p + 8) * stride);
    if (p + 16 < parts) a2 += __ldg(src + int64_t(p + 16) * stride);
  }
  sm[g * 32 + lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  return ((sm[lane] + sm[32 + lane]) + (sm[64 + lane] + sm[96 + lane])) + ((sm[128 + lane] + sm[160 + lane]) + (sm[192 + lane] + sm[224 + lane]));
}

}  // namespace hgn
