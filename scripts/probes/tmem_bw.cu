// Probe: TMEM read throughput (tcgen05.ld 32x32b) for 4 / 8 / 16 warps per CTA.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../hyper-graph-nets_b200/csrc/tc05.cuh"
using namespace tc05;
__global__ void k(int iters, long long* out, int* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tptr);
  fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t base = tptr + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 32;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[32];
    tmem_ld32(base + ((i & 1) * 128), v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= v[j];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345) sink[0] = 1;
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tptr);
}
int main() {
  long long* out; int* sink; cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
  for (int warps : {1, 4, 8, 16}) {
    const int iters = 2000;
    k<<<1, warps * 32>>>(iters, out, sink); cudaDeviceSynchronize();
    k<<<1, warps * 32>>>(iters, out, sink);
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    const double bytes = double(iters) * warps * 32 * 32 * 4;
    printf("warps=%2d: %lld cycles, %.1f cyc/iter, %.1f B/cyc per CTA  (%s)\n", warps, h, double(h) / iters, bytes / h, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
