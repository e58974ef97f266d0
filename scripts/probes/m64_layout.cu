// Round-2 preparation (compile-checked, NOT yet run): where does tcgen05.mma put a cta_group::1 M = 64 accumulator in tensor memory,
// and what does it cost?  A backward kernel with TWO 64-row tiles in flight needs both answers.
//   part 1  layout: D1[m][n] = m + 1 and D2[m][n] = n + 1 from rank-1 operands (exact in bf16 / fp32), M = 64, N = 128, K = 16; all 128
//           lanes x 128 columns of both accumulators are read back with 32x32b loads and the (lane, column) -> (m, n) map is printed
//           A third accumulator asks for lane offset 64 (taddr lane field = 64): can a second M = 64 tile live in lanes 64-127 of the SAME
//           columns?  (That is what lets two 64-row tiles share the 128 chain-accumulator columns of today's kernel.)
//   part 2  rate: cycles per MMA for M = 64 and M = 128 (N = 128, K = 16, operands in shared memory), 2 000 back-to-back instructions
// Build / run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hyper-graph-nets_b200/csrc -o scripts/probes/m64_layout scripts/probes/m64_layout.cu
#include <cstdio>
#include <vector>
#include "tc05.cuh"

using namespace tc05;

__global__ void __launch_bounds__(128, 1) layout_kernel(float* __restrict__ out, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA1 = smem;             // [128 rows][64 k] panel: A1[m][0] = m + 1
  uint8_t* sB1 = smem + 16384;     // B1[n][0] = 1
  uint8_t* sA2 = smem + 32768;     // A2[m][0] = 1
  uint8_t* sB2 = smem + 49152;     // B2[n][0] = n + 1
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4 * 16384 / 2; i += 128) reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16(0.f);
  __syncthreads();
  {
    const int r = tid;               // row r, column 0 of every panel
    *reinterpret_cast<__nv_bfloat16*>(sA1 + sw128_offset(r, 0)) = __float2bfloat16(float(r + 1));
    *reinterpret_cast<__nv_bfloat16*>(sB1 + sw128_offset(r, 0)) = __float2bfloat16(1.f);
    *reinterpret_cast<__nv_bfloat16*>(sA2 + sw128_offset(r, 0)) = __float2bfloat16(1.f);
    *reinterpret_cast<__nv_bfloat16*>(sB2 + sw128_offset(r, 0)) = __float2bfloat16(float(r + 1));
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t t0 = tmem_base_s;
  // clear both accumulator regions through all 128 lanes so that untouched cells read back as zero
  {
    uint32_t z[32];
    for (int j = 0; j < 32; ++j) z[j] = 0u;
    for (int c = 0; c < 384; c += 32) tmem_st32(t0 + (uint32_t(warp * 32) << 16) + c, z);
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  if (tid == 0) {
    const uint32_t id64 = make_idesc_bf16(64, 128, 0, 0);
    mma_ss(t0, sdesc_kmajor(smem_u32(sA1)), sdesc_kmajor(smem_u32(sB1)), id64, false);          // D1 = row id
    mma_ss(t0 + 128, sdesc_kmajor(smem_u32(sA2)), sdesc_kmajor(smem_u32(sB2)), id64, false);    // D2 = column id
    mma_ss(t0 + (64u << 16) + 256, sdesc_kmajor(smem_u32(sA1)), sdesc_kmajor(smem_u32(sB1)), id64, false);   // D3 = row id, lane offset 64
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < 384; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(t0 + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[tid * 384 + c0 + j] = __uint_as_float(v[j]);
  }
  fence_before_sync();
  __syncthreads();
  // part 2: issue rate
  if (tid == 0) {
    for (int pass = 0; pass < 2; ++pass) {
      const uint32_t id = make_idesc_bf16(pass == 0 ? 64 : 128, 128, 0, 0);
      mbar_init(&bar, 1); mbar_init_fence();
      const long long c0 = clock64();
      for (int i = 0; i < 2000; ++i) mma_ss(t0 + 384, sdesc_kmajor(smem_u32(sA1)), sdesc_kmajor(smem_u32(sB1)), id, i > 0);
      mma_commit(&bar);
      mbar_wait(&bar, 0);
      cycles[pass] = clock64() - c0;
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}

int main() {
  float* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 128 * 384 * 4); cudaMalloc(&d_cyc, 16);
  cudaMemset(d_out, 0, 128 * 384 * 4);
  const int smem_bytes = 65536 + 1024;
  cudaFuncSetAttribute(layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  layout_kernel<<<1, 128, smem_bytes>>>(d_out, d_cyc);
  const cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); return 2; }
  std::vector<float> h(128 * 384); long long cyc[2];
  cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost);
  printf("M = 64 accumulator: TMEM lane -> row m (0 = lane not written), read at column 0; and column -> n at the first written lane\n");
  int first_lane = -1;
  for (int lane = 0; lane < 128; ++lane) {
    const int m = int(h[lane * 384 + 0]);
    if (m != 0 && first_lane < 0) first_lane = lane;
    printf("%s%3d:%-3d", lane % 16 == 0 ? "\n  " : " ", lane, m);
  }
  printf("\n  columns of lane %d (accumulator 2 = column id):", first_lane);
  if (first_lane >= 0) for (int c = 0; c < 128; c += 8) printf(" c%d->n%d", c, int(h[first_lane * 384 + 128 + c]) - 1);
  int rows_seen = 0; bool seen[65] = {false};
  for (int lane = 0; lane < 128; ++lane) for (int c = 0; c < 128; ++c) { const int m = int(h[lane * 384 + c]); if (m >= 1 && m <= 64 && !seen[m]) { seen[m] = true; ++rows_seen; } }
  printf("\n  distinct rows found anywhere in the 128 x 128 region: %d of 64\n", rows_seen);
  printf("third accumulator (requested lane offset 64, columns 256..383): lane -> row m at column 256");
  for (int lane = 0; lane < 128; ++lane) printf("%s%3d:%-3d", lane % 16 == 0 ? "\n  " : " ", lane, int(h[lane * 384 + 256]));
  printf("\n");
  printf("issue + execute, 2000 MMAs N=128 K=16: M=64 %.1f cycles each, M=128 %.1f cycles each\n", cyc[0] / 2000.0, cyc[1] / 2000.0);
  return 0;
}
