// Descriptor probe for the tcgen05 path (development tool, run once on a B200 through gpurun):
// checks every operand form the MLP tile kernels rely on against a host GEMM and prints the error.
//   mode 0: A K-major (smem)   x B K-major (smem)         forward / dgrad-with-transposed-copy form
//   mode 1: A MN-major (smem)  x B K-major
//   mode 2: A K-major          x B MN-major               dgrad form (weights as stored, [out][in])
//   mode 3: A MN-major         x B MN-major               wgrad form (activations as stored, [row][feature])
//   mode 4: A from TMEM (packed bf16 pairs) x B K-major   H1/H2 kept in tensor memory
// D[m][n] = sum_k A[m][k] * B[n][k], M = N = K = 128.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../tc05.cuh"

using namespace tc05;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D,
             int mode, int swap_lbo_sbo) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 2 panels x 16 KiB
  uint8_t* sB = smem + 32768;         // 2 panels x 16 KiB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const bool a_mn = (mode == 1 || mode == 3), b_mn = (mode == 2 || mode == 3), a_tmem = (mode == 4);

  // operand staging with plain stores: element (r, c) of the smem tile; K-major tile = [m][k], MN-major = [k][m]
  for (int i = tid; i < 128 * 128; i += 128) {
    int r = i >> 7, c = i & 127;
    __nv_bfloat16 av = a_mn ? A[c * 128 + r] : A[r * 128 + c];
    __nv_bfloat16 bv = b_mn ? B[c * 128 + r] : B[r * 128 + c];
    *reinterpret_cast<__nv_bfloat16*>(sA + (c >> 6) * kPanelBytes128 + sw128_offset(r, c & 63)) = av;
    *reinterpret_cast<__nv_bfloat16*>(sB + (c >> 6) * kPanelBytes128 + sw128_offset(r, c & 63)) = bv;
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t tmem_a = tmem_base_s + 128;

  if (a_tmem) {
    // lane = row m; 32-bit column j holds (A[m][2j], A[m][2j+1])
    const int m = tid;
    for (int j0 = 0; j0 < 64; j0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const uint32_t*>(&A[m * 128 + 2 * (j0 + j)]);
      tmem_st8(tmem_a + (uint32_t(warp * 32) << 16) + j0, v);
    }
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }

  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 128, a_mn, b_mn);
    for (int ks = 0; ks < 8; ++ks) {
      uint64_t ad = 0, bd;
      auto mn_desc = [&](uint8_t* base) {
        uint32_t addr = smem_u32(base) + ks * 2048;
        return swap_lbo_sbo ? make_sdesc(addr, 1024, kPanelBytes128) : sdesc_mnmajor(addr, kPanelBytes128);
      };
      auto k_desc = [&](uint8_t* base) { return sdesc_kmajor(smem_u32(base) + (ks >> 2) * kPanelBytes128 + (ks & 3) * 32); };
      if (!a_tmem) ad = a_mn ? mn_desc(sA) : k_desc(sA);
      bd = b_mn ? mn_desc(sB) : k_desc(sB);
      if (a_tmem) mma_ts(tmem_d, tmem_a + ks * 8, bd, idesc, ks > 0);
      else        mma_ss(tmem_d, ad, bd, idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();

  const int m = tid;   // warp w owns TMEM lanes 32w .. 32w+31
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_d + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[m * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base_s);
}

static float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

int main() {
  const int n = 128;
  std::vector<__nv_bfloat16> hA(n * n), hB(n * n);
  srand(1);
  for (int i = 0; i < n * n; ++i) {
    hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
    hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
  }
  std::vector<float> ref(n * n);
  for (int m = 0; m < n; ++m)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int k = 0; k < n; ++k) s += double(bf2f(hA[m * n + k])) * double(bf2f(hB[j * n + k]));
      ref[m * n + j] = float(s);
    }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, n * n * 2); cudaMalloc(&dB, n * n * 2); cudaMalloc(&dD, n * n * 4);
  cudaMemcpy(dA, hA.data(), n * n * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), n * n * 2, cudaMemcpyHostToDevice);
  const int smem_bytes = 65536 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  int failures = 0;
  for (int mode = 0; mode < 5; ++mode) {
    for (int swap = 0; swap < ((mode >= 1 && mode <= 3) ? 2 : 1); ++swap) {
      cudaMemset(dD, 0, n * n * 4);
      probe_kernel<<<1, 128, smem_bytes>>>(dA, dB, dD, mode, swap);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) {
        printf("mode %d swap %d: CUDA error %s\n", mode, swap, cudaGetErrorString(err));
        return 2;   // context is gone after a trap
      }
      std::vector<float> out(n * n);
      cudaMemcpy(out.data(), dD, n * n * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int i = 0; i < n * n; ++i) { maxerr = fmax(maxerr, fabs(out[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
      bool ok = maxerr < 1e-3 * maxref;
      printf("mode %d swap_lbo_sbo %d: max_abs_err %.3e (max_ref %.3e) %s\n", mode, swap, maxerr, maxref, ok ? "OK" : "MISMATCH");
      if (!ok && swap == 0 && !(mode >= 1 && mode <= 3)) failures++;
    }
  }
  printf("probe done, hard failures=%d\n", failures);
  return 0;
}
