// Round-2 preparation (compile-checked here, NOT yet run: the round's GPU budget was spent): how fast can one SM gather 256-byte
// bf16 rows by index?  DESIGN.md s3.2 measured that the thread-per-row `ld.global.v8` gathers of the edge backward kernel retire
// about one 32-byte sector per cycle and stall the issuing epilogue warps.  This probe times, per SM and for the same index lists,
//   mode 0  the kernel's pattern: 256 threads, thread = row, 4 x 32-byte loads of its 128-byte half row (checksummed in registers)
//   mode 1  TMA gather:  cp.async.bulk.tensor.2d ... tile::gather4 (UTMALDG.2D.GATHER4 on sm_100a), one elected thread, 4 rows x
//           128 bytes per instruction into a 128B-swizzled shared-memory tile, double buffered, consumers read rows with ld.shared
// and checks both against a host checksum.  Build and run (one GPU):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scripts/probes/gather_rate scripts/probes/gather_rate.cu -lcuda
//   scripts/probes/gather_rate [rows_in_table=1000000] [gathers=5992002] [sorted=0|1]
// Open questions it answers: the box shape tile::gather4 wants ({64, 1} -- what CuTe's make_tma_copy_atom encodes -- then {64, 4}), bytes/cycle/SM of both
// modes for random and for receiver-sorted indices, and whether one issuing thread keeps up.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int kD = 128, kTile = 128, kThreads = 288;       // 8 consumer warps + 1 TMA warp
constexpr uint32_t kPanel = 128 * 128;                      // one 128-row x 64-column bf16 panel, 128B-swizzled
constexpr uint32_t kTileBytes = 2 * kPanel;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               :: "r"(dst), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}
// byte offset of 16-byte chunk c (0..7) of row r inside a 128B-swizzled panel
__device__ __forceinline__ uint32_t sw128(int r, int c) { return uint32_t(r) * 128u + uint32_t((c ^ (r & 7)) << 4); }

__global__ void __launch_bounds__(kThreads, 1)
gather_kernel(int mode, const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ idx, int64_t gathers, const __grid_constant__ CUtensorMap tm,
              unsigned long long* checksum, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[2], empty[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (gathers + kTile - 1) / kTile;
  const int64_t my_tiles = (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  if (tid == 0) { for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 256); } asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const long long t0 = clock64();
  unsigned long long sum = 0;
  if (mode == 0) {
    if (warp < 8) {
      const int r = (warp & 3) * 32 + lane, hh = warp >> 2;
      for (int64_t t = 0; t < my_tiles; ++t) {
        const int64_t g = (blockIdx.x + t * gridDim.x) * kTile + r;
        if (g < gathers) {
          const __nv_bfloat16* src = table + int64_t(__ldg(idx + g)) * kD + hh * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) {                       // the kernel's ldg256_l1: one 32-byte sector per lane and instruction
            uint32_t v[8];
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(src + 16 * k));
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += v[j];
          }
        }
      }
    }
  } else {
    if (warp == 8) {
      if (lane == 0) {
        for (int64_t t = 0; t < my_tiles; ++t) {
          const int s = int(t & 1);
          if (t >= 2) mbar_wait(&empty[s], uint32_t((t >> 1) - 1) & 1);
          const int64_t g0 = (blockIdx.x + t * gridDim.x) * kTile;
          mbar_expect_tx(&full[s], kTileBytes);
          const uint32_t dst = smem_u32(smem) + s * kTileBytes;
          for (int q = 0; q < kTile / 4; ++q) {
            int r[4];
            for (int k = 0; k < 4; ++k) { const int64_t g = g0 + 4 * q + k; r[k] = g < gathers ? __ldg(idx + g) : 0; }
            tma_gather4(dst + q * 4 * 128, &tm, 0, r[0], r[1], r[2], r[3], &full[s]);
            tma_gather4(dst + kPanel + q * 4 * 128, &tm, 64, r[0], r[1], r[2], r[3], &full[s]);
          }
        }
      }
    } else {
      const int r = (warp & 3) * 32 + lane, hh = warp >> 2;
      for (int64_t t = 0; t < my_tiles; ++t) {
        const int s = int(t & 1);
        mbar_wait(&full[s], uint32_t(t >> 1) & 1);
        const int64_t g = (blockIdx.x + t * gridDim.x) * kTile + r;
        const uint8_t* base = smem + s * kTileBytes + hh * kPanel;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint4 v = *reinterpret_cast<const uint4*>(base + sw128(r, k));
          if (g < gathers) sum += (unsigned long long)v.x + v.y + v.z + v.w;
        }
        mbar_arrive(&empty[s]);
      }
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  if (lane == 0 && sum) atomicAdd(checksum, sum);
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 1000000, gathers = argc > 2 ? atoll(argv[2]) : 5992002;
  const int sorted = argc > 3 ? atoi(argv[3]) : 0;
  std::vector<uint16_t> h_table(size_t(n) * kD);
  std::vector<int32_t> h_idx(gathers);
  uint64_t seed = 88172645463325252ull;
  auto rnd = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return seed; };
  for (auto& v : h_table) v = uint16_t(rnd());
  for (int64_t g = 0; g < gathers; ++g) h_idx[g] = sorted ? int32_t(g * n / gathers) : int32_t(rnd() % uint64_t(n));
  unsigned long long want = 0;
  for (int64_t g = 0; g < gathers; ++g) {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(h_table.data() + size_t(h_idx[g]) * kD);
    for (int k = 0; k < kD / 2; ++k) want += row[k];
  }
  __nv_bfloat16* d_table; int32_t* d_idx; unsigned long long* d_sum; long long* d_cycles;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaMalloc(&d_table, h_table.size() * 2)); CK(cudaMalloc(&d_idx, size_t(gathers) * 4)); CK(cudaMalloc(&d_sum, 8)); CK(cudaMalloc(&d_cycles, sms * 8));
  CK(cudaMemcpy(d_table, h_table.data(), h_table.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx, h_idx.data(), size_t(gathers) * 4, cudaMemcpyHostToDevice));
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(p);
  CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(2 * kTileBytes)));
  for (int box_rows = 1; box_rows <= 4; box_rows += 3) {
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {kD, cuuint64_t(n)}, gstride[1] = {kD * 2};
    const cuuint32_t box[2] = {64, cuuint32_t(box_rows)}, estride[2] = {1, 1};
    const CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_table, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("tensor map with box {64, %d}: encode rc=%d\n", box_rows, int(rc));
    if (rc != CUDA_SUCCESS) continue;
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 0 && box_rows != 1) continue;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(d_sum, 0, 8));
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        gather_kernel<<<sms, kThreads, 2 * kTileBytes>>>(mode, d_table, d_idx, gathers, tm, d_sum, d_cycles);
        cudaEventRecord(b);
        const cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("mode %d box rows %d: %s\n", mode, box_rows, cudaGetErrorString(err)); return 1; }
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        unsigned long long got = 0; std::vector<long long> cyc(sms);
        cudaMemcpy(&got, d_sum, 8, cudaMemcpyDeviceToHost); cudaMemcpy(cyc.data(), d_cycles, sms * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (long long c : cyc) mx = c > mx ? c : mx;
        printf("mode %d (%s) box rows %d rep %d: %.3f ms, %.1f GB/s, %.1f B/cycle/SM, checksum %s\n", mode, mode ? "TMA gather4" : "ld.global per row", box_rows, rep,
               ms, double(gathers) * 256 / ms / 1e6, double(gathers) * 256 / sms / double(mx), got == want ? "ok" : "MISMATCH");
      }
    }
  }
  return 0;
}
