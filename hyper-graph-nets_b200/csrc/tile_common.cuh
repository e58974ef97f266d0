// Constants, the packed-weight blob layout and small device helpers shared by the bf16 tcgen05 kernels
// (mlp_tc.cu: generic fused MLP tile; edge_tc.cu: projected edge update + node projection).
#pragma once
#include "common.cuh"
#include "tc05.cuh"

namespace hgn {

using namespace tc05;

constexpr int kTile = 128;                     // rows per tile = TMEM lanes
constexpr int kPanel = kPanelBytes128;         // 16 KiB: [128][64] bf16
constexpr int kChunkBytes = 2 * kPanel;        // one 128x128 bf16 operand: 32 KiB
constexpr int kStages = 2;
constexpr int kEpiThreads = 256;               // warps 0-7
constexpr int kProdThreads = 128;              // warps 8-11
constexpr int kTileThreads = kEpiThreads + kProdThreads + 32;
constexpr int kMaxResidentChunks = 3;
constexpr float kEps = 1e-5f;

struct PackedTc {   // byte offsets inside the packed blob
  size_t w0, w1, w2, params, total;   // params: b0 b1 b2 gamma beta (fp32 x 128 each)
  __host__ __device__ explicit PackedTc(int n_chunks) {
    size_t o = 0;
    w0 = o; o += size_t(n_chunks) * kD * kD * 2;
    w1 = o; o += size_t(kD) * kD * 2;
    w2 = o; o += size_t(kD) * kD * 2;
    params = o; o += 5 * kD * 4;
    total = o;
  }
};

// copy a [128][128] bf16 row-major block (row pitch `ld` elements) into two SW128 panels
__device__ __forceinline__ void load_weight_block(uint32_t smem_dst, const __nv_bfloat16* __restrict__ g, int64_t ld, int tid, int nthreads) {
  for (int q = tid; q < kTile * 16; q += nthreads) {
    const int row = q >> 4, c16 = q & 15;
    cp_async16(smem_dst + (c16 >> 3) * kPanel + sw128_chunk(row, c16 & 7), g + int64_t(row) * ld + c16 * 8);
  }
}

// ---- epilogue helpers shared by the projected edge kernels ----------------------------------------------------
__device__ __forceinline__ void st_shared128(uint32_t addr, const uint32_t* w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void ld_shared128(uint32_t addr, uint32_t* w) {
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr) : "memory");
}
// packed bf16 helpers (one SASS instruction each: F2FP.RELU.BF16.F32.PACK_AB, HFMA2.BF16_V2, HSET2 + HMUL2)
__device__ __forceinline__ uint32_t cvt_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t x, uint32_t y) {
  uint32_t d;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(y));
  return d;
}
// g * [h > 0] per 16-bit half: ReLU backward against the stored activation
__device__ __forceinline__ uint32_t relu_bwd_bf16x2(uint32_t g, uint32_t h) {
  uint32_t d;
  asm("{\n\t.reg .b32 m;\n\tset.gt.bf16x2.bf16x2 m, %2, %3;\n\tmul.rn.bf16x2 %0, %1, m;\n\t}" : "=r"(d) : "r"(g), "r"(h), "r"(0u));
  return d;
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u)); }
// 256-bit read-only load of a gathered table row.  Not allocated in L1: with 227 KiB of the SM's 256 KiB given to shared memory the L1 holds
// a few hundred lines, a tile gathers 1-2 k of them, and A/B on the cfg5 layer measured the allocating form 1.3 % (backward) / 2.7 % (forward)
// slower (bitwise identical results).
__device__ __forceinline__ void ldg256_l1(const void* p, uint32_t* v) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Column sums over the 32 lanes of a warp: x[j] is this lane's (= this row's) value in column j.  Returns, in lane j,
// the sum over all 32 lanes of column j (recursive halving: 31 shuffles).  Fixed association order.
__device__ __forceinline__ float warp_colsum32(float (&x)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? x[i] : x[i + off];
      const float keep = up ? x[i + off] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return x[0];
}


int tc_sm_count();                 // SMs of the current device (mlp_tc.cu)
uint32_t* debug_buffer_device();   // cabi.cu: host-mapped words, readable after a device trap
// fixed-order reductions of the per-CTA weight-gradient partials (mlp_f32.cu)
size_t colsum5_workspace_bytes(int64_t rows);                                  // segment.cu
int colsum5_bf16(const void* const* mats, float* const* outs, int64_t rows, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t st);
void launch_reduce_weight_partials(const float* partial, int parts, int n_chunks, float* gW0, float* gW1, float* gW2, cudaStream_t st);

}  // namespace hgn
