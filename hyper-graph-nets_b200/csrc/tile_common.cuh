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

// Optional "projected" first layer (edge update): layer 0 consumes only the chunks listed in `ch` (the edge rows) with the
// W0 column block starting at chunk `w0_chunk0` of a `w0_chunks`-wide packed W0, and the epilogue adds the per-node
// pre-projections  proj_s[senders[row]] + proj_r[receivers[row]]  (= v[s] W0[:,0:128]^T + v[r] W0[:,128:256]^T computed once
// per NODE by proj_tc_kernel instead of once per EDGE here).  w0_chunks == 0 -> plain mode (w0_chunks = n_chunks, chunk0 = 0).
struct PreAdd {
  const __nv_bfloat16 *proj_s, *proj_r;
  const int32_t *senders, *receivers;
  int w0_chunks, w0_chunk0;
};

int mlp_tc_forward_pre(int64_t rows, const hgn_chunks* ch, const void* packed, const void* resid, int64_t resid_off, void* out,
                       const PreAdd& pre, const char* name, cudaStream_t st);
int tc_sm_count();                 // SMs of the current device (mlp_tc.cu)
uint32_t* debug_buffer_device();   // cabi.cu: host-mapped words, readable after a device trap
// fixed-order reductions of the per-CTA weight-gradient partials (mlp_f32.cu)
void launch_reduce_weight_partials(const float* partial, int parts, int n_chunks, float* gW0, float* gW1, float* gW2, cudaStream_t st);

}  // namespace hgn
