// fp32 parity path of the fused "MLP tile" (HGN_F32): gather + 3-layer MLP + LayerNorm + residual with
// FFMA register tiles.  It exists to meet the 1e-5 (fp32) parity bar of the reference's fp32 torch path
// (src/migration/graphnet.py:22-48, meshgraphnet.py:53-60,93-108); the throughput path is mlp_tc.cu.
//
// Tile: 64 rows x 128 features per CTA, 256 threads, each thread a 4x8 register block.  The A operand
// (a gathered 64x128 chunk, later H1/H2/gradients) lives in shared memory; the weights stream through a
// double-buffered 32x128 shared-memory slice with cp.async.  The [rows, 128*n_chunks] concatenation of
// the reference is never formed: layer 0 accumulates one 128-wide chunk after the other.
#include "common.cuh"
#include "tc05.cuh"

namespace hgn {

using tc05::cp_async16;
using tc05::cp_async16_zfill;
using tc05::cp_async_commit;
using tc05::cp_async_wait;
using tc05::smem_u32;

constexpr int kTileRows = 64;
constexpr int kLdx = 132;                    // padded row pitch (floats) of the A operand tile
constexpr int kSliceK = 32;                  // weight slice: 32 k-rows x 128 outputs
constexpr int kThreads = 256;
constexpr float kLnEps = 1e-5f;

struct PackedF32 {   // offsets (floats) inside the packed blob for n_chunks
  int64_t w0t, w1t, w2t, w0, w1, w2, b0, b1, b2, gamma, beta, total;
  __host__ __device__ explicit PackedF32(int n_chunks) {
    const int64_t k0 = int64_t(n_chunks) * kD;
    int64_t o = 0;
    w0t = o; o += k0 * kD;
    w1t = o; o += kD * kD;
    w2t = o; o += kD * kD;
    w0 = o; o += k0 * kD;
    w1 = o; o += kD * kD;
    w2 = o; o += kD * kD;
    b0 = o; o += kD; b1 = o; o += kD; b2 = o; o += kD; gamma = o; o += kD; beta = o; o += kD;
    total = o;
  }
};

__global__ void pack_f32_kernel(int n_chunks, const float* W0, const float* b0, const float* W1, const float* b1, const float* W2,
                                const float* b2, const float* gamma, const float* beta, float* packed) {
  const PackedF32 L(n_chunks);
  const int64_t k0 = int64_t(n_chunks) * kD;
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i < k0 * kD) {            // W0 is [128][k0]
    const int64_t o = i / k0, k = i - o * k0;
    packed[L.w0 + i] = W0[i];
    packed[L.w0t + k * kD + o] = W0[i];
  }
  if (i < kD * kD) {
    const int64_t o = i / kD, k = i - o * kD;
    packed[L.w1 + i] = W1[i];
    packed[L.w1t + k * kD + o] = W1[i];
    packed[L.w2 + i] = W2[i];
    packed[L.w2t + k * kD + o] = W2[i];
  }
  if (i < kD) {
    packed[L.b0 + i] = b0[i]; packed[L.b1 + i] = b1[i]; packed[L.b2 + i] = b2[i];
    packed[L.gamma + i] = gamma[i]; packed[L.beta + i] = beta[i];
  }
}

// ---- building blocks ---------------------------------------------------------------------------
// stage a gathered 64x128 fp32 chunk into Xs (one cp.async group member; caller commits)
__device__ __forceinline__ void stage_chunk(float* Xs, const float* __restrict__ src, const int32_t* __restrict__ idx,
                                            int64_t row_offset, int64_t row0, int64_t rows, int tid) {
  const int c4 = tid & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r = (tid >> 5) + 8 * j;
    const int64_t grow = row0 + r;
    const bool valid = grow < rows;
    int64_t srow = 0;
    if (valid) srow = idx ? int64_t(idx[grow]) : grow + row_offset;
    cp_async16_zfill(smem_u32(Xs + r * kLdx + c4 * 4), src + srow * kD + c4 * 4, valid);
  }
}

__device__ __forceinline__ void stage_slice(float* Ws, const float* __restrict__ Wg, int64_t ld, int tid) {
  // 32 k-rows x 128 floats; row k at Wg + k*ld
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int q = tid + j * kThreads;       // 0..1023 float4s
    const int k = q >> 5, c4 = q & 31;
    cp_async16(smem_u32(Ws + k * kD + c4 * 4), Wg + int64_t(k) * ld + c4 * 4);
  }
}

// acc[4][8] += Xs[64 x 128] * B[128 x 128], B row k at Bg + k*ld (streamed through Ws[2][32][128]).
// Any cp.async issued by the caller before this call (the A tile) is covered by the first wait.
__device__ __forceinline__ void gemm_block(float (&acc)[4][8], const float* Xs, const float* __restrict__ Bg, int64_t ld,
                                           float* Ws, int tid) {
  const int tx = tid & 15, ty = tid >> 4;
  stage_slice(Ws, Bg, ld, tid);
  cp_async_commit();
#pragma unroll 1
  for (int s = 0; s < kD / kSliceK; ++s) {
    if (s + 1 < kD / kSliceK) {
      stage_slice(Ws + ((s + 1) & 1) * kSliceK * kD, Bg + int64_t(s + 1) * kSliceK * ld, ld, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* W = Ws + (s & 1) * kSliceK * kD + tx * 8;
    const float* X = Xs + (ty * 4) * kLdx + s * kSliceK;
#pragma unroll
    for (int kk = 0; kk < kSliceK; kk += 4) {
      float4 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(X + i * kLdx + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b0 = *reinterpret_cast<const float4*>(W + (kk + q) * kD);
        const float4 b1 = *reinterpret_cast<const float4*>(W + (kk + q) * kD + 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float av = q == 0 ? a[i].x : q == 1 ? a[i].y : q == 2 ? a[i].z : a[i].w;
          acc[i][0] = fmaf(av, b0.x, acc[i][0]); acc[i][1] = fmaf(av, b0.y, acc[i][1]);
          acc[i][2] = fmaf(av, b0.z, acc[i][2]); acc[i][3] = fmaf(av, b0.w, acc[i][3]);
          acc[i][4] = fmaf(av, b1.x, acc[i][4]); acc[i][5] = fmaf(av, b1.y, acc[i][5]);
          acc[i][6] = fmaf(av, b1.z, acc[i][6]); acc[i][7] = fmaf(av, b1.w, acc[i][7]);
        }
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

// bias (+ optional ReLU), then park the 4x8 block in the shared A tile for the next GEMM
__device__ __forceinline__ void bias_act_store(float (&acc)[4][8], const float* __restrict__ bias, bool relu, float* Xs, int tid) {
  const int tx = tid & 15, ty = tid >> 4;
  const float4 b0 = *reinterpret_cast<const float4*>(bias + tx * 8);
  const float4 b1 = *reinterpret_cast<const float4*>(bias + tx * 8 + 4);
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[i][j] + bb[j];
      acc[i][j] = relu ? fmaxf(v, 0.f) : v;
    }
    if (Xs) {
      float* p = Xs + (ty * 4 + i) * kLdx + tx * 8;
      *reinterpret_cast<float4*>(p) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
}

__device__ __forceinline__ float row_sum16(float v) {   // sum over the 16 lanes that share a row group
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v;
}

// forward through the three linears; leaves y (pre-LayerNorm, bias added) in acc.
// H1 -> h1s, H2 -> h2s (may both alias xs for the forward-only kernel).
__device__ __forceinline__ void mlp_recompute(float (&acc)[4][8], const hgn_chunks& ch, const float* __restrict__ packed,
                                              const PackedF32& L, int64_t row0, int64_t rows, float* xs, float* h1s, float* h2s,
                                              float* ws, int tid) {
  zero_acc(acc);
  for (int c = 0; c < ch.n_chunks; ++c) {
    stage_chunk(xs, static_cast<const float*>(ch.src[c]), ch.idx[c], ch.row_offset[c], row0, rows, tid);
    gemm_block(acc, xs, packed + L.w0t + int64_t(c) * kD * kD, kD, ws, tid);
  }
  bias_act_store(acc, packed + L.b0, true, h1s, tid);
  __syncthreads();
  zero_acc(acc);
  gemm_block(acc, h1s, packed + L.w1t, kD, ws, tid);
  bias_act_store(acc, packed + L.b1, true, h2s, tid);
  __syncthreads();
  zero_acc(acc);
  gemm_block(acc, h2s, packed + L.w2t, kD, ws, tid);
  bias_act_store(acc, packed + L.b2, false, nullptr, tid);
}

// ---- forward kernel ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
mlp_fwd_f32_kernel(int64_t rows, hgn_chunks ch, const float* __restrict__ packed, const float* __restrict__ resid,
                   int64_t resid_off, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem_f[];
  float* xs = smem_f;                         // [64][132]
  float* ws = smem_f + kTileRows * kLdx;      // [2][32][128]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const PackedF32 L(ch.n_chunks);
  const int64_t row0 = int64_t(blockIdx.x) * kTileRows;
  float acc[4][8];
  mlp_recompute(acc, ch, packed, L, row0, rows, xs, xs, xs, ws, tid);

  const float4 g0 = *reinterpret_cast<const float4*>(packed + L.gamma + tx * 8), g1 = *reinterpret_cast<const float4*>(packed + L.gamma + tx * 8 + 4);
  const float4 e0 = *reinterpret_cast<const float4*>(packed + L.beta + tx * 8), e1 = *reinterpret_cast<const float4*>(packed + L.beta + tx * 8 + 4);
  const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bet[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[i][j];
    const float mean = row_sum16(s) * (1.0f / kD);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = acc[i][j] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(row_sum16(q) * (1.0f / kD) + kLnEps);
    const int64_t grow = row0 + ty * 4 + i;
    if (grow < rows) {
      const float* rp = resid + (grow + resid_off) * kD + tx * 8;
      const float4 r0 = *reinterpret_cast<const float4*>(rp), r1 = *reinterpret_cast<const float4*>(rp + 4);
      const float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rr[j] + ((acc[i][j] - mean) * rstd * gam[j] + bet[j]);
      float* op = out + grow * kD + tx * 8;
      *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(op + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// ---- backward: recompute + data gradients ----------------------------------------------------------
// Writes grad_chunk[c] tiles, the pre-activation gradients G2 (=dY), G1, G0 and the activations H1, H2
// of this slab (inputs of the weight-gradient kernel), and per-CTA partial sums for gamma/beta.
struct BwdOut {
  float* grad_chunk[HGN_MAX_CHUNKS];
};

__device__ __forceinline__ void store_tile(float* __restrict__ g, const float (&v)[4][8], int64_t row0, int64_t rows, int64_t slab0, int tid) {
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t grow = row0 + ty * 4 + i;
    if (grow < rows) {
      float* p = g + (grow - slab0) * kD + tx * 8;
      *reinterpret_cast<float4*>(p) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
    }
  }
}

__device__ __forceinline__ void tile_to_smem(float* Xs, const float (&v)[4][8], int tid) {
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* p = Xs + (ty * 4 + i) * kLdx + tx * 8;
    *reinterpret_cast<float4*>(p) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
  }
}

__global__ void __launch_bounds__(kThreads)
mlp_bwd_f32_kernel(int64_t rows, int64_t slab0, int64_t slab_rows, hgn_chunks ch, const float* __restrict__ packed,
                   const float* __restrict__ grad_out, BwdOut go, float* __restrict__ G2, float* __restrict__ G1,
                   float* __restrict__ G0, float* __restrict__ H1, float* __restrict__ H2, float* __restrict__ ln_partial,
                   int accumulate_ln, int resid_chunk) {
  extern __shared__ __align__(16) float smem_f[];
  float* xs = smem_f;                              // A tile: chunk / gradients
  float* h1s = xs + kTileRows * kLdx;
  float* h2s = h1s + kTileRows * kLdx;
  float* ws = h2s + kTileRows * kLdx;              // [2][32][128], also reduction scratch
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const PackedF32 L(ch.n_chunks);
  const int64_t row0 = slab0 + int64_t(blockIdx.x) * kTileRows;
  const int64_t row_end = min(rows, slab0 + slab_rows);
  float acc[4][8];
  mlp_recompute(acc, ch, packed, L, row0, row_end, xs, h1s, h2s, ws, tid);

  // LayerNorm backward on the register tile
  const float4 g0 = *reinterpret_cast<const float4*>(packed + L.gamma + tx * 8), g1 = *reinterpret_cast<const float4*>(packed + L.gamma + tx * 8 + 4);
  const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  float dgam[8], dbet[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { dgam[j] = 0.f; dbet[j] = 0.f; }
  float dy[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[i][j];
    const float mean = row_sum16(s) * (1.0f / kD);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = acc[i][j] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(row_sum16(q) * (1.0f / kD) + kLnEps);
    const int64_t grow = row0 + ty * 4 + i;
    float dO[8];
    if (grow < row_end) {
      const float* gp = grad_out + grow * kD + tx * 8;
      const float4 a = *reinterpret_cast<const float4*>(gp), b = *reinterpret_cast<const float4*>(gp + 4);
      dO[0] = a.x; dO[1] = a.y; dO[2] = a.z; dO[3] = a.w; dO[4] = b.x; dO[5] = b.y; dO[6] = b.z; dO[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) dO[j] = 0.f;
    }
    float m1 = 0.f, m2 = 0.f, yh[8], dyh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      yh[j] = (acc[i][j] - mean) * rstd;
      dyh[j] = dO[j] * gam[j];
      m1 += dyh[j];
      m2 = fmaf(dyh[j], yh[j], m2);
      dgam[j] = fmaf(dO[j], yh[j], dgam[j]);
      dbet[j] += dO[j];
    }
    m1 = row_sum16(m1) * (1.0f / kD);
    m2 = row_sum16(m2) * (1.0f / kD);
#pragma unroll
    for (int j = 0; j < 8; ++j) dy[i][j] = rstd * (dyh[j] - m1 - yh[j] * m2);
  }
  // per-CTA gamma/beta partials: fixed-order reduction over the 16 row groups through shared memory
  __syncthreads();
  {
    float* red = ws;   // [16][256]
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[ty * 256 + tx * 8 + j] = dgam[j]; red[ty * 256 + 128 + tx * 8 + j] = dbet[j]; }
    __syncthreads();
    float s = 0.f;
    for (int g = 0; g < 16; ++g) s += red[g * 256 + tid];
    float* dst = ln_partial + int64_t(blockIdx.x) * 256 + tid;
    *dst = accumulate_ln ? *dst + s : s;
    __syncthreads();
  }

  // dY -> G2 ; dH2 = dY W2, masked by H2 > 0 -> G1 ; dH1 = G1 W1, masked -> G0 ; dX_c = G0 W0[:, c]
  store_tile(G2, dy, row0, row_end, slab0, tid);
  tile_to_smem(xs, dy, tid);
  __syncthreads();
  zero_acc(acc);
  gemm_block(acc, xs, packed + L.w2, kD, ws, tid);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = h2s[(ty * 4 + i) * kLdx + tx * 8 + j] > 0.f ? acc[i][j] : 0.f;
  store_tile(G1, acc, row0, row_end, slab0, tid);
  tile_to_smem(xs, acc, tid);
  __syncthreads();
  zero_acc(acc);
  gemm_block(acc, xs, packed + L.w1, kD, ws, tid);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = h1s[(ty * 4 + i) * kLdx + tx * 8 + j] > 0.f ? acc[i][j] : 0.f;
  store_tile(G0, acc, row0, row_end, slab0, tid);
  tile_to_smem(xs, acc, tid);
  // activations for the weight-gradient kernel
  {
    float t[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) t[i][j] = h1s[(ty * 4 + i) * kLdx + tx * 8 + j];
    store_tile(H1, t, row0, row_end, slab0, tid);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) t[i][j] = h2s[(ty * 4 + i) * kLdx + tx * 8 + j];
    store_tile(H2, t, row0, row_end, slab0, tid);
  }
  __syncthreads();
  const int64_t k0 = int64_t(ch.n_chunks) * kD;
  for (int c = 0; c < ch.n_chunks; ++c) {
    if (!go.grad_chunk[c]) continue;
    zero_acc(acc);
    gemm_block(acc, xs, packed + L.w0 + int64_t(c) * kD, k0, ws, tid);
    if (c == resid_chunk) {   // residual branch: d(out)/d(resid) = identity
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t grow = row0 + ty * 4 + i;
        if (grow < row_end) {
          const float* gp = grad_out + grow * kD + tx * 8;
          const float4 a = *reinterpret_cast<const float4*>(gp), b = *reinterpret_cast<const float4*>(gp + 4);
          acc[i][0] += a.x; acc[i][1] += a.y; acc[i][2] += a.z; acc[i][3] += a.w;
          acc[i][4] += b.x; acc[i][5] += b.y; acc[i][6] += b.z; acc[i][7] += b.w;
        }
      }
    }
    store_tile(go.grad_chunk[c], acc, row0, row_end, 0, tid);
  }
}

// ---- weight gradients: dW[o][i] = sum_r G[r][o] * Z[r][i] over a slab, split over row parts ----------
// grid (parts, n_z): z < n_chunks -> Z = gathered chunk z with G = G0 ; z = n_chunks -> (G1, H1) ;
// z = n_chunks+1 -> (G2, H2).  partial[(part*n_z + z)][128][128] (+ bias column sums for G0/G1/G2).
__global__ void __launch_bounds__(kThreads)
mlp_wgrad_f32_kernel(int64_t rows, int64_t slab0, int64_t slab_rows, int rows_per_part, hgn_chunks ch,
                     const float* __restrict__ G0, const float* __restrict__ G1, const float* __restrict__ G2,
                     const float* __restrict__ H1, const float* __restrict__ H2, float* __restrict__ partial,
                     float* __restrict__ bias_partial, int accumulate) {
  extern __shared__ __align__(16) float smem_f[];
  float* gs = smem_f;                 // [32][128]
  float* zs = smem_f + 32 * kD;       // [32][128]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;   // outputs o = ty*8.., i = tx*8..
  const int part = blockIdx.x, z = blockIdx.y, n_z = gridDim.y, nch = ch.n_chunks;
  const float* G = z < nch ? G0 : (z == nch ? G1 : G2);
  const int64_t row_end = min(rows, slab0 + slab_rows);
  const int64_t r_beg = slab0 + int64_t(part) * rows_per_part;
  const int64_t r_end = min(row_end, r_beg + rows_per_part);
  float acc[8][8], bsum[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) { bsum[a] = 0.f;
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f; }
  const bool do_bias = (z == 0 || z >= nch) && tx == 0;
  for (int64_t r0 = r_beg; r0 < r_end; r0 += 32) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = tid + j * kThreads;
      const int r = q >> 5, c4 = q & 31;
      const int64_t grow = r0 + r;
      const bool valid = grow < r_end;
      const int64_t lrow = valid ? grow - slab0 : 0;
      cp_async16_zfill(smem_u32(gs + r * kD + c4 * 4), G + lrow * kD + c4 * 4, valid);
      const float* zsrc;
      if (z < nch) {
        int64_t srow = 0;
        if (valid) srow = ch.idx[z] ? int64_t(ch.idx[z][grow]) : grow + ch.row_offset[z];
        zsrc = static_cast<const float*>(ch.src[z]) + srow * kD;
      } else {
        zsrc = (z == nch ? H1 : H2) + lrow * kD;
      }
      cp_async16_zfill(smem_u32(zs + r * kD + c4 * 4), zsrc + c4 * 4, valid);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const float4 ga = *reinterpret_cast<const float4*>(gs + r * kD + ty * 8), gb = *reinterpret_cast<const float4*>(gs + r * kD + ty * 8 + 4);
      const float4 za = *reinterpret_cast<const float4*>(zs + r * kD + tx * 8), zb = *reinterpret_cast<const float4*>(zs + r * kD + tx * 8 + 4);
      const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
      const float zv[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        bsum[a] += gv[a];
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(gv[a], zv[b], acc[a][b]);
      }
    }
  }
  float* P = partial + (int64_t(part) * n_z + z) * kD * kD;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    float* p = P + (ty * 8 + a) * kD + tx * 8;
    float4 v0 = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]), v1 = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
    if (accumulate) {
      const float4 p0 = *reinterpret_cast<const float4*>(p), p1 = *reinterpret_cast<const float4*>(p + 4);
      v0.x += p0.x; v0.y += p0.y; v0.z += p0.z; v0.w += p0.w; v1.x += p1.x; v1.y += p1.y; v1.z += p1.z; v1.w += p1.w;
    }
    *reinterpret_cast<float4*>(p) = v0;
    *reinterpret_cast<float4*>(p + 4) = v1;
  }
  if (do_bias) {
    const int which = z == 0 ? 0 : (z == nch ? 1 : 2);    // b0 <- G0, b1 <- G1, b2 <- G2
    float* bp = bias_partial + (int64_t(part) * 3 + which) * kD + ty * 8;
#pragma unroll
    for (int a = 0; a < 8; ++a) bp[a] = accumulate ? bp[a] + bsum[a] : bsum[a];
  }
}

// out[i] = sum_p in[p*stride + i]  (fixed order; 32 outputs per 256-thread block, common.cuh ordered_sum_block8)
__global__ void __launch_bounds__(256)
reduce_parts_kernel(const float* __restrict__ in, int parts, int64_t stride, int64_t n, float* __restrict__ out) {
  __shared__ float sm[256];
  const int64_t i = blockIdx.x * int64_t(32) + (threadIdx.x & 31);
  const float s = ordered_sum_block8(i < n ? in + i : nullptr, parts, stride, sm);
  if (threadIdx.x < 32 && i < n) out[i] = s;
}
// dW0[o][c*128 + i] = sum_p partial[p][z=c][o][i]
__global__ void __launch_bounds__(256)
reduce_w0_kernel(const float* __restrict__ partial, int parts, int n_z, int n_chunks, float* __restrict__ dW0) {
  __shared__ float sm[256];
  const int64_t k0 = int64_t(n_chunks) * kD;
  const int64_t i = blockIdx.x * int64_t(32) + (threadIdx.x & 31);
  const bool on = i < kD * k0;
  const int64_t o = on ? i / k0 : 0, col = on ? i - o * k0 : 0;
  const int c = int(col / kD), ii = int(col - int64_t(c) * kD);
  const float s = ordered_sum_block8(on ? partial + (int64_t(c) * kD + o) * kD + ii : nullptr, parts, int64_t(n_z) * kD * kD, sm);
  if (threadIdx.x < 32 && on) dW0[i] = s;
}

// partial layout [parts][n_chunks + 2][128][128]: z < n_chunks -> W0 chunk z ; n_chunks -> W1 ; n_chunks + 1 -> W2
void launch_reduce_weight_partials(const float* partial, int parts, int n_chunks, float* gW0, float* gW1, float* gW2, cudaStream_t st) {
  const int n_z = n_chunks + 2;
  const int64_t k0 = int64_t(n_chunks) * kD, blk = int64_t(kD) * kD;
  HGN_TIMED("reduce_weight_partials", st);
  reduce_w0_kernel<<<unsigned(ceil_div(kD * k0, 32)), 256, 0, st>>>(partial, parts, n_z, n_chunks, gW0);
  reduce_parts_kernel<<<unsigned(ceil_div(blk, 32)), 256, 0, st>>>(partial + int64_t(n_chunks) * blk, parts, int64_t(n_z) * blk, blk, gW1);
  reduce_parts_kernel<<<unsigned(ceil_div(blk, 32)), 256, 0, st>>>(partial + int64_t(n_chunks + 1) * blk, parts, int64_t(n_z) * blk, blk, gW2);
}

struct BwdLayoutF32 {
  int64_t slab_rows, tiles, parts;
  int rows_per_part;
  size_t g2, g1, g0, h1, h2, partial, bias_partial, ln_partial, total;
};

static BwdLayoutF32 bwd_layout_f32(int64_t rows, int n_chunks) {
  BwdLayoutF32 L{};
  L.slab_rows = rows < (int64_t(1) << 19) ? (rows > 0 ? rows : 1) : (int64_t(1) << 19);   // <= 512K rows per pass
  L.slab_rows = ceil_div(L.slab_rows, kTileRows) * kTileRows;
  L.tiles = L.slab_rows / kTileRows;
  L.parts = L.slab_rows >= 64 * 256 ? 64 : ceil_div(L.slab_rows, 256);
  L.rows_per_part = int(ceil_div(ceil_div(L.slab_rows, L.parts), 32) * 32);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t act = size_t(L.slab_rows) * kD * 4;
  L.g2 = take(act); L.g1 = take(act); L.g0 = take(act); L.h1 = take(act); L.h2 = take(act);
  L.partial = take(size_t(L.parts) * (n_chunks + 2) * kD * kD * 4);
  L.bias_partial = take(size_t(L.parts) * 3 * kD * 4);
  L.ln_partial = take(size_t(L.tiles) * 256 * 4);
  L.total = off;
  return L;
}

size_t mlp_f32_packed_bytes(int n_chunks) { return size_t(PackedF32(n_chunks).total) * 4; }

int mlp_f32_pack(int n_chunks, const float* W0, const float* b0, const float* W1, const float* b1, const float* W2,
                 const float* b2, const float* gamma, const float* beta, void* packed, cudaStream_t st) {
  const int64_t n = int64_t(n_chunks) * kD * kD;
  pack_f32_kernel<<<unsigned(ceil_div(n, 256)), 256, 0, st>>>(n_chunks, W0, b0, W1, b1, W2, b2, gamma, beta, static_cast<float*>(packed));
  HGN_LAUNCH_OK("pack_f32");
  return HGN_OK;
}

int mlp_f32_forward(int64_t rows, const hgn_chunks* ch, const void* packed, const void* resid, int64_t resid_off, void* out,
                    cudaStream_t st) {
  const size_t smem = size_t(kTileRows * kLdx + 2 * kSliceK * kD) * 4;
  static bool configured = false;
  if (!configured) {
    HGN_CUDA_OK(cudaFuncSetAttribute(mlp_fwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    configured = true;
  }
  HGN_TIMED("mlp_fwd_f32", st);
  mlp_fwd_f32_kernel<<<unsigned(ceil_div(rows, kTileRows)), kThreads, smem, st>>>(rows, *ch, static_cast<const float*>(packed),
                                                                               static_cast<const float*>(resid), resid_off,
                                                                               static_cast<float*>(out));
  HGN_LAUNCH_OK("mlp_fwd_f32");
  return HGN_OK;
}

size_t mlp_f32_backward_workspace_bytes(int64_t rows, int n_chunks) { return bwd_layout_f32(rows, n_chunks).total; }

int mlp_f32_backward(int64_t rows, const hgn_chunks* ch, const void* packed, const void* grad_out, int resid_chunk, void* const* grad_chunk,
                     float* gW0, float* gb0, float* gW1, float* gb1, float* gW2, float* gb2, float* ggamma, float* gbeta,
                     void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int nch = ch->n_chunks;
  const BwdLayoutF32 L = bwd_layout_f32(rows, nch);
  if (workspace_bytes < L.total) { set_error("mlp_backward(f32): workspace %zu < %zu", workspace_bytes, L.total); return HGN_ERR_WORKSPACE; }
  char* ws = static_cast<char*>(workspace);
  float *G2 = (float*)(ws + L.g2), *G1 = (float*)(ws + L.g1), *G0 = (float*)(ws + L.g0), *H1 = (float*)(ws + L.h1), *H2 = (float*)(ws + L.h2);
  float *partial = (float*)(ws + L.partial), *bias_partial = (float*)(ws + L.bias_partial), *ln_partial = (float*)(ws + L.ln_partial);
  const size_t smem_bwd = size_t(3 * kTileRows * kLdx + 2 * kSliceK * kD) * 4;
  const size_t smem_wg = size_t(2 * 32 * kD) * 4;
  static bool configured = false;
  if (!configured) {
    HGN_CUDA_OK(cudaFuncSetAttribute(mlp_bwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bwd)));
    configured = true;
  }
  BwdOut go{};
  for (int c = 0; c < nch; ++c) go.grad_chunk[c] = grad_chunk ? static_cast<float*>(grad_chunk[c]) : nullptr;
  const int n_z = nch + 2;
  int pass = 0;
  for (int64_t slab0 = 0; slab0 < rows || pass == 0; slab0 += L.slab_rows, ++pass) {
    const int64_t this_rows = rows - slab0 < L.slab_rows ? rows - slab0 : L.slab_rows;
    if (this_rows > 0) {
      HGN_TIMED("mlp_bwd_f32", st);
      mlp_bwd_f32_kernel<<<unsigned(ceil_div(this_rows, kTileRows)), kThreads, smem_bwd, st>>>(
          rows, slab0, L.slab_rows, *ch, static_cast<const float*>(packed), static_cast<const float*>(grad_out), go, G2, G1, G0, H1,
          H2, ln_partial, pass > 0, resid_chunk);
      HGN_LAUNCH_OK("mlp_bwd_f32");
    }
    if (pass == 0 && this_rows < L.slab_rows) {
      // tiles that were not launched must not contribute stale partials
      const int64_t launched = this_rows > 0 ? ceil_div(this_rows, kTileRows) : 0;
      if (launched < L.tiles)
        HGN_CUDA_OK(cudaMemsetAsync(ln_partial + launched * 256, 0, size_t(L.tiles - launched) * 256 * 4, st));
    }
    dim3 grid(unsigned(L.parts), unsigned(n_z));
    { HGN_TIMED("mlp_wgrad_f32", st);
    mlp_wgrad_f32_kernel<<<grid, kThreads, smem_wg, st>>>(rows, slab0, L.slab_rows, L.rows_per_part, *ch, G0, G1, G2, H1, H2, partial,
                                                          bias_partial, pass > 0);
    }
    HGN_LAUNCH_OK("mlp_wgrad_f32");
    if (rows == 0) break;
  }
  launch_reduce_weight_partials(partial, int(L.parts), nch, gW0, gW1, gW2, st);
  reduce_parts_kernel<<<kD / 32, 256, 0, st>>>(bias_partial, int(L.parts), 3 * kD, kD, gb0);
  reduce_parts_kernel<<<kD / 32, 256, 0, st>>>(bias_partial + kD, int(L.parts), 3 * kD, kD, gb1);
  reduce_parts_kernel<<<kD / 32, 256, 0, st>>>(bias_partial + 2 * kD, int(L.parts), 3 * kD, kD, gb2);
  reduce_parts_kernel<<<kD / 32, 256, 0, st>>>(ln_partial, int(L.tiles), 256, kD, ggamma);
  reduce_parts_kernel<<<kD / 32, 256, 0, st>>>(ln_partial + kD, int(L.tiles), 256, kD, gbeta);
  HGN_LAUNCH_OK("mlp_bwd_f32 reductions");
  return HGN_OK;
}

}  // namespace hgn
