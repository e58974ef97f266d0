// bf16 tcgen05 "projected" edge update (HGN_BF16 only) -- the per-layer hot kernel of the processor.
//
// Reference arithmetic (src/migration/graphnet.py:22-32):  e' = e + LN(MLP([v[s] | v[r] | e])).
// The first linear is split by input block,  W0 = [Ws | Wr | We]  (columns 0:128 | 128:256 | 256:384):
//     pre0_e = Ws v[s_e] + Wr v[r_e] + We e_e + b0 = Ps[s_e] + Pr[r_e] + We e_e + b0,     Ps = v Ws^T, Pr = v Wr^T
// The two node-side products are computed once per NODE (proj kernel below; E ~ 6 N on a triangle mesh) and the
// edge kernels only gather-add their rows.  The same split is used backwards: with G0 = d loss / d pre0,
//     d v  += segsum_senders(G0) Ws + segsum_receivers(G0) Wr          (node-level dgrad, proj kernel with B MN-major)
//     dWs  = segsum_senders(G0)^T v,  dWr = segsum_receivers(G0)^T v   (node-level wgrad)
// so the edge backward kernel needs only We, W1, W2 and writes only d e and G0.
//
// Edge backward kernel (persistent, one CTA per SM, one 128-edge tile at a time):
//   warps 0-7   epilogue: row = TMEM lane, warp w handles lane quadrant w%4 and column half w/4; bias/ReLU/LayerNorm
//               forward+backward arithmetic in fp32; every intermediate tile (H1, H2, dY, dH2', dH1') is written as bf16
//               into one of four rotating 32 KiB shared-memory buffers in the 128B-swizzled operand layout
//   warps 8-9   producers: cp.async of the next tile's edge rows into the buffer that becomes free after step 4, and the
//               bias-gradient column sums of the dY / dH2' / dH1' tiles read back from shared memory
//   warp 10     MMA issuer: per tile 6 chain GEMMs (recompute L0 L1 L2, dgrad dH2 dH1 dXe; A operand K-major from the
//               buffers, B = resident We/W1/W2 panels, read MN-major for the dgrads) and 3 weight-gradient GEMMs
//               (dW2 += dY^T H2, dW1 += dH2'^T H1, dWe += dH1'^T e; both operands MN-major from the SAME buffers) whose
//               fp32 accumulators stay in TMEM for the CTA's whole lifetime.
// TMEM (512 columns): chain accumulator [0,128) | dW2 [128,256) | dW1 [256,384) | dWe [384,512).
// Nothing but d e and G0 (and per-CTA fp32 partials at the end) is written to HBM; the previous design wrote six
// [E,128] workspaces per layer.  All reductions have a fixed order (per-CTA partials, then a sequential sum over CTAs).
#include <cuda.h>
#include <stdlib.h>

#include "tile_common.cuh"

namespace hgn {

// =========================================================================================================
// node projection and its data gradient:  out_o[rows,128] = sum_c in_c[rows,128] * B(chunk chunk0 + o + c)
//   forward (b_mn = 0): n_in = 1, n_out = 2:  Ps = v Ws^T, Pr = v Wr^T          (B K-major:  out[n] = sum_k in[k] W[n][k])
//   dgrad   (b_mn = 1): n_in = 2, n_out = 1:  dv = Gs Ws + Gr Wr                (B MN-major: out[k] = sum_n in[n] W[n][k])
// HBM-bound (256 B read + 512 B written per node row forward).  2-stage cp.async ring, double-buffered TMEM accumulators.
// =========================================================================================================
constexpr int kLinEpi = 128, kLinProd = 64;
constexpr int kLinThreads = kLinEpi + kLinProd + 32;

struct LinArgs {
  const __nv_bfloat16* in[2];
  __nv_bfloat16* out[2];
  const __nv_bfloat16* add;      // optional [rows,128]: added to out[0] in the epilogue (fp32), e.g. the node update's share of d loss / d v
  int n_in, n_out, b_mn, w0_chunks, chunk0;
};

__global__ void __launch_bounds__(kLinThreads, 1)
proj_tc_kernel(int64_t rows, int64_t num_tiles, const uint8_t* __restrict__ packed, LinArgs la) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t stage_bytes = uint32_t(la.n_in) * kChunkBytes;
  const uint32_t w_off = 0, st_off = 2 * kChunkBytes, bars_off = st_off + 2 * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + bars_off);   // full[2] empty[2] accf[2] acce[2] tmem
  const PackedTc P(la.w0_chunks);
  {
    const __nv_bfloat16* w0g = reinterpret_cast<const __nv_bfloat16*>(packed + P.w0) + la.chunk0 * kD;
    const int n_w = la.n_in > la.n_out ? la.n_in : la.n_out;
    for (int c = 0; c < n_w; ++c) load_weight_block(sbase + w_off + c * kChunkBytes, w0g + c * kD, int64_t(la.w0_chunks) * kD, tid, kLinThreads);
    cp_async_commit();
    if (tid == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&bars[0 + s], kLinProd); mbar_init(&bars[2 + s], 1); mbar_init(&bars[4 + s], 1); mbar_init(&bars[6 + s], kLinEpi);
      }
      mbar_init_fence();
    }
    if (warp == 6) tmem_alloc<512>(reinterpret_cast<uint32_t*>(&bars[8]));
    cp_async_wait<0>();
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars[8]);
  const int64_t my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp >= 4 && warp < 6) {
    // ---- producers -----------------------------------------------------------------------------------
    const int ptid = tid - kLinEpi;
    if (ptid < 2) mbar_arrive(&bars[2 + ptid]);          // ring starts empty
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int stage = int(it & 1);
      mbar_wait(&bars[2 + stage], uint32_t(it >> 1) & 1, 40);
      const int64_t row0 = (blockIdx.x + it * gridDim.x) * kTile;
      for (int c = 0; c < la.n_in; ++c) {
        const uint32_t dst = sbase + st_off + stage * stage_bytes + c * kChunkBytes;
        const __nv_bfloat16* src = la.in[c];
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
          const int qd = ptid + kLinProd * j, row = qd >> 4, c16 = qd & 15;
          const int64_t grow = row0 + row;
          const bool valid = grow < rows;
          cp_async16_zfill(dst + (c16 >> 3) * kPanel + sw128_chunk(row, c16 & 7), src + (valid ? grow : 0) * kD + c16 * 8, valid);
        }
      }
      cp_async_commit();
      if (it > 0) {                                       // keep two tiles of loads in flight
        cp_async_wait<1>();
        fence_async_smem();
        mbar_arrive(&bars[0 + int((it - 1) & 1)]);
      }
    }
    cp_async_wait<0>();
    fence_async_smem();
    if (my_tiles > 0) mbar_arrive(&bars[0 + int((my_tiles - 1) & 1)]);
  } else if (warp == 6) {
    // ---- MMA issuer ----------------------------------------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 128, 0, la.b_mn);
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int s = int(it & 1);
        mbar_wait(&bars[0 + s], uint32_t(it >> 1) & 1, 41);
        if (it >= 2) mbar_wait(&bars[6 + s], uint32_t((it >> 1) - 1) & 1, 42);     // accumulator set s drained
        fence_after_sync();
        for (int o = 0; o < la.n_out; ++o) {
          const uint32_t acc = tmem_base + s * 256 + o * 128;
          for (int c = 0; c < la.n_in; ++c) {
            const uint32_t a_addr = sbase + st_off + s * stage_bytes + c * kChunkBytes;
            const uint32_t b_addr = sbase + w_off + (o + c) * kChunkBytes;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t ad = sdesc_kmajor(a_addr + (ks >> 2) * kPanel + (ks & 3) * 32);
              const uint64_t bd = la.b_mn ? sdesc_mnmajor(b_addr + ks * 2048, kPanel) : sdesc_kmajor(b_addr + (ks >> 2) * kPanel + (ks & 3) * 32);
              mma_ss(acc, ad, bd, idesc, (c | ks) != 0);
            }
          }
        }
        mma_commit(&bars[2 + s]);
        mma_commit(&bars[4 + s]);
      }
    }
  } else if (warp < 4) {
    // ---- epilogue: accumulator -> bf16 rows ----------------------------------------------------------
    const uint32_t lane_addr = uint32_t(warp * 32) << 16;
    const int r = warp * 32 + lane;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int s = int(it & 1);
      mbar_wait(&bars[4 + s], uint32_t(it >> 1) & 1, 43);
      fence_after_sync();
      const int64_t grow = (blockIdx.x + it * gridDim.x) * kTile + r;
      const bool valid = grow < rows;
      for (int o = 0; o < la.n_out; ++o) {
        __nv_bfloat16* op = la.out[o] + (valid ? grow : 0) * kD;
#pragma unroll 1
        for (int cg = 0; cg < 4; ++cg) {
          uint32_t v[32], w[16];
          tmem_ld32(tmem_base + lane_addr + s * 256 + o * 128 + cg * 32, v);
          // tcgen05.ld / tcgen05.wait are .sync.aligned: every lane of the warp executes them together, so only the global loads
          // are guarded by the per-lane `valid` (a tail tile of a small graph has valid and invalid rows inside one warp)
          const bool add_row = la.add != nullptr && o == 0 && valid;
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = 0u;
          if (add_row) { ldg256(la.add + grow * kD + cg * 32, w); ldg256(la.add + grow * kD + cg * 32 + 16, w + 8); }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 x = unpack_bf16x2(w[j]);
            w[j] = add_row ? pack_bf16(__uint_as_float(v[2 * j]) + x.x, __uint_as_float(v[2 * j + 1]) + x.y)
                           : pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          }
          if (valid) { stg256(op + cg * 32, w); stg256(op + cg * 32 + 16, w + 8); }
        }
      }
      fence_before_sync();
      mbar_arrive(&bars[6 + s]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 6) tmem_dealloc<512>(tmem_base);
}

// =========================================================================================================
// edge backward
// =========================================================================================================
constexpr int kEbEpiThreads = 256, kEbProdThreads = 64;
constexpr int kEbThreads = kEbEpiThreads + kEbProdThreads + 32;    // 11 warps
constexpr uint32_t kEbWe = 0, kEbW1 = kChunkBytes, kEbW2 = 2 * kChunkBytes, kEbBuf = 3 * kChunkBytes;
constexpr uint32_t kEbParams = kEbBuf + 4 * kChunkBytes;           // b0 b1 b2 gamma beta (fp32 x 128 each)
constexpr uint32_t kEbBars = kEbParams + 5 * kD * 4;
constexpr uint32_t kEbSmem = kEbBars + 128;                        // 232 064 B of the 232 448 available
// kEbG + k / kEbCs + k (k = 0 dY, 1 dH2', 2 dH1'): one barrier per tile-in-buffer hand-over, so each completes exactly one
// phase per tile and no waiter can fall two phases behind (a parity wait cannot tell phase n from phase n + 2)
enum { kEbFull = 0, kEbWg = 1, kEbAcc = 2, kEbEpi = 3, kEbG = 4, kEbCs = 7, kEbTmem = 10, kEbAfree = 11, kEbFinal = 12, kEbDe = 13, kEbDeFree = 14, kEbG0Free = 15 };

struct EdgeBwdArgs {
  const __nv_bfloat16 *edge, *proj_s, *proj_r;
  const int32_t *senders, *receivers;
  const __nv_bfloat16* grad_out;    // [E,128] dense part of d loss / d e' (may be null)
  const __nv_bfloat16* grad_agg;    // [N,128] gathered through receivers: gradient of the 'sum' aggregate (may be null)
  __nv_bfloat16 *grad_edge, *grad_pre0;
  float* w_partial;                 // [grid][3][128][128]  z = 0: dWe, 1: dW1, 2: dW2
  float* epi_colpart;               // [grid][4][2][128]    beta, gamma partial column sums per lane quadrant
  float* prod_colpart;              // [grid][3][128]       db2, db1, db0
  long long* timeline;              // development: clock64 stamps of block 0 ([tile][32]) when HGN_TC_ABLATE has bit 64
  int w0_chunks, w0_chunk0;         // W0 is [128][128 w0_chunks]; the dense input multiplies chunk w0_chunk0
  int ablate;                       // development switches (HGN_TC_ABLATE): 1 no table/gradient loads, 2 no HBM stores,
                                    // 4 no LayerNorm-vector column sums, 8 no bias column sums, 16 no weight-gradient MMAs
};

__device__ __forceinline__ void stamp(const EdgeBwdArgs& a, int64_t t, int slot) {
  if (a.timeline != nullptr && blockIdx.x == 0 && t < 8) a.timeline[t * 48 + slot] = clock64();
}

__global__ void __launch_bounds__(kEbThreads, 1)
edge_bwd_tc_kernel(int64_t rows, int64_t num_tiles, const uint8_t* __restrict__ packed, EdgeBwdArgs a,
                   const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_g0,
                   const __grid_constant__ CUtensorMap tm_de) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kEbBars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PackedTc P(a.w0_chunks);
  float* prm = reinterpret_cast<float*>(smem + kEbParams);
  {
    const float* pg = reinterpret_cast<const float*>(packed + P.params);
    for (int i = tid; i < 5 * kD; i += kEbThreads) prm[i] = pg[i];
    const __nv_bfloat16* w0g = reinterpret_cast<const __nv_bfloat16*>(packed + P.w0);
    load_weight_block(sbase + kEbWe, w0g + a.w0_chunk0 * kD, int64_t(a.w0_chunks) * kD, tid, kEbThreads);
    load_weight_block(sbase + kEbW1, reinterpret_cast<const __nv_bfloat16*>(packed + P.w1), kD, tid, kEbThreads);
    load_weight_block(sbase + kEbW2, reinterpret_cast<const __nv_bfloat16*>(packed + P.w2), kD, tid, kEbThreads);
    cp_async_commit();
    if (tid == 0) {
      mbar_init(&bars[kEbFull], 1);                           // one arrive.expect_tx per tile, completed by the TMA bytes
      mbar_init(&bars[kEbWg], 1);
      mbar_init(&bars[kEbFinal], 1);
      mbar_init(&bars[kEbDe], kEbEpiThreads / 32);            // epilogue barriers: one arrival per WARP (lane 0 after __syncwarp):
      mbar_init(&bars[kEbDeFree], 1);
      mbar_init(&bars[kEbG0Free], 1);
      mbar_init(&bars[kEbAcc], 1);
      mbar_init(&bars[kEbEpi], kEbEpiThreads / 32);           // 256 arrivals on one mbarrier serialise in the shared-memory unit
      mbar_init(&bars[kEbAfree], kEbEpiThreads / 32);
      for (int k = 0; k < 3; ++k) { mbar_init(&bars[kEbG + k], kEbEpiThreads / 32); mbar_init(&bars[kEbCs + k], kEbProdThreads / 32); }
      mbar_init_fence();
    }
    if (warp == 10) tmem_alloc<512>(reinterpret_cast<uint32_t*>(&bars[kEbTmem]));
    cp_async_wait<0>();
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars[kEbTmem]);
  const int64_t my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  // buffer roles of tile t: 0 = S (edge rows), 1 = A (H1), 2 = B (H2, then dH2'), 3 = C (dY, then dH1'); they rotate by one
  // buffer per tile so that S(t+1) = A(t), the first buffer that falls free (after step 4), receives the prefetch
  auto buf = [&](int role, int64_t t) -> uint32_t { return sbase + kEbBuf + uint32_t((role + t) & 3) * kChunkBytes; };

  if (warp == 8 || warp == 9) {
    // =============================== producers =========================================================
    const int ptid = tid - kEbEpiThreads, pw = warp - 8;
    // column sums of a bf16 tile in a buffer: warp pw owns panel pw (columns 64 pw ..); lane l reads the 16-byte piece l & 7
    // (8 columns) of rows 4 i + (l >> 3).  Each lane keeps fp32 partials over ALL its tiles; the four row groups are
    // combined once, after the last tile (fixed order).
    float cs[3][8];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[k][j] = 0.f;
    auto colsum = [&](uint32_t base, float (&acc8)[8]) {
      if (a.ablate & 8) return;
      const uint32_t pbase = base + pw * kPanel;
      const int c = lane & 7, ro = lane >> 3;
      float2 t[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        uint32_t w[4];
        ld_shared128(pbase + sw128_chunk(4 * i + ro, c), w);
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = __fadd2_rn(t[j], unpack_bf16x2(w[j]));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc8[2 * j] += t[j].x; acc8[2 * j + 1] += t[j].y; }
    };
    // edge rows of tile t: two TMA boxes (128 rows x 64 columns, 128B-swizzled) into the tile's S buffer
    auto load_e = [&](int64_t t) {
      if (ptid == 0) {
        const uint32_t dst = buf(0, t);
        const int y = int((blockIdx.x + t * gridDim.x) * kTile);
        mbar_expect_tx(&bars[kEbFull], kChunkBytes);
        tma_load_2d(dst, &tm_e, 0, y, &bars[kEbFull]);
        tma_load_2d(dst + kPanel, &tm_e, 64, y, &bars[kEbFull]);
      }
    };
    if (my_tiles > 0) load_e(0);
    for (int64_t t = 0; t < my_tiles; ++t) {
      const uint32_t par = uint32_t(t) & 1;
      const bool more = t + 1 < my_tiles;
      if (more && !(a.ablate & 33)) {
        // everything tile t+1 will read from HBM is pulled into L2 a whole tile ahead, by these otherwise idle warps (never
        // by the epilogue warps: their proxy fences wait for outstanding prefetches): edge rows, dense gradient rows, and --
        // through the tile's sender / receiver indices -- the rows of the two node tables and of the aggregate gradient
        const int64_t row0 = (blockIdx.x + (t + 1) * gridDim.x) * kTile;
        auto pf = [](const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); };
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int64_t grow = row0 + ptid + 64 * j;
          if (grow < rows) {
            const int64_t si = a.senders != nullptr ? int64_t(__ldg(a.senders + grow)) : grow;
            const int64_t ri = a.receivers != nullptr ? int64_t(__ldg(a.receivers + grow)) : grow;
            pf(a.edge + grow * kD); pf(a.edge + grow * kD + 64);
            if (a.grad_out != nullptr) { pf(a.grad_out + grow * kD); pf(a.grad_out + grow * kD + 64); }
            pf(a.proj_s + si * kD); pf(a.proj_s + si * kD + 64);
            if (a.proj_r != nullptr) { pf(a.proj_r + ri * kD); pf(a.proj_r + ri * kD + 64); }
            if (a.grad_agg != nullptr) { pf(a.grad_agg + ri * kD); pf(a.grad_agg + ri * kD + 64); }
          }
        }
      }
      mbar_wait(&bars[kEbG + 0], par, 50);
      colsum(buf(3, t), cs[0]);                              // dY
      __syncwarp(); if (lane == 0) mbar_arrive(&bars[kEbCs + 0]);
      mbar_wait(&bars[kEbG + 1], par, 51);
      colsum(buf(2, t), cs[1]);                              // dH2'
      __syncwarp(); if (lane == 0) mbar_arrive(&bars[kEbCs + 1]);
      mbar_wait(&bars[kEbAfree], par, 52);                   // H1 has been read back and the dW1 MMAs are done:
      if (more) load_e(t + 1);                               // buffer A(t) = S(t+1) takes the next tile's edge rows
      mbar_wait(&bars[kEbG + 2], par, 53);                   // dH1' = G0 is in buffer C (and fenced for the async proxy)
      if (ptid == 0 && !(a.ablate & 2)) {                    // G0 goes to HBM straight from the operand buffer
        const int y = int((blockIdx.x + t * gridDim.x) * kTile);
        tma_store_2d(&tm_g0, buf(3, t), 0, y);
        tma_store_2d(&tm_g0, buf(3, t) + kPanel, 64, y);
        tma_store_commit();
      }
      colsum(buf(3, t), cs[2]);
      if (ptid == 0) tma_store_wait_read<0>();               // the store has read the buffer before E1 of the next tile reuses it
      __syncwarp(); if (lane == 0) mbar_arrive(&bars[kEbCs + 2]);
    }
    if (ptid == 0) tma_store_wait<0>();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = cs[k][j];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        cs[k][j] = v;
      }
      if (lane < 8) {
        float* dst = a.prod_colpart + (int64_t(blockIdx.x) * 3 + k) * kD + pw * 64 + lane * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(cs[k][0], cs[k][1], cs[k][2], cs[k][3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(cs[k][4], cs[k][5], cs[k][6], cs[k][7]);
      }
    }
  } else if (warp == 10) {
    // =============================== MMA issuer =========================================================
    if (lane == 0) {
      const uint32_t id_kk = make_idesc_bf16(128, 128, 0, 0), id_kmn = make_idesc_bf16(128, 128, 0, 1), id_mm = make_idesc_bf16(128, 128, 1, 1);
      const uint32_t acc = tmem_base, dW2 = tmem_base + 128, dW1 = tmem_base + 256, dWe = tmem_base + 384;
      uint32_t epi_phase = 0;
      auto wait_epi = [&]() {
        mbar_spin(&bars[kEbEpi], epi_phase++ & 1, 60);
        fence_after_sync();
      };
      auto chain = [&](uint32_t a_addr, uint32_t b_addr, bool b_mn) {      // acc = A[128 x 128] (K-major) * B
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = sdesc_kmajor(a_addr + (ks >> 2) * kPanel + (ks & 3) * 32);
          const uint64_t bd = b_mn ? sdesc_mnmajor(b_addr + ks * 2048, kPanel) : sdesc_kmajor(b_addr + (ks >> 2) * kPanel + (ks & 3) * 32);
          mma_ss(acc, ad, bd, b_mn ? id_kmn : id_kk, ks != 0);
        }
      };
      auto wgrad = [&](uint32_t d, uint32_t g_addr, uint32_t z_addr, bool first) {   // d (+)= G^T Z over the tile's 128 rows
        if (a.ablate & 16) return;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_ss(d, sdesc_mnmajor(g_addr + ks * 2048, kPanel), sdesc_mnmajor(z_addr + ks * 2048, kPanel), id_mm, !(first && ks == 0));
      };
      for (int64_t t = 0; t < my_tiles; ++t) {
        const uint32_t S = buf(0, t), A = buf(1, t), B = buf(2, t), C = buf(3, t);
        const bool first = t == 0;
        mbar_wait(&bars[kEbFull], uint32_t(t) & 1, 61);
        fence_after_sync();
        stamp(a, t, 0);
        if (!first) wait_epi();                                  // previous tile's last epilogue has drained the accumulator
        stamp(a, t, 1);
        chain(S, sbase + kEbWe, false);  mma_commit(&bars[kEbAcc]);                                    // 0: e We^T
        // the previous tile's dWe MMAs go behind this tile's step 0 (whose commit E0 waits for): they read C(t-1) = B(t) and S(t-1) = C(t),
        // first overwritten by E1 / E2 of this tile, i.e. after the commits of steps 1 / 2, which cover them
        if (!first) wgrad(dWe, buf(3, t - 1), buf(0, t - 1), t == 1);
        stamp(a, t, 2);
        wait_epi(); stamp(a, t, 3); chain(A, sbase + kEbW1, false);  mma_commit(&bars[kEbAcc]);                        // 1: H1 W1^T
        wait_epi(); stamp(a, t, 4); chain(B, sbase + kEbW2, false);  mma_commit(&bars[kEbAcc]);                        // 2: H2 W2^T
        // steps 3 and 4 commit after their weight-gradient MMAs: the epilogue warps spend that time on the LayerNorm vector
        // column sums anyway, and phases E3 / E4 overwrite buffers those MMAs read.  Step 5 commits right after the chain.
        wait_epi(); stamp(a, t, 5); chain(C, sbase + kEbW2, true);   wgrad(dW2, C, B, first); mma_commit(&bars[kEbAcc]);   // 3: dY W2 ; dW2
        wait_epi(); stamp(a, t, 6); chain(B, sbase + kEbW1, true);   wgrad(dW1, B, A, first); mma_commit(&bars[kEbAcc]);   // 4: dH2' W1 ; dW1
        wait_epi(); stamp(a, t, 7); chain(C, sbase + kEbWe, true);   mma_commit(&bars[kEbAcc]);                            // 5: dH1' We  (dWe: deferred)
        stamp(a, t, 8);
      }
      if (my_tiles > 0) {
        wgrad(dWe, buf(3, my_tiles - 1), buf(0, my_tiles - 1), my_tiles == 1);
        mma_commit(&bars[kEbFinal]);                           // every MMA of this CTA, for the accumulator drain
      }
    }
  } else {
    // =============================== epilogue ============================================================
    // Instruction-lean on purpose (the phases below sit on the serial critical path of a tile and the FP32 pipes issue one
    // warp instruction per two cycles): packed fp32x2 arithmetic, convert+ReLU+pack in one instruction, ReLU backward as
    // HSET2/HMUL2 against the activation words still sitting in shared memory (no mask registers), table / gradient rows
    // fetched a phase ahead, and the warp column-sum butterflies run in the shadow of the MMA steps.
    const int q = warp & 3, hh = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t acc = tmem_base + lane_addr + hh * 64;
    const float2 *b0 = reinterpret_cast<const float2*>(prm + hh * 64), *b1 = reinterpret_cast<const float2*>(prm + kD + hh * 64),
                 *b2 = reinterpret_cast<const float2*>(prm + 2 * kD + hh * 64), *gam = reinterpret_cast<const float2*>(prm + 3 * kD + hh * 64);
    uint32_t acc_phase = 0;
    float cbeta[2] = {0.f, 0.f}, cgamma[2] = {0.f, 0.f};
    auto wait_acc = [&](int tag) {
      mbar_wait(&bars[kEbAcc], acc_phase++ & 1, tag);         // try_wait (suspends): measured 2 % faster than busy polling here -- eight
                                                               // polling warps take issue slots from the column-sum warps and the MMA thread
      fence_after_sync();
    };
    // k-th column-sum hand-over of (local) tile tt: the producers have finished reading that tile from its buffer
    auto wait_cs = [&](int k, int64_t tt) { mbar_wait(&bars[kEbCs + k], uint32_t(tt) & 1, 70 + k); };
    auto done = [&](int producers_k) {   // producers_k >= 0: the tile just written is also the producers' column-sum input k
      fence_async_smem();          // generic-proxy tile writes -> visible to the tensor core's async-proxy reads
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars[kEbEpi]);
        if (producers_k >= 0) mbar_arrive(&bars[kEbG + producers_k]);
      }
    };
    const uint32_t row_off = hh * kPanel;                      // my 64 columns = panel hh of every buffer
    auto store_row = [&](uint32_t bufaddr, const uint32_t* w) {   // 64 columns (32 packed words) of row r
#pragma unroll
      for (int k = 0; k < 8; ++k) st_shared128(bufaddr + row_off + sw128_chunk(r, k), w + 4 * k);
    };
    auto load_row = [&](uint32_t bufaddr, uint32_t* w) {
#pragma unroll
      for (int k = 0; k < 8; ++k) ld_shared128(bufaddr + row_off + sw128_chunk(r, k), w + 4 * k);
    };
    auto tile_row = [&](int64_t tt) { return (blockIdx.x + tt * gridDim.x) * kTile + r; };
    auto load_tables = [&](int32_t si, int32_t ri, uint32_t* pq) {          // Ps[s] and Pr[r]: my 64 columns of each
      const __nv_bfloat16* psrow = a.proj_s + int64_t(si) * kD + hh * 64;
#pragma unroll
      for (int k = 0; k < 4; ++k) ldg256_l1(psrow + 16 * k, pq + 8 * k);
      if (a.proj_r != nullptr) {
        const __nv_bfloat16* prrow = a.proj_r + int64_t(ri) * kD + hh * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) ldg256_l1(prrow + 16 * k, pq + 32 + 8 * k);
      }
    };
    const bool ld_on = !(a.ablate & 1);
    // software pipeline across tiles: the indices of the NEXT tile are read during the current one and its table rows pulled
    // into L2 (prefetch, no registers); the register loads are issued one MMA step ahead of their use
    auto row_index = [&](const int32_t* idx, int64_t grow) -> int32_t { return idx != nullptr ? __ldg(idx + grow) : int32_t(grow); };
    int32_t si = 0, ri = 0;                                   // 32-bit until used: widening at the load would wait for it
    if (my_tiles > 0) {
      const int64_t g0 = tile_row(0);
      if (g0 < rows) { si = row_index(a.senders, g0); ri = row_index(a.receivers, g0); }
    }
    for (int64_t t = 0; t < my_tiles; ++t) {
      const uint32_t A = buf(1, t), B = buf(2, t), C = buf(3, t);
      // row-half exchange area (LayerNorm statistics): the first 4 KiB of buffer C, which is idle until dY is written into it
      float2* xch = reinterpret_cast<float2*>(smem + (C - sbase));
      const int64_t grow = tile_row(t);
      const bool valid = grow < rows;
      const int64_t ri_cur = ri;
      const int64_t gnext = tile_row(t + 1);
      const bool vnext = t + 1 < my_tiles && gnext < rows;
      const bool has_do = valid && a.grad_out != nullptr && ld_on, has_ga = valid && a.grad_agg != nullptr && ld_on;
      const __nv_bfloat16* dorow = a.grad_out + grow * kD + hh * 64;
      const __nv_bfloat16* garow = a.grad_agg + ri_cur * kD + hh * 64;
      // ---- E0: H1 = relu(e We^T + Ps[s] + Pr[r] + b0) -> A ---------------------------------------------------
      {
        uint32_t pq[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) pq[j] = 0u;
        if (ld_on) load_tables(si, ri, pq);
        if (t > 0) {                                            // the TMA store of d e has read this buffer
          if (tid == 0) { tma_store_wait_read<0>(); mbar_arrive(&bars[kEbDeFree]); }
          mbar_wait(&bars[kEbDeFree], uint32_t(t - 1) & 1, 73);
        }
        wait_acc(100); if (tid == 0) stamp(a, t, 10);
        uint32_t h[32];
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          uint32_t v[32];
          tmem_ld32(acc + cg * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float2 x = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            x = __fadd2_rn(x, unpack_bf16x2(pq[cg * 16 + j]));
            x = __fadd2_rn(x, unpack_bf16x2(pq[32 + cg * 16 + j]));
            x = __fadd2_rn(x, b0[cg * 16 + j]);
            h[cg * 16 + j] = cvt_relu_bf16x2(x.x, x.y);
          }
        }
        store_row(A, h);
        if (tid == 0) stamp(a, t, 11);
        if (lane == 0) stamp(a, t, 32 + warp);
        done(-1);
      }
      // dO = grad_out[row] + grad_agg[receiver]: requested two phases before its first use
      uint32_t dreg[32];
      {
        uint32_t dq[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) dq[j] = 0u;
        if (has_do) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ldg256(dorow + 16 * k, dq + 8 * k);
        }
        if (has_ga) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ldg256_l1(garow + 16 * k, dq + 32 + 8 * k);
        }
        // ---- E1: H2 = relu(H1 W1^T + b1) -> B -------------------------------------------------------------------
        if (t > 0) {
          wait_cs(2, t - 1);                                    // previous tile's dH1' column sum has left this buffer
        }
        wait_acc(101); if (tid == 0) stamp(a, t, 12);
        uint32_t h[32];
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          uint32_t v[32];
          tmem_ld32(acc + cg * 32, v);
          tmem_ld_wait();
          if (tid == 0) stamp(a, t, 26 + 2 * cg);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 x = __fadd2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), b1[cg * 16 + j]);
            h[cg * 16 + j] = cvt_relu_bf16x2(x.x, x.y);
          }
          if (tid == 0) stamp(a, t, 27 + 2 * cg);
        }
        store_row(B, h);
        if (tid == 0) stamp(a, t, 13);
        done(-1);
        // one rounding to bf16, the value every later use sees (EXPERIMENT: consumed after the phase's fence)
#pragma unroll
        for (int j = 0; j < 32; ++j) dreg[j] = add_bf16x2(dq[j], dq[32 + j]);
      }
      // ---- E2: y = H2 W2^T + b2 ; LayerNorm forward statistics and backward -> dY -> C --------------------------------
      uint32_t preg[32];                                        // dO * yhat (bf16): gamma-gradient terms, summed after the phase
      {
        wait_acc(102); if (tid == 0) stamp(a, t, 14);
        float2 y[32];
        {
          uint32_t v[32];
          tmem_ld32(acc, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = __fadd2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), b2[j]);
          tmem_ld32(acc + 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) y[16 + j] = __fadd2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), b2[16 + j]);
        }
        // statistics of my 64 columns (shifted by the first value), merged with the other half of the row (Chan's update)
        const float c0 = y[0].x;
        const float2 nc = make_float2(-c0, -c0);
        float2 s1a = make_float2(0.f, 0.f), s1b = s1a, s2a = s1a, s2b = s1a;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 d0 = __fadd2_rn(y[j], nc), d1 = __fadd2_rn(y[j + 1], nc);
          s1a = __fadd2_rn(s1a, d0); s1b = __fadd2_rn(s1b, d1);
          s2a = __ffma2_rn(d0, d0, s2a); s2b = __ffma2_rn(d1, d1, s2b);
        }
        const float s1 = (s1a.x + s1a.y) + (s1b.x + s1b.y), s2 = (s2a.x + s2a.y) + (s2b.x + s2b.y);
        const float mean_h = c0 + s1 * (1.0f / 64.0f);
        const float m2h = s2 - s1 * s1 * (1.0f / 64.0f);
        xch[hh * kTile + r] = make_float2(mean_h, m2h);
        epi_bar_sync();
        const float2 oth = xch[(1 - hh) * kTile + r];
        const float mean = 0.5f * (mean_h + oth.x);
        const float dm = mean_h - oth.x;
        const float rstd = rsqrtf(fmaxf(m2h + oth.y + 32.0f * dm * dm, 0.f) * (1.0f / kD) + kEps);
        const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mean * rstd, -mean * rstd);
        float2 m1a = make_float2(0.f, 0.f), m2a = m1a;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          y[j] = __ffma2_rn(y[j], rs2, nm2);                    // yhat
          const float2 d = unpack_bf16x2(dreg[j]);
          const float2 z = __fmul2_rn(d, gam[j]);
          m1a = __fadd2_rn(m1a, z);
          m2a = __ffma2_rn(z, y[j], m2a);
          const float2 p = __fmul2_rn(d, y[j]);
          preg[j] = pack_bf16(p.x, p.y);
        }
        xch[2 * kTile + hh * kTile + r] = make_float2(m1a.x + m1a.y, m2a.x + m2a.y);
        epi_bar_sync();
        const float2 o2 = xch[2 * kTile + (1 - hh) * kTile + r];
        const float m1 = (m1a.x + m1a.y + o2.x) * (1.0f / kD), m2 = (m2a.x + m2a.y + o2.y) * (1.0f / kD);
        const float2 nm1 = make_float2(-m1, -m1), nmm2 = make_float2(-m2, -m2);
        uint32_t o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float2 z = __fmul2_rn(unpack_bf16x2(dreg[j]), gam[j]);
          const float2 u = __fmul2_rn(__ffma2_rn(y[j], nmm2, __fadd2_rn(z, nm1)), rs2);     // rstd (dO gamma - m1 - yhat m2)
          o[j] = pack_bf16(u.x, u.y);
        }
        epi_bar_sync();                                         // every thread has read both exchanges: dY may overwrite them
        store_row(C, o);
        if (tid == 0) stamp(a, t, 15);
        done(0);
      }
      // (in the shadow of MMA step 3) gamma gradient: column sums of dO * yhat over this warp's 32 rows
      if (!(a.ablate & 4)) {
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          float p[32];
#pragma unroll
          for (int j = 0; j < 16; ++j) { p[2 * j] = bf16_lo(preg[cg * 16 + j]); p[2 * j + 1] = bf16_hi(preg[cg * 16 + j]); }
          cgamma[cg] += warp_colsum32(p, lane);
        }
      }
      // ---- E3: dH2' = (dY W2) * [H2 > 0] -> B ---------------------------------------------------------------------
      {
        wait_acc(103); if (tid == 0) stamp(a, t, 16);
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {                        // per 32-column group: mask words (H2, about to be replaced by its
          uint32_t v[32], hw[16], o[16];                        // own gradient) are fetched under the TMEM load
          tmem_ld32(acc + cg * 32, v);
#pragma unroll
          for (int k = 0; k < 4; ++k) ld_shared128(B + row_off + sw128_chunk(r, 4 * cg + k), hw + 4 * k);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            o[j] = relu_bwd_bf16x2(pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), hw[j]);
#pragma unroll
          for (int k = 0; k < 4; ++k) st_shared128(B + row_off + sw128_chunk(r, 4 * cg + k), o + 4 * k);
        }
        if (tid == 0) stamp(a, t, 17);
        done(1);
      }
      // (in the shadow of MMA step 4) beta gradient: column sums of dO
      if (!(a.ablate & 4)) {
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          float p[32];
#pragma unroll
          for (int j = 0; j < 16; ++j) { p[2 * j] = bf16_lo(dreg[cg * 16 + j]); p[2 * j + 1] = bf16_hi(dreg[cg * 16 + j]); }
          cbeta[cg] += warp_colsum32(p, lane);
        }
      }
      // ---- E4: G0 = dH1' = (dH2' W1) * [H1 > 0] -> C and -> HBM -----------------------------------------------------------
      {
        wait_cs(0, t);                                          // the producers' dY column sum has left buffer C
        wait_acc(104); if (tid == 0) stamp(a, t, 18);
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          uint32_t v[32], hw[16], o[16];
          tmem_ld32(acc + cg * 32, v);
#pragma unroll
          for (int k = 0; k < 4; ++k) ld_shared128(A + row_off + sw128_chunk(r, 4 * cg + k), hw + 4 * k);   // H1 mask words
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            o[j] = relu_bwd_bf16x2(pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), hw[j]);
#pragma unroll
          for (int k = 0; k < 4; ++k) st_shared128(C + row_off + sw128_chunk(r, 4 * cg + k), o + 4 * k);
        }
        __syncwarp(); if (lane == 0) mbar_arrive(&bars[kEbAfree]);   // this warp has read its H1 rows: the last reader of buffer A in this tile
        if (tid == 0) stamp(a, t, 19);
        if (lane == 0) stamp(a, t, 40 + warp);
        done(2);
        // indices of the next tile's row (an L2 hit: the producers touched them a tile ago); first used by the next E0
        si = 0; ri = 0;
        if (vnext) { si = row_index(a.senders, gnext); ri = row_index(a.receivers, gnext); }
      }
      // ---- E5: d e = dH1' We + dO -> HBM -------------------------------------------------------------------------------
      {
        wait_acc(105); if (tid == 0) stamp(a, t, 20);
        uint32_t v0[32], v1[32];
        tmem_ld32(acc, v0);
        tmem_ld32(acc + 32, v1);
        tmem_ld_wait();
        // the accumulator is in registers: the next tile's step 0 may overwrite it while this phase does its arithmetic and
        // stores (no shared-memory writes here, so no proxy fence)
        fence_before_sync();
        __syncwarp(); if (lane == 0) mbar_arrive(&bars[kEbEpi]);
        uint32_t o[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 x0 = __fadd2_rn(make_float2(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1])), unpack_bf16x2(dreg[j]));
          const float2 x1 = __fadd2_rn(make_float2(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1])), unpack_bf16x2(dreg[16 + j]));
          o[j] = pack_bf16(x0.x, x0.y);
          o[16 + j] = pack_bf16(x1.x, x1.y);
        }
        // staged in buffer B (dH2' is dead: the chain commit above came after the dW1 MMAs, and the producers' column sum of it,
        // long finished, is awaited here rather than assumed) and stored by TMA
        wait_cs(1, t);
        store_row(B, o);
        fence_async_smem();
        __syncwarp(); if (lane == 0) mbar_arrive(&bars[kEbDe]);
        if (tid == 0) {                                         // one thread stores the staged tile once all 256 rows are in
          mbar_wait(&bars[kEbDe], uint32_t(t) & 1, 77);
          if (!(a.ablate & 2)) {
            const int y = int((blockIdx.x + t * gridDim.x) * kTile);
            tma_store_2d(&tm_de, B, 0, y);
            tma_store_2d(&tm_de, B + kPanel, 64, y);
            tma_store_commit();
          }
          stamp(a, t, 21);
        }
      }
    }
    if (tid == 0) tma_store_wait<0>();
    // ---- drain the weight-gradient accumulators and the LayerNorm vector partials ---------------------------------------
    if (my_tiles > 0) {
      mbar_wait(&bars[kEbFinal], 0, 75);
      fence_after_sync();
    }
#pragma unroll 1
    for (int z = 0; z < 3; ++z) {
      const uint32_t col0 = 384u - 128u * uint32_t(z);           // z = 0: dWe, 1: dW1, 2: dW2
      float* dst = a.w_partial + ((int64_t(blockIdx.x) * 3 + z) * kD + r) * kD + hh * 64;
#pragma unroll 1
      for (int cg = 0; cg < 2; ++cg) {
        uint32_t v[32];
        if (my_tiles > 0) {
          tmem_ld32(tmem_base + lane_addr + col0 + hh * 64 + cg * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4*>(dst + cg * 32 + 4 * k) = make_float4(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]),
                                                                          __uint_as_float(v[4 * k + 2]), __uint_as_float(v[4 * k + 3]));
      }
    }
#pragma unroll
    for (int cg = 0; cg < 2; ++cg) {
      float* cp = a.epi_colpart + ((int64_t(blockIdx.x) * 4 + q) * 2) * kD + hh * 64 + cg * 32 + lane;
      cp[0] = cbeta[cg];
      cp[kD] = cgamma[cg];
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 10) tmem_dealloc<512>(tmem_base);
}

// fixed-order reduction of the per-CTA partials of edge_bwd_tc_kernel
__global__ void __launch_bounds__(256)
edge_bwd_reduce_kernel(const float* __restrict__ w_partial, const float* __restrict__ epi_colpart,
                       const float* __restrict__ prod_colpart, int parts, int w0_chunks, int w0_chunk0, int skip_we, float* __restrict__ gW0, float* __restrict__ gW1,
                       float* __restrict__ gW2, float* gb0, float* gb1, float* gb2, float* ggamma, float* gbeta) {
  __shared__ float sm[256];
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);         // 32 outputs per block (common.cuh, ordered_sum_block8)
  const float* src = nullptr;
  int n = parts;
  int64_t stride = 0;
  if (i < 3 * kD * kD) {
    if (!(skip_we && i < kD * kD)) { src = w_partial + i; stride = int64_t(3) * kD * kD; }    // dWe may come from the streaming weight-gradient kernel
  } else if (i < 3 * kD * kD + 5 * kD) {
    const int k = (i - 3 * kD * kD) / kD, c = i % kD;
    if (k < 3) { src = prod_colpart + int64_t(k) * kD + c; stride = int64_t(3) * kD; }
    else { src = epi_colpart + int64_t(k - 3) * kD + c; n = parts * 4; stride = int64_t(2) * kD; }
  }
  const float s = ordered_sum_block8(src, n, stride, sm);
  if (threadIdx.x >= 32 || src == nullptr) return;
  if (i < 3 * kD * kD) {
    const int z = i / (kD * kD), o = (i / kD) % kD, c = i % kD;
    if (z == 0) gW0[(int64_t(o) * w0_chunks + w0_chunk0) * kD + c] = s;
    else if (z == 1) gW1[o * kD + c] = s;
    else gW2[o * kD + c] = s;
  } else {
    const int k = (i - 3 * kD * kD) / kD, c = i % kD;
    (k == 0 ? gb2 : k == 1 ? gb1 : k == 2 ? gb0 : k == 3 ? gbeta : ggamma)[c] = s;
  }
}

// gW0[:, col0 : col0+128] = sum_p partial[p][z]   (partials of the pair weight-gradient kernel, [parts][n_z][128][128])
__global__ void __launch_bounds__(256)
reduce_w0_block_kernel(const float* __restrict__ partial, int parts, int n_z, int z, float* __restrict__ gW0, int ld, int col0) {
  __shared__ float sm[256];
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);
  const float s = ordered_sum_block8(i < kD * kD ? partial + int64_t(z) * kD * kD + i : nullptr, parts, int64_t(n_z) * kD * kD, sm);
  if (threadIdx.x < 32 && i < kD * kD) gW0[int64_t(i / kD) * ld + col0 + (i % kD)] = s;
}

// ---- host side -----------------------------------------------------------------------------------------------------------
void launch_edge_bwd_reduce(const float* w_partial, const float* epi_colpart, const float* prod_colpart, int parts, int w0_chunks, int w0_chunk0,
                            float* gW0, float* gW1, float* gW2, float* gb0, float* gb1, float* gb2, float* ggamma, float* gbeta, cudaStream_t st) {
  edge_bwd_reduce_kernel<<<(3 * kD * kD + 5 * kD + 31) / 32, 256, 0, st>>>(w_partial, epi_colpart, prod_colpart, parts, w0_chunks, w0_chunk0, 0,
                                                                            gW0, gW1, gW2, gb0, gb1, gb2, ggamma, gbeta);
}

int tc_pair_wgrad(int64_t rows, const void* Ga, const void* Za, const void* Gb, const void* Zb, float* partial, int parts,
                  cudaStream_t st);                              // mlp_tc.cu
int tc_pair_wgrad_parts(int64_t rows);

static int configure_edge_kernels() {
  static bool configured = false;
  if (!configured) {
    uint32_t* dbg = debug_buffer_device();
    HGN_CUDA_OK(cudaMemcpyToSymbol(tc05::g_debug_words, &dbg, sizeof(dbg)));
    HGN_CUDA_OK(cudaFuncSetAttribute(edge_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kEbSmem)));
    HGN_CUDA_OK(cudaFuncSetAttribute(proj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kChunkBytes + 4 * kChunkBytes + 128));
    configured = true;
  }
  return HGN_OK;
}

static int launch_proj(int64_t rows, const void* packed, const LinArgs& la, const char* name, cudaStream_t st) {
  if (int rc = configure_edge_kernels()) return rc;
  const int64_t tiles = ceil_div(rows, kTile);
  const unsigned grid = unsigned(tiles < tc_sm_count() ? tiles : tc_sm_count());
  const size_t smem = size_t(2 + 2 * la.n_in) * kChunkBytes + 128;
  HGN_TIMED(name, st);
  proj_tc_kernel<<<grid, kLinThreads, smem, st>>>(rows, tiles, static_cast<const uint8_t*>(packed), la);
  HGN_LAUNCH_OK(name);
  return HGN_OK;
}

int edge_project_forward_tc(int64_t num_nodes, const void* v, const void* packed, void* proj_s, void* proj_r, cudaStream_t st) {
  LinArgs la{};
  la.in[0] = static_cast<const __nv_bfloat16*>(v);
  la.out[0] = static_cast<__nv_bfloat16*>(proj_s);
  la.out[1] = static_cast<__nv_bfloat16*>(proj_r);
  la.n_in = 1; la.n_out = 2; la.b_mn = 0; la.w0_chunks = 3; la.chunk0 = 0;
  return launch_proj(num_nodes, packed, la, "edge_project_fwd", st);
}

size_t edge_project_backward_workspace_tc(int64_t num_nodes) {
  return size_t(tc_pair_wgrad_parts(num_nodes)) * 2 * kD * kD * sizeof(float) + 256;
}

int edge_project_backward_tc(int64_t num_nodes, const void* v, const void* packed, const void* grad_s, const void* grad_r, const void* grad_v_add,
                             void* grad_v, float* grad_W0, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (workspace_bytes < edge_project_backward_workspace_tc(num_nodes)) { set_error("edge_project_backward: workspace too small"); return HGN_ERR_WORKSPACE; }
  LinArgs la{};
  la.in[0] = static_cast<const __nv_bfloat16*>(grad_s);
  la.in[1] = static_cast<const __nv_bfloat16*>(grad_r);
  la.out[0] = static_cast<__nv_bfloat16*>(grad_v);
  la.add = static_cast<const __nv_bfloat16*>(grad_v_add);
  la.n_in = 2; la.n_out = 1; la.b_mn = 1; la.w0_chunks = 3; la.chunk0 = 0;
  if (int rc = launch_proj(num_nodes, packed, la, "edge_project_dgrad", st)) return rc;
  float* partial = static_cast<float*>(workspace);
  const int parts = tc_pair_wgrad_parts(num_nodes);
  if (int rc = tc_pair_wgrad(num_nodes, grad_s, v, grad_r, v, partial, parts, st)) return rc;
  {
    HGN_TIMED("reduce_weight_partials", st);
    reduce_w0_block_kernel<<<kD * kD / 32, 256, 0, st>>>(partial, parts, 2, 1, grad_W0, 3 * kD, 0);        // Gs^T v -> columns 0:128
    reduce_w0_block_kernel<<<kD * kD / 32, 256, 0, st>>>(partial, parts, 2, 0, grad_W0, 3 * kD, kD);       // Gr^T v -> columns 128:256
  }
  HGN_LAUNCH_OK("edge_project_backward reductions");
  return HGN_OK;
}

int make_rows_tensor_map(CUtensorMap* tm, const void* base, int64_t rows);   // edge_fwd_tc.cu
int edge_fwd_tc_launch(int64_t rows, const void* dense, const void* proj_s, const void* proj_r, const int32_t* senders, const int32_t* receivers,
                       const void* packed, int w0_chunks, int w0_chunk0, void* out, const char* name, cudaStream_t st);                                                                                  // edge_fwd_tc.cu
int edge_update_forward_tc(int64_t num_edges, const void* edge, const void* proj_s, const void* proj_r, const int32_t* senders,
                           const int32_t* receivers, const void* packed, void* out, cudaStream_t st) {
  return edge_fwd_tc_launch(num_edges, edge, proj_s, proj_r, senders, receivers, packed, 3, 2, out, "edge_fwd_tc", st);
}

struct EdgeBwdLayout { size_t w_partial, epi, prod, total; int grid; };
static EdgeBwdLayout edge_bwd_layout(int64_t rows) {
  EdgeBwdLayout L{};
  const int64_t tiles = ceil_div(rows > 0 ? rows : 1, kTile);
  L.grid = int(tiles < tc_sm_count() ? tiles : tc_sm_count());
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.w_partial = take(size_t(L.grid) * 3 * kD * kD * 4);
  L.epi = take(size_t(L.grid) * 4 * 2 * kD * 4);
  L.prod = take(size_t(L.grid) * 3 * kD * 4);
  L.total = off;
  return L;
}

size_t edge_update_backward_workspace_tc(int64_t num_edges) { return edge_bwd_layout(num_edges).total; }

// Backward of  out = dense + LN(MLP(We dense + Ps[s] + Pr[r] + b0 ...)):  grad_dense, grad_pre0 (= G0) and every weight gradient
// except the W0 blocks that multiply the node tables (those follow from G0 at node level).  gW0's block w0_chunk0 is written.
static int projected_backward_launch(int64_t rows, const void* dense, const void* proj_s, const void* proj_r, const int32_t* senders,
                                     const int32_t* receivers, const void* packed, int w0_chunks, int w0_chunk0, const void* grad_out,
                                     const void* grad_agg, void* grad_dense, void* grad_pre0, float* gW0, float* gb0, float* gW1, float* gb1,
                                     float* gW2, float* gb2, float* ggamma, float* gbeta, void* workspace, const char* name, cudaStream_t st) {
  const EdgeBwdLayout L = edge_bwd_layout(rows);
  if (int rc = configure_edge_kernels()) return rc;
  char* ws = static_cast<char*>(workspace);
  EdgeBwdArgs a{};
  a.edge = static_cast<const __nv_bfloat16*>(dense);
  a.proj_s = static_cast<const __nv_bfloat16*>(proj_s);
  a.proj_r = static_cast<const __nv_bfloat16*>(proj_r);
  a.senders = senders;
  a.receivers = receivers;
  a.grad_out = static_cast<const __nv_bfloat16*>(grad_out);
  a.grad_agg = static_cast<const __nv_bfloat16*>(grad_agg);
  a.grad_edge = static_cast<__nv_bfloat16*>(grad_dense);
  a.grad_pre0 = static_cast<__nv_bfloat16*>(grad_pre0);
  a.w_partial = reinterpret_cast<float*>(ws + L.w_partial);
  a.epi_colpart = reinterpret_cast<float*>(ws + L.epi);
  a.prod_colpart = reinterpret_cast<float*>(ws + L.prod);
  a.w0_chunks = w0_chunks;
  a.w0_chunk0 = w0_chunk0;
  { const char* ab = getenv("HGN_TC_ABLATE"); a.ablate = ab ? atoi(ab) : 0; }
  static long long* tl_dev = nullptr;
  if (a.ablate & 64) {
    if (tl_dev == nullptr) cudaMalloc(&tl_dev, 8 * 48 * sizeof(long long));
    cudaMemsetAsync(tl_dev, 0, 8 * 48 * sizeof(long long), st);
    a.timeline = tl_dev;
  }
  const int64_t tiles = ceil_div(rows, kTile);
  CUtensorMap tm_e, tm_g0, tm_de;                       // rows == 0: maps over one (never accessed) row keep the encoder happy
  if (int rc = make_rows_tensor_map(&tm_e, dense != nullptr ? dense : workspace, rows > 0 ? rows : 1)) return rc;
  if (int rc = make_rows_tensor_map(&tm_g0, grad_pre0 != nullptr ? grad_pre0 : workspace, rows > 0 ? rows : 1)) return rc;
  if (int rc = make_rows_tensor_map(&tm_de, grad_dense != nullptr ? grad_dense : workspace, rows > 0 ? rows : 1)) return rc;
  {
    HGN_TIMED(name, st);
    edge_bwd_tc_kernel<<<unsigned(L.grid), kEbThreads, kEbSmem, st>>>(rows, tiles, static_cast<const uint8_t*>(packed), a, tm_e, tm_g0, tm_de);
  }
  HGN_LAUNCH_OK(name);
  if (a.timeline != nullptr) {
    long long h[8 * 48];
    cudaMemcpyAsync(h, a.timeline, sizeof(h), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    for (int t = 0; t < 8; ++t) {
      fprintf(stderr, "tile %d:", t);
      for (int k = 0; k < 48; ++k) fprintf(stderr, " %lld", h[t * 48 + k] ? h[t * 48 + k] - h[0] : -1);
      fprintf(stderr, "\n");
    }
  }
  {
    HGN_TIMED("reduce_weight_partials", st);
    edge_bwd_reduce_kernel<<<(3 * kD * kD + 5 * kD + 31) / 32, 256, 0, st>>>(a.w_partial, a.epi_colpart, a.prod_colpart, L.grid, a.w0_chunks, a.w0_chunk0, 0,
                                                                              gW0, gW1, gW2, gb0, gb1, gb2, ggamma, gbeta);
  }
  HGN_LAUNCH_OK("edge_bwd_reduce");
  return HGN_OK;
}

int edge_update_backward_tc(int64_t num_edges, const void* edge, const void* proj_s, const void* proj_r, const int32_t* senders,
                            const int32_t* receivers, const void* packed, const void* grad_out,
                            const void* grad_agg, void* grad_edge, void* grad_pre0, float* gW0, float* gb0, float* gW1, float* gb1, float* gW2,
                            float* gb2, float* ggamma, float* gbeta, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const EdgeBwdLayout L = edge_bwd_layout(num_edges);
  if (workspace_bytes < L.total) { set_error("edge_update_backward: workspace %zu < %zu", workspace_bytes, L.total); return HGN_ERR_WORKSPACE; }
  return projected_backward_launch(num_edges, edge, proj_s, proj_r, senders, receivers, packed, 3, 2, grad_out, grad_agg, grad_edge, grad_pre0,
                                   gW0, gb0, gW1, gb1, gW2, gb2, ggamma, gbeta, workspace, "edge_bwd_tc", st);
}

// ---- node update in 'sum' mode through the same kernels (graphnet.py:34-48 with one aggregate) ---------------------------
//   v' = v + LN(MLP([v | agg]))  with  W0 = [Wv | Wa]:  pre0 = Wv v + Q + b0,  Q = agg Wa^T  (one projection per node row, gathered
//   through the identity), so the fused forward / backward kernels of the edge update serve the node update unchanged.
// n_agg = 1..4 aggregates (one 'sum', or sum/mean/max/min of 'pna'): W0 = [Wv | Wa_1 | ... | Wa_n].  The aggregates' blocks of the
// first linear are applied per node row into at most two tables, q1 = agg_1 Wa_1^T (+ agg_2 Wa_2^T) and q2 = agg_3 Wa_3^T (+ agg_4
// Wa_4^T), which the fused kernel gathers through the identity in the places of Ps[s] and Pr[r].
int node_update_forward_tc(int64_t num_nodes, const void* v, int n_agg, const void* const* aggs, const void* packed, void* q1, void* q2,
                           void* out, cudaStream_t st) {
  for (int t = 0; t * 2 < n_agg; ++t) {
    LinArgs la{};
    la.n_in = n_agg - 2 * t >= 2 ? 2 : 1;
    for (int c = 0; c < la.n_in; ++c) la.in[c] = static_cast<const __nv_bfloat16*>(aggs[2 * t + c]);
    la.out[0] = static_cast<__nv_bfloat16*>(t == 0 ? q1 : q2);
    la.n_out = 1; la.b_mn = 0; la.w0_chunks = 1 + n_agg; la.chunk0 = 1 + 2 * t;
    if (int rc = launch_proj(num_nodes, packed, la, "node_project_fwd", st)) return rc;
  }
  return edge_fwd_tc_launch(num_nodes, v, q1, n_agg > 2 ? q2 : nullptr, nullptr, nullptr, packed, 1 + n_agg, 0, out, "node_fwd_tc", st);
}

struct NodeBwdLayout { size_t edge, g0, partial, total; int parts; };
static NodeBwdLayout node_bwd_layout(int64_t n) {
  NodeBwdLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.parts = tc_pair_wgrad_parts(n);
  L.edge = take(edge_bwd_layout(n).total);
  L.g0 = take(size_t(n > 0 ? n : 1) * kD * 2);
  L.partial = take(size_t(L.parts) * 2 * kD * kD * 4);
  L.total = off;
  return L;
}
size_t node_update_backward_workspace_tc(int64_t num_nodes) { return node_bwd_layout(num_nodes).total; }

int node_update_backward_tc(int64_t num_nodes, const void* v, int n_agg, const void* const* aggs, const void* q1, const void* q2,
                            const void* packed, const void* grad_out, void* grad_v, void* const* grad_aggs,
                            float* gW0, float* gb0, float* gW1, float* gb1, float* gW2, float* gb2, float* ggamma, float* gbeta,
                            void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const NodeBwdLayout L = node_bwd_layout(num_nodes);
  if (workspace_bytes < L.total) { set_error("node_update_backward: workspace %zu < %zu", workspace_bytes, L.total); return HGN_ERR_WORKSPACE; }
  char* ws = static_cast<char*>(workspace);
  void* g0 = ws + L.g0;
  const int w0_chunks = 1 + n_agg;
  if (int rc = projected_backward_launch(num_nodes, v, q1, n_agg > 2 ? q2 : nullptr, nullptr, nullptr, packed, w0_chunks, 0, grad_out, nullptr,
                                                grad_v, g0, gW0, gb0, gW1, gb1, gW2, gb2, ggamma, gbeta, ws + L.edge, "node_bwd_tc", st)) {
    return rc;
  }
  float* partial = reinterpret_cast<float*>(ws + L.partial);
  for (int t = 0; t * 2 < n_agg; ++t) {
    const int n = n_agg - 2 * t >= 2 ? 2 : 1;
    LinArgs la{};                                // d agg_j = G0 Wa_j for the (up to) two aggregates of table t
    la.in[0] = static_cast<const __nv_bfloat16*>(g0);
    for (int o = 0; o < n; ++o) la.out[o] = static_cast<__nv_bfloat16*>(grad_aggs[2 * t + o]);
    la.n_in = 1; la.n_out = n; la.b_mn = 1; la.w0_chunks = w0_chunks; la.chunk0 = 1 + 2 * t;
    if (int rc = launch_proj(num_nodes, packed, la, "node_project_dgrad", st)) return rc;
    // d Wa_j = G0^T agg_j  (pair kernel: partial z = 1 holds the first pair, z = 0 the second)
    if (int rc = tc_pair_wgrad(num_nodes, g0, aggs[2 * t], n == 2 ? g0 : nullptr, n == 2 ? aggs[2 * t + 1] : nullptr, partial, L.parts, st)) return rc;
    {
      HGN_TIMED("reduce_weight_partials", st);
      reduce_w0_block_kernel<<<kD * kD / 32, 256, 0, st>>>(partial, L.parts, 2, 1, gW0, w0_chunks * kD, (1 + 2 * t) * kD);
      if (n == 2) reduce_w0_block_kernel<<<kD * kD / 32, 256, 0, st>>>(partial, L.parts, 2, 0, gW0, w0_chunks * kD, (2 + 2 * t) * kD);
    }
    HGN_LAUNCH_OK("node_update_backward reductions");
  }
  return HGN_OK;
}

}  // namespace hgn
