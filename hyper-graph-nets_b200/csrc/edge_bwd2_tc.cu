// bf16 tcgen05 backward of the projected edge / node update, two tiles interleaved (HGN_BF16 only).
//
// Same arithmetic as edge_tc.cu's edge_bwd_tc_kernel (recompute L0 L1 L2, LayerNorm backward, dgrad chain, weight gradients):
//     step/phase 0: H1 = relu(e We^T + Ps[s] + Pr[r] + b0)     3: dH2' = (dY W2)  * [H2 > 0]     dW2 += dY^T H2
//                1: H2 = relu(H1 W1^T + b1)                     4: G0   = (dH2' W1) * [H1 > 0]     dW1 += dH2'^T H1
//                2: y = H2 W2^T + b2, dY = LN'(y; dO)           5: d e  = G0 We + dO
// The one-tile kernel runs a tile's six GEMM -> epilogue hand-overs strictly one after the other, so the tensor pipe idles
// during every epilogue phase and the epilogue warps during every GEMM (23 % tensor pipe, 34 % issue slots).  Here tile A
// (= local tile k, in its forward half: phases 0 1 2) and tile B (= k - 1, in its backward half: 3 4 5) alternate:
//     epilogue warps:  E3(B) E0(A) E4(B) E1(A) E5(B) E2(A) | next round ...
//     MMA issuer:      ... c3(B)+w3(B) | c4(B)+w4(B) | c1(A) | c5(B) | c2(A) | c0(next A) | ...     (c = chain GEMM, w = weight gradient)
// so each GEMM runs under the other tile's epilogue phase.  What makes that fit:
//   * shared memory: three weight blocks + FOUR tile buffers as before.  Tile t uses buffer t&3 for e, then H1, then its d e
//     staging tile; (t+2)&3 for H2, then dH2', then G0; (t+3)&3 for dY.  Every hand-over between the two tiles in flight is a
//     buffer whose last reader (a weight-gradient GEMM or a column sum) has provably finished (see the waits below).
//   * TMEM: two chain accumulators (one per tile in flight) + dW2 + dW1 = 512 columns.  dWe = G0^T e, the one weight gradient
//     whose two operands live in HBM anyway, is left to the streaming weight-gradient kernel (mlp_tc.cu) after this one.
//   * registers: no per-tile state survives a phase except the receiver index; dO is gathered again for phase 5 (an L2 hit) and
//     the LayerNorm vector column sums are taken right after phase 2.
// Roles: warps 0-7 epilogue (row = TMEM lane, 64 columns per thread), warps 8-9 bias column sums out of the operand buffers +
// L2 prefetch + (thread 0 of warp 8) all TMA traffic, warp 10 MMA issuer.
#include <cuda.h>
#include <stdlib.h>

#include "tile_common.cuh"

namespace hgn {

constexpr int kB2EpiThreads = 256, kB2ProdThreads = 64;
constexpr int kB2Threads = kB2EpiThreads + kB2ProdThreads + 32;    // 11 warps
constexpr uint32_t kB2We = 0, kB2W1 = kChunkBytes, kB2W2 = 2 * kChunkBytes, kB2Buf = 3 * kChunkBytes;
constexpr uint32_t kB2Params = kB2Buf + 4 * kChunkBytes;           // b0 b1 b2 gamma beta (fp32 x 128 each)
constexpr uint32_t kB2Bars = kB2Params + 5 * kD * 4;
constexpr uint32_t kB2Smem = kB2Bars + 256;                        // 32 barrier slots
// kB2G + j / kB2Cs + j (j = 0 dY, 1 dH2', 2 G0): tile written / its bias column sum (and, for G0, its TMA store) done
// kB2Full + (t & 1): edge rows of tile t landed (two tile loads can be in flight).  kB2Epi + (i & 1): epilogue phase i done --
// two barriers in alternation, so that a parity wait could only alias if the epilogue got FOUR phases ahead of the MMA thread
// (it cannot: every phase needs a GEMM that the MMA thread issues only after observing the phase before the previous one).
enum { kB2Full = 0, kB2Acc = 2, kB2Epi = 4, kB2G = 6, kB2Cs = 9, kB2Tmem = 12, kB2De = 13, kB2DeFree = 14, kB2Sfree = 15, kB2Final = 16 };

struct Bwd2Args {
  const __nv_bfloat16 *proj_s, *proj_r;
  const int32_t *senders, *receivers;
  const __nv_bfloat16* grad_out;    // [rows,128] dense part of d loss / d out (may be null)
  const __nv_bfloat16* grad_agg;    // [N,128] gathered through receivers (may be null)
  float* w_partial;                 // [grid][3][128][128]  z = 1: dW1, 2: dW2 (z = 0 unused here)
  float* epi_colpart;               // [grid][4][2][128]    beta, gamma partial column sums per lane quadrant
  float* prod_colpart;              // [grid][3][128]       db2, db1, db0
  int w0_chunks, w0_chunk0;
  long long* timeline;              // development (HGN_TC_ABLATE bit 64): clock64 stamps of block 0, [round][32]
};

__global__ void __launch_bounds__(kB2Threads, 1)
edge_bwd2_tc_kernel(int64_t rows, int64_t num_tiles, const uint8_t* __restrict__ packed, Bwd2Args a,
                    const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_g0,
                    const __grid_constant__ CUtensorMap tm_de) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kB2Bars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PackedTc P(a.w0_chunks);
  float* prm = reinterpret_cast<float*>(smem + kB2Params);
  {
    const float* pg = reinterpret_cast<const float*>(packed + P.params);
    for (int i = tid; i < 5 * kD; i += kB2Threads) prm[i] = pg[i];
    const __nv_bfloat16* w0g = reinterpret_cast<const __nv_bfloat16*>(packed + P.w0);
    load_weight_block(sbase + kB2We, w0g + a.w0_chunk0 * kD, int64_t(a.w0_chunks) * kD, tid, kB2Threads);
    load_weight_block(sbase + kB2W1, reinterpret_cast<const __nv_bfloat16*>(packed + P.w1), kD, tid, kB2Threads);
    load_weight_block(sbase + kB2W2, reinterpret_cast<const __nv_bfloat16*>(packed + P.w2), kD, tid, kB2Threads);
    cp_async_commit();
    if (tid == 0) {
      mbar_init(&bars[kB2Full], 1); mbar_init(&bars[kB2Full + 1], 1);
      mbar_init(&bars[kB2Acc], 1); mbar_init(&bars[kB2Acc + 1], 1);
      mbar_init(&bars[kB2Epi], kB2EpiThreads / 32); mbar_init(&bars[kB2Epi + 1], kB2EpiThreads / 32);   // one arrival per epilogue warp
      for (int j = 0; j < 3; ++j) { mbar_init(&bars[kB2G + j], kB2EpiThreads / 32); mbar_init(&bars[kB2Cs + j], kB2ProdThreads / 32); }
      mbar_init(&bars[kB2De], kB2EpiThreads / 32);
      mbar_init(&bars[kB2DeFree], 1);
      mbar_init(&bars[kB2Sfree], 1);
      mbar_init(&bars[kB2Final], 1);
      mbar_init_fence();
    }
    if (warp == 10) tmem_alloc<512>(reinterpret_cast<uint32_t*>(&bars[kB2Tmem]));
    cp_async_wait<0>();
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars[kB2Tmem]);
  const int64_t T = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;       // my tiles: local index t -> global tile b + t G
  auto buf = [&](int64_t i) -> uint32_t { return sbase + kB2Buf + uint32_t(i & 3) * kChunkBytes; };
  auto buf_h1 = [&](int64_t t) { return buf(t); };           // e, then H1, then the d e staging tile
  auto buf_h2 = [&](int64_t t) { return buf(t + 2); };       // H2, then dH2', then G0
  auto buf_dy = [&](int64_t t) { return buf(t + 3); };       // dY
  auto tile_y = [&](int64_t t) -> int { return int((blockIdx.x + t * gridDim.x) * kTile); };
  auto stamp = [&](int64_t k, int slot) { if (a.timeline != nullptr && blockIdx.x == 0 && k < 10) a.timeline[k * 32 + slot] = clock64(); };

  if (warp == 8 || warp == 9) {
    // =============================== producers =========================================================
    const int ptid = tid - kB2EpiThreads, pw = warp - 8;
    auto load_e = [&](int64_t t) {
      uint64_t* full = &bars[kB2Full + int(t & 1)];
      mbar_expect_tx(full, kChunkBytes);
      tma_load_2d(buf_h1(t), &tm_e, 0, tile_y(t), full);
      tma_load_2d(buf_h1(t) + kPanel, &tm_e, 64, tile_y(t), full);
    };
    auto store_tile = [&](const CUtensorMap* tm, uint32_t src, int64_t t) {
      tma_store_2d(tm, src, 0, tile_y(t));
      tma_store_2d(tm, src + kPanel, 64, tile_y(t));
      tma_store_commit();
    };
    float cs[3][8];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[k][j] = 0.f;
    // column sums of a bf16 tile in a buffer: warp pw owns panel pw; lane l reads the 16-byte piece l & 7 of rows 4 i + (l >> 3)
    auto colsum = [&](uint32_t base, float (&acc8)[8]) {
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
    // L2 prefetch of everything a tile gathers from HBM.  The tile's indices are read one call earlier (fetch_idx) so that the
    // prefetch addresses never wait on an index load.
    int32_t pf_s[2] = {0, 0}, pf_r[2] = {0, 0};
    auto fetch_idx = [&](int64_t t) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int64_t grow = int64_t(tile_y(t)) + ptid + 64 * j;
        const bool ok = t < T && grow < rows;
        pf_s[j] = ok ? (a.senders != nullptr ? __ldg(a.senders + grow) : int32_t(grow)) : -1;
        pf_r[j] = ok ? (a.receivers != nullptr ? __ldg(a.receivers + grow) : int32_t(grow)) : -1;
      }
    };
    auto prefetch_tile = [&](int64_t t) {     // uses the indices fetched for tile t
      auto pf = [](const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); };
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (pf_s[j] < 0) continue;
        const int64_t grow = int64_t(tile_y(t)) + ptid + 64 * j, si = pf_s[j], ri = pf_r[j];
        if (a.grad_out != nullptr) { pf(a.grad_out + grow * kD); pf(a.grad_out + grow * kD + 64); }
        pf(a.proj_s + si * kD); pf(a.proj_s + si * kD + 64);
        if (a.proj_r != nullptr) { pf(a.proj_r + ri * kD); pf(a.proj_r + ri * kD + 64); }
        if (a.grad_agg != nullptr) { pf(a.grad_agg + ri * kD); pf(a.grad_agg + ri * kD + 64); }
      }
    };
    if (ptid == 0) {
      if (T > 0) load_e(0);
      if (T > 1) load_e(1);
    }
    fetch_idx(0); prefetch_tile(0);
    fetch_idx(1); prefetch_tile(1);
    fetch_idx(2);
    for (int64_t t = 0; t < T; ++t) {
      const uint32_t par = uint32_t(t) & 1;
      prefetch_tile(t + 2);                                  // indices read during the previous iteration
      fetch_idx(t + 3);
      mbar_wait(&bars[kB2G + 0], par, 50);
      colsum(buf_dy(t), cs[0]);                              // dY
      __syncwarp(); if (lane == 0) mbar_arrive(&bars[kB2Cs + 0]);
      mbar_wait(&bars[kB2G + 1], par, 51);
      colsum(buf_h2(t), cs[1]);                              // dH2'
      __syncwarp(); if (lane == 0) mbar_arrive(&bars[kB2Cs + 1]);
      mbar_wait(&bars[kB2G + 2], par, 53);                   // G0 is in its buffer (and fenced for the async proxy)
      if (ptid == 0) store_tile(&tm_g0, buf_h2(t), t);
      colsum(buf_h2(t), cs[2]);
      if (ptid == 0) tma_store_wait_read<0>();               // the G0 store has read the buffer before anything reuses it
      __syncwarp(); if (lane == 0) mbar_arrive(&bars[kB2Cs + 2]);
      if (ptid == 0) {
        mbar_wait(&bars[kB2Sfree], par, 54);                 // the last chain GEMM of tile t has read G0 ...
        mbar_wait(&bars[kB2Cs + 2], par, 56);                // ... and so have BOTH column-sum warps: the buffer takes e of tile t + 2
        if (t + 2 < T) load_e(t + 2);
        mbar_wait(&bars[kB2De], par, 55);                    // d e of tile t is staged in its first buffer
        store_tile(&tm_de, buf_h1(t), t);
        tma_store_wait_read<0>();
        mbar_arrive(&bars[kB2DeFree]);                       // ... which dY of tile t + 1 may now overwrite
      }
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
      const uint32_t dW2 = tmem_base + 256, dW1 = tmem_base + 384;
      uint32_t epi_seen = 0;                                 // epilogue phases observed so far
      auto wait_epi_count = [&](uint32_t n) {                // returns once n epilogue phases have completed (phase i: barrier i & 1)
        while (epi_seen < n) {
          mbar_spin(&bars[kB2Epi + int(epi_seen & 1)], (epi_seen >> 1) & 1, 60);
          ++epi_seen;
        }
        fence_after_sync();
      };
      auto chain = [&](int64_t t, uint32_t a_addr, uint32_t b_addr, bool b_mn) {      // acc(t) = A[128 x 128] (K-major) * B
        const uint32_t acc = tmem_base + uint32_t(t & 1) * 128;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = sdesc_kmajor(a_addr + (ks >> 2) * kPanel + (ks & 3) * 32);
          const uint64_t bd = b_mn ? sdesc_mnmajor(b_addr + ks * 2048, kPanel) : sdesc_kmajor(b_addr + (ks >> 2) * kPanel + (ks & 3) * 32);
          mma_ss(acc, ad, bd, b_mn ? id_kmn : id_kk, ks != 0);
        }
      };
      auto wgrad = [&](uint32_t d, uint32_t g_addr, uint32_t z_addr, bool first) {   // d (+)= G^T Z over the tile's 128 rows
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_ss(d, sdesc_mnmajor(g_addr + ks * 2048, kPanel), sdesc_mnmajor(z_addr + ks * 2048, kPanel), id_mm, !(first && ks == 0));
      };
      auto commit_acc = [&](int64_t t) { mma_commit(&bars[kB2Acc + int(t & 1)]); };
      auto step0 = [&](int64_t t) {                          // e We^T; the tile's edge rows have landed
        mbar_wait(&bars[kB2Full + int(t & 1)], uint32_t(t >> 1) & 1, 61);
        fence_after_sync();
        chain(t, buf_h1(t), sbase + kB2We, false);
        commit_acc(t);
      };
      if (T > 0) step0(0);
      // Epilogue phases of round k, in order (absent tiles have no phases): k == 0: E0 E1 E2 (tile A = 0);  0 < k < T: E3(B) E0(A)
      // E4(B) E1(A) E5(B) E2(A);  k == T: E3 E4 E5 (tile B = T - 1).  `base` = phases completed before round k.
      for (int64_t k = 0; k <= T; ++k) {
        const bool hasA = k < T, hasB = k >= 1;
        const int64_t A = k, B = k - 1;
        const uint32_t base = k == 0 ? 0u : uint32_t(3 + 6 * (k - 1));
        if (hasB) {                                          // 3: dY W2 ; dW2     (after E2(B), the last phase of the previous round)
          wait_epi_count(base); stamp(k, 0);
          chain(B, buf_dy(B), sbase + kB2W2, true); wgrad(dW2, buf_dy(B), buf_h2(B), B == 0); commit_acc(B);
          wait_epi_count(base + 1); stamp(k, 1);             // 4: dH2' W1 ; dW1   (after E3(B))
          chain(B, buf_h2(B), sbase + kB2W1, true); wgrad(dW1, buf_h2(B), buf_h1(B), B == 0); commit_acc(B);
        }
        if (hasA) {                                          // 1: H1 W1^T          (after E0(A))
          wait_epi_count(base + (hasB ? 2 : 1)); stamp(k, 2);
          chain(A, buf_h1(A), sbase + kB2W1, false); commit_acc(A);
        }
        if (hasB) {                                          // 5: G0 We            (after E4(B))
          wait_epi_count(base + (hasA ? 3 : 2)); stamp(k, 3);
          chain(B, buf_h2(B), sbase + kB2We, true); commit_acc(B); mma_commit(&bars[kB2Sfree]);
        }
        if (hasA) {                                          // 2: H2 W2^T          (after E1(A))
          wait_epi_count(base + (hasB ? 4 : 2)); stamp(k, 4);
          chain(A, buf_h2(A), sbase + kB2W2, false); commit_acc(A);
        }
        if (k + 1 < T) {                                     // 0 of the next tile  (after E5(B): its accumulator slot is drained)
          if (hasB) wait_epi_count(base + 5);
          stamp(k, 5);
          step0(k + 1);
          stamp(k, 6);
        }
      }
      if (T > 0) mma_commit(&bars[kB2Final]);                // every MMA of this CTA, for the accumulator drain
    }
  } else {
    // =============================== epilogue ============================================================
    const int q = warp & 3, hh = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const float2 *b0 = reinterpret_cast<const float2*>(prm + hh * 64), *b1 = reinterpret_cast<const float2*>(prm + kD + hh * 64),
                 *b2 = reinterpret_cast<const float2*>(prm + 2 * kD + hh * 64), *gam = reinterpret_cast<const float2*>(prm + 3 * kD + hh * 64);
    uint32_t acc_phase[2] = {0, 0};
    uint32_t epi_index = 0;                                    // index of the epilogue phase in progress (barrier epi_index & 1)
    int32_t si_next = 0, ri_next = 0;                          // gather indices of the next tile A, read at the end of phase 2
    float cbeta[2] = {0.f, 0.f}, cgamma[2] = {0.f, 0.f};
    auto acc_of = [&](int64_t t) -> uint32_t { return tmem_base + lane_addr + uint32_t(t & 1) * 128 + hh * 64; };
    auto wait_acc = [&](int64_t t, int tag) {
      const int s = int(t & 1);
      mbar_spin(&bars[kB2Acc + s], acc_phase[s]++ & 1, tag);
      fence_after_sync();
    };
    auto wait_cs = [&](int j, int64_t t) { mbar_wait(&bars[kB2Cs + j], uint32_t(t) & 1, 70 + j); };
    auto arrive = [&](int which) { __syncwarp(); if (lane == 0) mbar_arrive(&bars[which]); };
    auto done = [&](int producers_j) {   // producers_j >= 0: the tile just written is also the producers' column-sum input j
      fence_async_smem();          // generic-proxy tile writes -> visible to the tensor core's / TMA's async-proxy reads
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars[kB2Epi + int(epi_index & 1)]);
        if (producers_j >= 0) mbar_arrive(&bars[kB2G + producers_j]);
      }
      ++epi_index;
    };
    const uint32_t row_off = hh * kPanel;                      // my 64 columns = panel hh of every buffer
    auto store_row = [&](uint32_t bufaddr, const uint32_t* w) {
#pragma unroll
      for (int k = 0; k < 8; ++k) st_shared128(bufaddr + row_off + sw128_chunk(r, k), w + 4 * k);
    };
    auto load_row = [&](uint32_t bufaddr, uint32_t* w) {
#pragma unroll
      for (int k = 0; k < 8; ++k) ld_shared128(bufaddr + row_off + sw128_chunk(r, k), w + 4 * k);
    };
    auto tile_row = [&](int64_t t) -> int64_t { return int64_t(tile_y(t)) + r; };
    auto row_index = [&](const int32_t* idx, int64_t grow) -> int32_t { return idx != nullptr ? __ldg(idx + grow) : int32_t(grow); };
    // dO = grad_out[row] + grad_agg[receiver]: one rounding to bf16, the value phases 2 and 5 both see
    auto load_dq = [&](int64_t t, int32_t ri, uint32_t* dq) {
      const int64_t grow = tile_row(t);
#pragma unroll
      for (int j = 0; j < 64; ++j) dq[j] = 0u;
      if (grow < rows) {
        if (a.grad_out != nullptr) {
          const __nv_bfloat16* dorow = a.grad_out + grow * kD + hh * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) ldg256(dorow + 16 * k, dq + 8 * k);
        }
        if (a.grad_agg != nullptr) {
          const __nv_bfloat16* garow = a.grad_agg + int64_t(ri) * kD + hh * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) ldg256_l1(garow + 16 * k, dq + 32 + 8 * k);
        }
      }
    };

    // ---- E0: H1 = relu(e We^T + Ps[s] + Pr[r] + b0) -> over e ---------------------------------------------------
    auto load_tables = [&](int32_t si, int32_t ri, uint32_t* pq) {          // Ps[s] and Pr[r]: my 64 columns of each
#pragma unroll
      for (int j = 0; j < 64; ++j) pq[j] = 0u;
      const __nv_bfloat16* psrow = a.proj_s + int64_t(si) * kD + hh * 64;
#pragma unroll
      for (int k = 0; k < 4; ++k) ldg256_l1(psrow + 16 * k, pq + 8 * k);
      if (a.proj_r != nullptr) {
        const __nv_bfloat16* prrow = a.proj_r + int64_t(ri) * kD + hh * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) ldg256_l1(prrow + 16 * k, pq + 32 + 8 * k);
      }
    };
    auto E0 = [&](int64_t t, int32_t si, int32_t ri) {
      uint32_t pq[64];
      load_tables(si, ri, pq);
      wait_acc(t, 100);                                         // step 0 has consumed e: H1 takes its buffer
      const uint32_t acc = acc_of(t);
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
      store_row(buf_h1(t), h);
      done(-1);
    };
    // ---- E1: H2 = relu(H1 W1^T + b1) -> the buffer dY of the previous tile has left -------------------------------
    auto E1 = [&](int64_t t) {
      if (t > 0) wait_cs(0, t - 1);                             // the producers' dY column sum of tile t - 1 (same buffer) is done
      wait_acc(t, 101);
      const uint32_t acc = acc_of(t);
      uint32_t h[32];
#pragma unroll
      for (int cg = 0; cg < 2; ++cg) {
        uint32_t v[32];
        tmem_ld32(acc + cg * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 x = __fadd2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), b1[cg * 16 + j]);
          h[cg * 16 + j] = cvt_relu_bf16x2(x.x, x.y);
        }
      }
      store_row(buf_h2(t), h);
      done(-1);
    };
    // ---- E2: y = H2 W2^T + b2 ; LayerNorm forward statistics and backward -> dY ; LN vector column sums -----------
    auto E2 = [&](int64_t t, int32_t ri) {
      uint32_t dq[64];
      load_dq(t, ri, dq);                                       // lands while the statistics pass runs
      wait_acc(t, 102);
      const uint32_t acc = acc_of(t);
      const uint32_t C = buf_dy(t);
      float2* xch = reinterpret_cast<float2*>(smem + (C - sbase));   // row-half exchange area: head of the dY buffer, idle until dY is stored
      uint32_t dreg[32], preg[32];
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
      if (t > 0) mbar_wait(&bars[kB2DeFree], uint32_t(t - 1) & 1, 73);   // d e of tile t - 1 has left this buffer (exchange area + dY)
      xch[hh * kTile + r] = make_float2(mean_h, m2h);
      epi_bar_sync();
      const float2 oth = xch[(1 - hh) * kTile + r];
      const float mean = 0.5f * (mean_h + oth.x);
      const float dm = mean_h - oth.x;
      const float rstd = rsqrtf(fmaxf(m2h + oth.y + 32.0f * dm * dm, 0.f) * (1.0f / kD) + kEps);
      const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mean * rstd, -mean * rstd);
      // dO = grad_out[row] + grad_agg[receiver]: one rounding to bf16, the value phase 5 reproduces
#pragma unroll
      for (int j = 0; j < 32; ++j) dreg[j] = add_bf16x2(dq[j], dq[32 + j]);
      float2 m1a = make_float2(0.f, 0.f), m2a = m1a;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        y[j] = __ffma2_rn(y[j], rs2, nm2);                      // yhat
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
      {
        uint32_t o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float2 z = __fmul2_rn(unpack_bf16x2(dreg[j]), gam[j]);
          const float2 u = __fmul2_rn(__ffma2_rn(y[j], nmm2, __fadd2_rn(z, nm1)), rs2);     // rstd (dO gamma - m1 - yhat m2)
          o[j] = pack_bf16(u.x, u.y);
        }
        epi_bar_sync();                                         // every thread has read both exchanges: dY may overwrite them
        store_row(C, o);
      }
      done(0);
      if (tid == 0) stamp(t, 15);
      if (t + 1 < T) {                                            // gather indices of the next tile: they land under the column sums below
        const int64_t gn = tile_row(t + 1);
        si_next = 0; ri_next = 0;
        if (gn < rows) { si_next = row_index(a.senders, gn); ri_next = row_index(a.receivers, gn); }
      }
      // LayerNorm vector gradients: column sums of dO * yhat (gamma) and dO (beta) over this warp's 32 rows
#pragma unroll
      for (int cg = 0; cg < 2; ++cg) {
        float p[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) { p[2 * j] = bf16_lo(preg[cg * 16 + j]); p[2 * j + 1] = bf16_hi(preg[cg * 16 + j]); }
        cgamma[cg] += warp_colsum32(p, lane);
#pragma unroll
        for (int j = 0; j < 16; ++j) { p[2 * j] = bf16_lo(dreg[cg * 16 + j]); p[2 * j + 1] = bf16_hi(dreg[cg * 16 + j]); }
        cbeta[cg] += warp_colsum32(p, lane);
      }
    };
    // ---- E3: dH2' = (dY W2) * [H2 > 0] -> over H2 ---------------------------------------------------------------------
    auto relu_bwd_phase = [&](int64_t t, uint32_t mask_buf, uint32_t dst_buf, int tag, int producers_j) {
      wait_acc(t, tag);                                         // the commit follows the step's weight-gradient MMAs: they have read dst_buf
      const uint32_t acc = acc_of(t);
#pragma unroll
      for (int cg = 0; cg < 2; ++cg) {
        uint32_t v[32], hw[16], o[16];
        tmem_ld32(acc + cg * 32, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) ld_shared128(mask_buf + row_off + sw128_chunk(r, 4 * cg + k), hw + 4 * k);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          o[j] = relu_bwd_bf16x2(pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), hw[j]);
#pragma unroll
        for (int k = 0; k < 4; ++k) st_shared128(dst_buf + row_off + sw128_chunk(r, 4 * cg + k), o + 4 * k);
      }
      done(producers_j);
    };
    // ---- E5: d e = G0 We + dO -> staged over H1, stored by the producers' TMA -----------------------------------------
    auto E5 = [&](int64_t t, int32_t ri) {
      uint32_t dreg[32];
      uint32_t dq[64];
      load_dq(t, ri, dq);
      wait_acc(t, 105);
#pragma unroll
      for (int j = 0; j < 32; ++j) dreg[j] = add_bf16x2(dq[j], dq[32 + j]);
      const uint32_t acc = acc_of(t);
      uint32_t v0[32], v1[32];
      tmem_ld32(acc, v0);
      tmem_ld32(acc + 32, v1);
      tmem_ld_wait();
      fence_before_sync();
      arrive(kB2Epi + int(epi_index & 1));                      // accumulator in registers: the slot's next tile may start
      ++epi_index;
      uint32_t o[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 x0 = __fadd2_rn(make_float2(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1])), unpack_bf16x2(dreg[j]));
        const float2 x1 = __fadd2_rn(make_float2(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1])), unpack_bf16x2(dreg[16 + j]));
        o[j] = pack_bf16(x0.x, x0.y);
        o[16 + j] = pack_bf16(x1.x, x1.y);
      }
      store_row(buf_h1(t), o);                                  // H1 is dead: dW1's MMAs and this thread's mask read are behind us
      fence_async_smem();
      arrive(kB2De);
    };

    // gather indices of tile A (read during the previous round) and, kept from its phase 2, of tile B for its phase 5
    int32_t si_A = 0, ri_A = 0, ri_B = 0;
    if (T > 0) {
      const int64_t g0 = tile_row(0);
      if (g0 < rows) { si_A = row_index(a.senders, g0); ri_A = row_index(a.receivers, g0); }
    }
    for (int64_t k = 0; k <= T; ++k) {
      const bool hasA = k < T, hasB = k >= 1;
      const int64_t A = k, B = k - 1;
      if (tid == 0) stamp(k, 8);
      if (hasB) relu_bwd_phase(B, buf_h2(B), buf_h2(B), 103, 1);                                      // E3(B): dH2' over H2
      if (tid == 0) stamp(k, 9);
      if (hasA) E0(A, si_A, ri_A);                                                                    // E0(A)
      if (tid == 0) stamp(k, 10);
      if (hasB) { wait_cs(1, B); relu_bwd_phase(B, buf_h1(B), buf_h2(B), 104, 2); }                   // E4(B): G0 over dH2' (its column sum is done)
      if (tid == 0) stamp(k, 11);
      if (hasA) E1(A);                                                                                // E1(A)
      if (tid == 0) stamp(k, 12);
      if (hasB) E5(B, ri_B);                                                                          // E5(B): d e staged over H1
      if (tid == 0) stamp(k, 13);
      if (hasA) E2(A, ri_A);                                                                          // E2(A) (+ next tile's indices, LN vector sums)
      if (tid == 0) stamp(k, 14);
      ri_B = ri_A;
      si_A = si_next; ri_A = ri_next;
    }
    // ---- drain the weight-gradient accumulators and the LayerNorm vector partials ---------------------------------------
    if (T > 0) {
      mbar_wait(&bars[kB2Final], 0, 75);
      fence_after_sync();
    }
#pragma unroll 1
    for (int z = 1; z < 3; ++z) {
      const uint32_t col0 = z == 2 ? 256u : 384u;                // z = 1: dW1, 2: dW2
      float* dst = a.w_partial + ((int64_t(blockIdx.x) * 3 + z) * kD + r) * kD + hh * 64;
#pragma unroll 1
      for (int cg = 0; cg < 2; ++cg) {
        uint32_t v[32];
        if (T > 0) {
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

// ---- host side -----------------------------------------------------------------------------------------------------------
int make_rows_tensor_map(CUtensorMap* tm, const void* base, int64_t rows);   // edge_fwd_tc.cu

int edge_bwd2_launch(int64_t rows, int64_t tiles, int grid, const void* packed, const void* dense, const void* proj_s, const void* proj_r,
                     const int32_t* senders, const int32_t* receivers, int w0_chunks, int w0_chunk0, const void* grad_out, const void* grad_agg,
                     void* grad_dense, void* grad_pre0, float* w_partial, float* epi_colpart, float* prod_colpart, const char* name,
                     cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    uint32_t* dbg = debug_buffer_device();
    HGN_CUDA_OK(cudaMemcpyToSymbol(tc05::g_debug_words, &dbg, sizeof(dbg)));
    HGN_CUDA_OK(cudaFuncSetAttribute(edge_bwd2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kB2Smem)));
    configured = true;
  }
  Bwd2Args a{};
  a.proj_s = static_cast<const __nv_bfloat16*>(proj_s);
  a.proj_r = static_cast<const __nv_bfloat16*>(proj_r);
  a.senders = senders;
  a.receivers = receivers;
  a.grad_out = static_cast<const __nv_bfloat16*>(grad_out);
  a.grad_agg = static_cast<const __nv_bfloat16*>(grad_agg);
  a.w_partial = w_partial;
  a.epi_colpart = epi_colpart;
  a.prod_colpart = prod_colpart;
  a.w0_chunks = w0_chunks;
  a.w0_chunk0 = w0_chunk0;
  const int64_t map_rows = rows > 0 ? rows : 1;      // rows == 0: maps over one (never accessed) row keep the encoder happy
  CUtensorMap tm_e, tm_g0, tm_de;
  if (int rc = make_rows_tensor_map(&tm_e, rows > 0 ? dense : w_partial, map_rows)) return rc;
  if (int rc = make_rows_tensor_map(&tm_g0, rows > 0 ? grad_pre0 : w_partial, map_rows)) return rc;
  if (int rc = make_rows_tensor_map(&tm_de, rows > 0 ? grad_dense : w_partial, map_rows)) return rc;
  static long long* tl_dev = nullptr;
  { const char* ab = getenv("HGN_TC_ABLATE"); if (ab != nullptr && (atoi(ab) & 64)) {
      if (tl_dev == nullptr) cudaMalloc(&tl_dev, 10 * 32 * sizeof(long long));
      cudaMemsetAsync(tl_dev, 0, 10 * 32 * sizeof(long long), st);
      a.timeline = tl_dev;
  } }
  {
  HGN_TIMED(name, st);
  edge_bwd2_tc_kernel<<<unsigned(grid), kB2Threads, kB2Smem, st>>>(rows, tiles, static_cast<const uint8_t*>(packed), a, tm_e, tm_g0, tm_de);
  }
  HGN_LAUNCH_OK(name);
  if (a.timeline != nullptr) {
    long long h[10 * 32];
    cudaMemcpyAsync(h, a.timeline, sizeof(h), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    for (int t = 0; t < 10; ++t) {
      fprintf(stderr, "bwd2 round %d:", t);
      for (int k = 0; k < 16; ++k) fprintf(stderr, " %lld", h[t * 32 + k] ? h[t * 32 + k] - h[8] : -1);
      fprintf(stderr, "\n");
    }
  }
  return HGN_OK;
}

}  // namespace hgn
