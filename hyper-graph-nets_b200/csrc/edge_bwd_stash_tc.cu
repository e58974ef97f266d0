// bf16 tcgen05 backward of the projected edge / node update from STASHED hidden activations (HGN_BF16 only).
//
// The forward kernel (edge_fwd_tc.cu) optionally writes H1 = relu(pre0) and H2 = relu(H1 W1^T + b1) as bf16 [rows,128].
// With them the backward needs neither the node tables nor the first two GEMMs of the recompute path (edge_tc.cu):
//     y   = H2 W2^T + b2                      step 2 (recomputed: the LayerNorm backward needs y, and y is one GEMM away)
//     dY  = LN'(y; dO)                         E2,   dO = grad_out[row] + grad_agg[receiver[row]]
//     dH2'= (dY W2)  * [H2 > 0]                step 3 / E3      dW2 += dY^T H2
//     G0  = (dH2' W1) * [H1 > 0]               step 4 / E4      dW1 += dH2'^T H1      (G0 = d loss / d pre0 -> HBM)
//     d e = G0 We + dO                         step 5 / E5      dWe += G0^T e
// Four MMA steps and epilogue phases per 128-row tile instead of six; the price is 2 x 256 B per row written by the forward
// and read here (the kernels are latency-bound, not HBM-bound: see DESIGN.md).
//
// Persistent, one CTA per SM, one tile at a time (shared memory holds the three weight blocks and four 32 KiB tile buffers):
//   warps 0-7   epilogue (row = TMEM lane, 64 columns per thread), as in edge_tc.cu
//   warps 8-9   bias-gradient column sums of dY / dH2' / G0 out of the operand buffers, L2 prefetch of the next tile's
//               gradient rows, and (thread 0 of warp 8) every TMA transfer: H1 / H2 / e tiles in, G0 and d e tiles out
//   warp 10     MMA issuer
// Buffers (fixed roles): S = e | A = H1, later the d e staging tile | B = H2, then dH2' | C = dY, then G0.
// The next tile's H2 is loaded as soon as the dW1 MMAs are done with B, e after the dWe MMAs, H1 after the d e store has read A.
// TMEM (512 columns): chain accumulator [0,128) | dW2 [128,256) | dW1 [256,384) | dWe [384,512).
#include <cuda.h>
#include <stdlib.h>

#include "tile_common.cuh"

namespace hgn {

constexpr int kSbEpiThreads = 256, kSbProdThreads = 64;
constexpr int kSbThreads = kSbEpiThreads + kSbProdThreads + 32;    // 11 warps
constexpr uint32_t kSbWe = 0, kSbW1 = kChunkBytes, kSbW2 = 2 * kChunkBytes, kSbBuf = 3 * kChunkBytes;
constexpr uint32_t kSbParams = kSbBuf + 4 * kChunkBytes;           // b0 b1 b2 gamma beta (fp32 x 128 each)
constexpr uint32_t kSbBars = kSbParams + 5 * kD * 4;
constexpr uint32_t kSbSmem = kSbBars + 128;
enum { kSbFullS = 0, kSbFullA = 1, kSbFullB = 2, kSbAcc = 3, kSbEpi = 4, kSbG = 5, kSbCs = 8, kSbTmem = 11, kSbBfree = 12, kSbSfree = 13,
       kSbDe = 14, kSbFinal = 15 };

struct StashBwdArgs {
  const int32_t* receivers;         // gather index of grad_agg (null = identity)
  const __nv_bfloat16* grad_out;    // [rows,128] dense part of d loss / d out (may be null)
  const __nv_bfloat16* grad_agg;    // [N,128] gathered through receivers (may be null)
  float* w_partial;                 // [grid][3][128][128]  z = 0: dWe, 1: dW1, 2: dW2
  float* epi_colpart;               // [grid][4][2][128]    beta, gamma partial column sums per lane quadrant
  float* prod_colpart;              // [grid][3][128]       db2, db1, db0
  int w0_chunks, w0_chunk0;         // W0 is [128][128 w0_chunks]; the dense input multiplies chunk w0_chunk0
};

__global__ void __launch_bounds__(kSbThreads, 1)
edge_bwd_stash_tc_kernel(int64_t rows, int64_t num_tiles, const uint8_t* __restrict__ packed, StashBwdArgs a,
                         const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_h1,
                         const __grid_constant__ CUtensorMap tm_h2, const __grid_constant__ CUtensorMap tm_g0,
                         const __grid_constant__ CUtensorMap tm_de) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSbBars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PackedTc P(a.w0_chunks);
  float* prm = reinterpret_cast<float*>(smem + kSbParams);
  {
    const float* pg = reinterpret_cast<const float*>(packed + P.params);
    for (int i = tid; i < 5 * kD; i += kSbThreads) prm[i] = pg[i];
    const __nv_bfloat16* w0g = reinterpret_cast<const __nv_bfloat16*>(packed + P.w0);
    load_weight_block(sbase + kSbWe, w0g + a.w0_chunk0 * kD, int64_t(a.w0_chunks) * kD, tid, kSbThreads);
    load_weight_block(sbase + kSbW1, reinterpret_cast<const __nv_bfloat16*>(packed + P.w1), kD, tid, kSbThreads);
    load_weight_block(sbase + kSbW2, reinterpret_cast<const __nv_bfloat16*>(packed + P.w2), kD, tid, kSbThreads);
    cp_async_commit();
    if (tid == 0) {
      mbar_init(&bars[kSbFullS], 1); mbar_init(&bars[kSbFullA], 1); mbar_init(&bars[kSbFullB], 1);
      mbar_init(&bars[kSbAcc], 1);
      mbar_init(&bars[kSbEpi], kSbEpiThreads);
      for (int k = 0; k < 3; ++k) { mbar_init(&bars[kSbG + k], kSbEpiThreads); mbar_init(&bars[kSbCs + k], kSbProdThreads); }
      mbar_init(&bars[kSbBfree], 1); mbar_init(&bars[kSbSfree], 1);
      mbar_init(&bars[kSbDe], kSbEpiThreads);
      mbar_init(&bars[kSbFinal], 1);
      mbar_init_fence();
    }
    if (warp == 10) tmem_alloc<512>(reinterpret_cast<uint32_t*>(&bars[kSbTmem]));
    cp_async_wait<0>();
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars[kSbTmem]);
  const int64_t my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const uint32_t S = sbase + kSbBuf, A = S + kChunkBytes, B = A + kChunkBytes, C = B + kChunkBytes;
  auto tile_y = [&](int64_t t) -> int { return int((blockIdx.x + t * gridDim.x) * kTile); };

  if (warp == 8 || warp == 9) {
    // =============================== producers =========================================================
    const int ptid = tid - kSbEpiThreads, pw = warp - 8;
    auto tma_in = [&](const CUtensorMap* tm, uint32_t dst, int64_t t, uint64_t* bar) {
      mbar_expect_tx(bar, kChunkBytes);
      tma_load_2d(dst, tm, 0, tile_y(t), bar);
      tma_load_2d(dst + kPanel, tm, 64, tile_y(t), bar);
    };
    auto tma_out = [&](const CUtensorMap* tm, uint32_t src, int64_t t) {
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
    if (ptid == 0 && my_tiles > 0) {
      tma_in(&tm_h2, B, 0, &bars[kSbFullB]);
      tma_in(&tm_h1, A, 0, &bars[kSbFullA]);
      tma_in(&tm_e, S, 0, &bars[kSbFullS]);
    }
    for (int64_t t = 0; t < my_tiles; ++t) {
      const uint32_t par = uint32_t(t) & 1;
      const bool more = t + 1 < my_tiles;
      if (more) {
        // the gradient rows tile t+1 will read are pulled into L2 a tile ahead (never by the epilogue warps: their proxy
        // fences wait for outstanding prefetches)
        const int64_t row0 = (blockIdx.x + (t + 1) * gridDim.x) * kTile;
        auto pf = [](const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); };
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int64_t grow = row0 + ptid + 64 * j;
          if (grow < rows) {
            if (a.grad_out != nullptr) { pf(a.grad_out + grow * kD); pf(a.grad_out + grow * kD + 64); }
            if (a.grad_agg != nullptr) {
              const int64_t ri = a.receivers != nullptr ? int64_t(__ldg(a.receivers + grow)) : grow;
              pf(a.grad_agg + ri * kD); pf(a.grad_agg + ri * kD + 64);
            }
          }
        }
      }
      mbar_wait(&bars[kSbG + 0], par, 50);
      colsum(C, cs[0]);                                      // dY
      mbar_arrive(&bars[kSbCs + 0]);
      mbar_wait(&bars[kSbG + 1], par, 51);
      colsum(B, cs[1]);                                      // dH2'
      mbar_arrive(&bars[kSbCs + 1]);
      if (ptid == 0 && more) {                               // the dW1 MMAs are done with B: the next tile's H2 may land
        mbar_wait(&bars[kSbBfree], par, 52);
        tma_in(&tm_h2, B, t + 1, &bars[kSbFullB]);
      }
      mbar_wait(&bars[kSbG + 2], par, 53);                   // G0 is in buffer C (and fenced for the async proxy)
      if (ptid == 0) tma_out(&tm_g0, C, t);
      colsum(C, cs[2]);
      if (ptid == 0) tma_store_wait_read<0>();               // the store has read C before E2 of the next tile reuses it
      mbar_arrive(&bars[kSbCs + 2]);
      if (ptid == 0) {
        if (more) {                                          // the dWe MMAs are done with S
          mbar_wait(&bars[kSbSfree], par, 54);
          tma_in(&tm_e, S, t + 1, &bars[kSbFullS]);
        }
        mbar_wait(&bars[kSbDe], par, 55);                    // d e is staged in A
        tma_out(&tm_de, A, t);
        if (more) {
          tma_store_wait_read<0>();
          tma_in(&tm_h1, A, t + 1, &bars[kSbFullA]);
        }
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
      const uint32_t acc = tmem_base, dW2 = tmem_base + 128, dW1 = tmem_base + 256, dWe = tmem_base + 384;
      uint32_t epi_phase = 0;
      auto wait_epi = [&]() {
        mbar_spin(&bars[kSbEpi], epi_phase++ & 1, 60);
        fence_after_sync();
      };
      auto wait_full = [&](int which, int64_t t) {
        mbar_wait(&bars[which], uint32_t(t) & 1, 61);
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
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          mma_ss(d, sdesc_mnmajor(g_addr + ks * 2048, kPanel), sdesc_mnmajor(z_addr + ks * 2048, kPanel), id_mm, !(first && ks == 0));
      };
      for (int64_t t = 0; t < my_tiles; ++t) {
        const bool first = t == 0;
        wait_full(kSbFullB, t);
        if (!first) wait_epi();                                  // previous tile's last phase has drained the accumulator
        chain(B, sbase + kSbW2, false); mma_commit(&bars[kSbAcc]);                                            // 2: H2 W2^T
        wait_epi(); chain(C, sbase + kSbW2, true); wgrad(dW2, C, B, first); mma_commit(&bars[kSbAcc]);           // 3: dY W2 ; dW2
        wait_epi(); wait_full(kSbFullA, t);
        chain(B, sbase + kSbW1, true); wgrad(dW1, B, A, first); mma_commit(&bars[kSbAcc]); mma_commit(&bars[kSbBfree]);   // 4: dH2' W1 ; dW1
        wait_epi(); wait_full(kSbFullS, t);
        chain(C, sbase + kSbWe, true); mma_commit(&bars[kSbAcc]); wgrad(dWe, C, S, first); mma_commit(&bars[kSbSfree]);   // 5: G0 We ; dWe
      }
      if (my_tiles > 0) mma_commit(&bars[kSbFinal]);           // every MMA of this CTA, for the accumulator drain
    }
  } else {
    // =============================== epilogue ============================================================
    const int q = warp & 3, hh = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t acc = tmem_base + lane_addr + hh * 64;
    const float2 *b2 = reinterpret_cast<const float2*>(prm + 2 * kD + hh * 64), *gam = reinterpret_cast<const float2*>(prm + 3 * kD + hh * 64);
    uint32_t acc_phase = 0;
    float cbeta[2] = {0.f, 0.f}, cgamma[2] = {0.f, 0.f};
    auto wait_acc = [&](int tag) {
      mbar_wait(&bars[kSbAcc], acc_phase++ & 1, tag);
      fence_after_sync();
    };
    auto wait_cs = [&](int k, int64_t tt) { mbar_wait(&bars[kSbCs + k], uint32_t(tt) & 1, 70 + k); };
    auto done = [&](int producers_k) {
      fence_async_smem();          // generic-proxy tile writes -> visible to the tensor core's / TMA's async-proxy reads
      fence_before_sync();
      mbar_arrive(&bars[kSbEpi]);
      mbar_arrive(&bars[kSbG + producers_k]);
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
    auto tile_row = [&](int64_t tt) { return (blockIdx.x + tt * gridDim.x) * kTile + r; };
    float2* xch = reinterpret_cast<float2*>(smem + kSbBuf + 3 * kChunkBytes);   // LayerNorm row-half exchange: head of buffer C
    int32_t ri = 0;
    if (my_tiles > 0 && a.grad_agg != nullptr) {
      const int64_t g0 = tile_row(0);
      if (g0 < rows) ri = a.receivers != nullptr ? __ldg(a.receivers + g0) : int32_t(g0);
    }
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t grow = tile_row(t);
      const bool valid = grow < rows;
      const int64_t gnext = tile_row(t + 1);
      const bool vnext = t + 1 < my_tiles && gnext < rows;
      // dO = grad_out[row] + grad_agg[receiver]: requested now, consumed after the accumulator wait of the first phase
      uint32_t dreg[32];
      {
        uint32_t dq[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) dq[j] = 0u;
        if (valid && a.grad_out != nullptr) {
          const __nv_bfloat16* dorow = a.grad_out + grow * kD + hh * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) ldg256(dorow + 16 * k, dq + 8 * k);
        }
        if (valid && a.grad_agg != nullptr) {
          const __nv_bfloat16* garow = a.grad_agg + int64_t(ri) * kD + hh * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) ldg256_l1(garow + 16 * k, dq + 32 + 8 * k);
        }
        if (t > 0) wait_cs(2, t - 1);                           // G0 of the previous tile has left buffer C (column sum + TMA store)
        wait_acc(102);
        // one rounding to bf16, the value every later use sees
#pragma unroll
        for (int j = 0; j < 32; ++j) dreg[j] = add_bf16x2(dq[j], dq[32 + j]);
      }
      // ---- E2: y = H2 W2^T + b2 ; LayerNorm forward statistics and backward -> dY -> C --------------------------------
      uint32_t preg[32];                                        // dO * yhat (bf16): gamma-gradient terms, summed after the phase
      {
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
        done(0);
      }
      // (in the shadow of MMA step 3) gamma gradient: column sums of dO * yhat over this warp's 32 rows
#pragma unroll
      for (int cg = 0; cg < 2; ++cg) {
        float p[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) { p[2 * j] = bf16_lo(preg[cg * 16 + j]); p[2 * j + 1] = bf16_hi(preg[cg * 16 + j]); }
        cgamma[cg] += warp_colsum32(p, lane);
      }
      // ---- E3: dH2' = (dY W2) * [H2 > 0] -> B ---------------------------------------------------------------------
      {
        wait_acc(103);
        mbar_wait(&bars[kSbFullB], uint32_t(t) & 1, 79);        // observe H2's TMA bytes (already seen by the MMA thread)
        uint32_t hw[32], o[32];
        load_row(B, hw);
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          uint32_t v[32];
          tmem_ld32(acc + cg * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            o[cg * 16 + j] = relu_bwd_bf16x2(pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), hw[cg * 16 + j]);
        }
        store_row(B, o);
        done(1);
      }
      // (in the shadow of MMA step 4) beta gradient: column sums of dO
#pragma unroll
      for (int cg = 0; cg < 2; ++cg) {
        float p[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) { p[2 * j] = bf16_lo(dreg[cg * 16 + j]); p[2 * j + 1] = bf16_hi(dreg[cg * 16 + j]); }
        cbeta[cg] += warp_colsum32(p, lane);
      }
      // ---- E4: G0 = (dH2' W1) * [H1 > 0] -> C (the producers' TMA sends it to HBM) ----------------------------------------
      {
        wait_cs(0, t);                                          // the producers' dY column sum has left buffer C
        wait_acc(104);
        mbar_wait(&bars[kSbFullA], uint32_t(t) & 1, 78);        // observe H1's TMA bytes (already seen by the MMA thread)
        uint32_t hw[32], o[32];
        load_row(A, hw);
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          uint32_t v[32];
          tmem_ld32(acc + cg * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            o[cg * 16 + j] = relu_bwd_bf16x2(pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), hw[cg * 16 + j]);
        }
        store_row(C, o);
        done(2);
        // index of the next tile's row (first used at the next tile's start)
        ri = 0;
        if (vnext && a.grad_agg != nullptr) ri = a.receivers != nullptr ? __ldg(a.receivers + gnext) : int32_t(gnext);
      }
      // ---- E5: d e = G0 We + dO -> staged in A, stored by the producers' TMA --------------------------------------------
      {
        wait_acc(105);
        uint32_t v0[32], v1[32];
        tmem_ld32(acc, v0);
        tmem_ld32(acc + 32, v1);
        tmem_ld_wait();
        fence_before_sync();
        mbar_arrive(&bars[kSbEpi]);                             // accumulator in registers: the next tile's step 2 may start
        uint32_t o[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 x0 = __fadd2_rn(make_float2(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1])), unpack_bf16x2(dreg[j]));
          const float2 x1 = __fadd2_rn(make_float2(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1])), unpack_bf16x2(dreg[16 + j]));
          o[j] = pack_bf16(x0.x, x0.y);
          o[16 + j] = pack_bf16(x1.x, x1.y);
        }
        // A (H1) is dead: this step's chain commit came after the dW1 MMAs, and this thread read its mask row in E4
        store_row(A, o);
        fence_async_smem();
        mbar_arrive(&bars[kSbDe]);
      }
    }
    // ---- drain the weight-gradient accumulators and the LayerNorm vector partials ---------------------------------------
    if (my_tiles > 0) {
      mbar_wait(&bars[kSbFinal], 0, 75);
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

// ---- host side -----------------------------------------------------------------------------------------------------------
int make_rows_tensor_map(CUtensorMap* tm, const void* base, int64_t rows);   // edge_fwd_tc.cu
void launch_edge_bwd_reduce(const float* w_partial, const float* epi_colpart, const float* prod_colpart, int parts, int w0_chunks, int w0_chunk0,
                            float* gW0, float* gW1, float* gW2, float* gb0, float* gb1, float* gb2, float* ggamma, float* gbeta,
                            cudaStream_t st);                                 // edge_tc.cu

// workspace layout identical to the recompute path's (edge_tc.cu: edge_bwd_layout)
int stash_backward_launch(int64_t rows, const void* dense, const void* h1, const void* h2, const int32_t* receivers, const void* packed,
                          int w0_chunks, int w0_chunk0, const void* grad_out, const void* grad_agg, void* grad_dense, void* grad_pre0,
                          float* gW0, float* gb0, float* gW1, float* gb1, float* gW2, float* gb2, float* ggamma, float* gbeta,
                          float* w_partial, float* epi_colpart, float* prod_colpart, int grid, const char* name, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    uint32_t* dbg = debug_buffer_device();
    HGN_CUDA_OK(cudaMemcpyToSymbol(tc05::g_debug_words, &dbg, sizeof(dbg)));
    HGN_CUDA_OK(cudaFuncSetAttribute(edge_bwd_stash_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSbSmem)));
    configured = true;
  }
  StashBwdArgs a{};
  a.receivers = receivers;
  a.grad_out = static_cast<const __nv_bfloat16*>(grad_out);
  a.grad_agg = static_cast<const __nv_bfloat16*>(grad_agg);
  a.w_partial = w_partial;
  a.epi_colpart = epi_colpart;
  a.prod_colpart = prod_colpart;
  a.w0_chunks = w0_chunks;
  a.w0_chunk0 = w0_chunk0;
  const int64_t tiles = ceil_div(rows, kTile);
  const int64_t map_rows = rows > 0 ? rows : 1;      // rows == 0: maps over one (never accessed) row keep the encoder happy
  CUtensorMap tm_e, tm_h1, tm_h2, tm_g0, tm_de;
  if (int rc = make_rows_tensor_map(&tm_e, rows > 0 ? dense : w_partial, map_rows)) return rc;
  if (int rc = make_rows_tensor_map(&tm_h1, rows > 0 ? h1 : w_partial, map_rows)) return rc;
  if (int rc = make_rows_tensor_map(&tm_h2, rows > 0 ? h2 : w_partial, map_rows)) return rc;
  if (int rc = make_rows_tensor_map(&tm_g0, rows > 0 ? grad_pre0 : w_partial, map_rows)) return rc;
  if (int rc = make_rows_tensor_map(&tm_de, rows > 0 ? grad_dense : w_partial, map_rows)) return rc;
  {
    HGN_TIMED(name, st);
    edge_bwd_stash_tc_kernel<<<unsigned(grid), kSbThreads, kSbSmem, st>>>(rows, tiles, static_cast<const uint8_t*>(packed), a, tm_e, tm_h1, tm_h2,
                                                                        tm_g0, tm_de);
  }
  HGN_LAUNCH_OK(name);
  {
    HGN_TIMED("reduce_weight_partials", st);
    launch_edge_bwd_reduce(w_partial, epi_colpart, prod_colpart, grid, w0_chunks, w0_chunk0, gW0, gW1, gW2, gb0, gb1, gb2, ggamma, gbeta, st);
  }
  HGN_LAUNCH_OK("edge_bwd_reduce");
  return HGN_OK;
}

}  // namespace hgn
