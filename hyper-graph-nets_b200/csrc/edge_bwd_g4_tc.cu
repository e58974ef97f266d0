// EXPERIMENTAL (opt-in: HGN_EDGE_BWD_TMA_GATHER=1) -- written after the round's GPU budget was spent: it compiles for sm_100a and has
// NOT run yet.  The default backward kernel is edge_bwd_tc_kernel (edge_tc.cu); nothing selects this one unless the switch is set.
//
// Edge backward with the table-row gathers moved from the load pipe to TMA.  DESIGN.md s3.2: in edge_bwd_tc_kernel every epilogue
// thread fetches its edge's rows of Ps[s], Pr[r] and grad_agg[r] with 32-byte ld.global's; a warp instruction then touches 32
// different 128-byte lines, the SM retires about one such sector per cycle, and the issuing warps stall for ~3.6 k of the ~19.4 k
// cycles a tile takes.  Here the two otherwise idle producer warps issue `cp.async.bulk.tensor.2d ... tile::gather4`
// (UTMALDG.2D.GATHER4: four 128-byte rows per instruction, every lane owns four rows of the tile) into tile buffers that are idle at
// that point of the schedule -- no extra shared memory:
//     Ps[s]        -> buffer A(t)   (free once the previous tile's d e store has read it; E0 reads its row and overwrites it with H1)
//     Pr[r]        -> buffer B(t)   (free once the dWe MMAs and the G0 column sums of the previous tile are done; E1 writes H2)
//     grad_agg[r]  -> buffer C(t)   (free once the dWe MMAs of the previous tile are done; read into registers before E1, E2 writes dY)
// The dense gradient row (grad_out) stays a register load.  Everything else -- the six chain GEMMs, the three weight-gradient GEMMs
// with TMEM-resident accumulators, the reductions, the outputs -- is edge_bwd_tc_kernel's; see edge_tc.cu for the description.
// New barriers: PsFull / PrFull / GaFull (TMA byte counts) and WeDone (a tcgen05.commit after the dWe MMAs of step 5).
#include <cuda.h>
#include <stdlib.h>

#include "tile_common.cuh"

namespace hgn {

__device__ __forceinline__ void tma_gather4(uint32_t smem_dst, const void* tmap, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               :: "r"(smem_dst), "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}

constexpr int kEgEpiThreads = 256, kEgProdThreads = 64;
constexpr int kEgThreads = kEgEpiThreads + kEgProdThreads + 32;    // 11 warps
constexpr uint32_t kEgWe = 0, kEgW1 = kChunkBytes, kEgW2 = 2 * kChunkBytes, kEgBuf = 3 * kChunkBytes;
constexpr uint32_t kEgParams = kEgBuf + 4 * kChunkBytes;           // b0 b1 b2 gamma beta (fp32 x 128 each)
constexpr uint32_t kEgBars = kEgParams + 5 * kD * 4;
constexpr uint32_t kEgSmem = kEgBars + 192;                        // 232 128 B of the 232 448 available (24 barrier slots)
// kEgG + k / kEgCs + k (k = 0 dY, 1 dH2', 2 dH1'): one barrier per tile-in-buffer hand-over, so each completes exactly one
// phase per tile and no waiter can fall two phases behind (a parity wait cannot tell phase n from phase n + 2)
enum { kEgFull = 0, kEgWg = 1, kEgAcc = 2, kEgEpi = 3, kEgG = 4, kEgCs = 7, kEgTmem = 10, kEgAfree = 11, kEgFinal = 12, kEgDe = 13, kEgDeFree = 14, kEgG0Free = 15,
       kEgPsFull = 16, kEgPrFull = 17, kEgGaFull = 18, kEgWeDone = 19 };

struct EdgeBwdG4Args {
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

__device__ __forceinline__ void stamp_g4(const EdgeBwdG4Args& a, int64_t t, int slot) {
  if (a.timeline != nullptr && blockIdx.x == 0 && t < 8) a.timeline[t * 48 + slot] = clock64();
}

__global__ void __launch_bounds__(kEgThreads, 1)
edge_bwd_g4_tc_kernel(int64_t rows, int64_t num_tiles, const uint8_t* __restrict__ packed, EdgeBwdG4Args a,
                      const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_g0,
                      const __grid_constant__ CUtensorMap tm_de, const __grid_constant__ CUtensorMap tm_ps,
                      const __grid_constant__ CUtensorMap tm_pr, const __grid_constant__ CUtensorMap tm_ga) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kEgBars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PackedTc P(a.w0_chunks);
  float* prm = reinterpret_cast<float*>(smem + kEgParams);
  {
    const float* pg = reinterpret_cast<const float*>(packed + P.params);
    for (int i = tid; i < 5 * kD; i += kEgThreads) prm[i] = pg[i];
    const __nv_bfloat16* w0g = reinterpret_cast<const __nv_bfloat16*>(packed + P.w0);
    load_weight_block(sbase + kEgWe, w0g + a.w0_chunk0 * kD, int64_t(a.w0_chunks) * kD, tid, kEgThreads);
    load_weight_block(sbase + kEgW1, reinterpret_cast<const __nv_bfloat16*>(packed + P.w1), kD, tid, kEgThreads);
    load_weight_block(sbase + kEgW2, reinterpret_cast<const __nv_bfloat16*>(packed + P.w2), kD, tid, kEgThreads);
    cp_async_commit();
    if (tid == 0) {
      mbar_init(&bars[kEgFull], 1);                           // one arrive.expect_tx per tile, completed by the TMA bytes
      mbar_init(&bars[kEgWg], 1);
      mbar_init(&bars[kEgFinal], 1);
      mbar_init(&bars[kEgDe], kEgEpiThreads / 32);            // epilogue barriers: one arrival per WARP (lane 0 after __syncwarp):
      mbar_init(&bars[kEgDeFree], 1);
      mbar_init(&bars[kEgG0Free], 1);
      mbar_init(&bars[kEgPsFull], 1); mbar_init(&bars[kEgPrFull], 1); mbar_init(&bars[kEgGaFull], 1); mbar_init(&bars[kEgWeDone], 1);
      mbar_init(&bars[kEgAcc], 1);
      mbar_init(&bars[kEgEpi], kEgEpiThreads / 32);           // 256 arrivals on one mbarrier serialise in the shared-memory unit
      mbar_init(&bars[kEgAfree], kEgEpiThreads / 32);
      for (int k = 0; k < 3; ++k) { mbar_init(&bars[kEgG + k], kEgEpiThreads / 32); mbar_init(&bars[kEgCs + k], kEgProdThreads / 32); }
      mbar_init_fence();
    }
    if (warp == 10) tmem_alloc<512>(reinterpret_cast<uint32_t*>(&bars[kEgTmem]));
    cp_async_wait<0>();
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars[kEgTmem]);
  const int64_t my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  // buffer roles of tile t: 0 = S (edge rows), 1 = A (H1), 2 = B (H2, then dH2'), 3 = C (dY, then dH1'); they rotate by one
  // buffer per tile so that S(t+1) = A(t), the first buffer that falls free (after step 4), receives the prefetch
  auto buf = [&](int role, int64_t t) -> uint32_t { return sbase + kEgBuf + uint32_t((role + t) & 3) * kChunkBytes; };

  if (warp == 8 || warp == 9) {
    // =============================== producers =========================================================
    const int ptid = tid - kEgEpiThreads, pw = warp - 8;
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
        mbar_expect_tx(&bars[kEgFull], kChunkBytes);
        tma_load_2d(dst, &tm_e, 0, y, &bars[kEgFull]);
        tma_load_2d(dst + kPanel, &tm_e, 64, y, &bars[kEgFull]);
      }
    };
    // ---- table rows by TMA: `tile::gather4` pulls four rows of 128 bytes per instruction (UTMALDG.2D.GATHER4) into a 128B-swizzled
    // panel, so the thread-per-row ld.global gathers (one 32-byte sector per lane and instruction, ~1 sector per cycle per SM,
    // issued by -- and stalling -- the epilogue warps) leave the load pipe.  Every lane of the issuing warp owns four rows of the
    // tile.  Destinations are buffers that are idle at that time: Ps[s] -> A(t) (until E0 overwrites it in place with H1),
    // Pr[r] -> B(t) (until E1 writes H2), grad_agg[r] -> C(t) (read into registers before E1; E2 then uses C for dY).
    auto tile_index = [&](const int32_t* idx, int64_t grow) -> int { return grow < rows ? (idx != nullptr ? __ldg(idx + grow) : int(grow)) : 0; };
    auto gather_tile = [&](const CUtensorMap* tm, const int32_t* idx, int64_t tt, uint32_t dst, uint64_t* bar) {
      const int64_t g0 = (blockIdx.x + tt * gridDim.x) * kTile + 4 * lane;
      const int r0 = tile_index(idx, g0), r1 = tile_index(idx, g0 + 1), r2 = tile_index(idx, g0 + 2), r3 = tile_index(idx, g0 + 3);
      if (lane == 0) mbar_expect_tx(bar, kChunkBytes);
      __syncwarp();
      tma_gather4(dst + uint32_t(lane) * 512u, tm, 0, r0, r1, r2, r3, bar);
      tma_gather4(dst + kPanel + uint32_t(lane) * 512u, tm, 64, r0, r1, r2, r3, bar);
    };
    const bool has_pr = a.proj_r != nullptr, has_ga = a.grad_agg != nullptr;
    if (my_tiles > 0) {
      load_e(0);
      if (pw == 0) {
        gather_tile(&tm_ps, a.senders, 0, buf(1, 0), &bars[kEgPsFull]);
        if (has_pr) gather_tile(&tm_pr, a.receivers, 0, buf(2, 0), &bars[kEgPrFull]);
      } else if (has_ga) {
        gather_tile(&tm_ga, a.receivers, 0, buf(3, 0), &bars[kEgGaFull]);
      }
    }
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
      mbar_wait(&bars[kEgG + 0], par, 50);
      colsum(buf(3, t), cs[0]);                              // dY
      __syncwarp(); if (lane == 0) mbar_arrive(&bars[kEgCs + 0]);
      mbar_wait(&bars[kEgG + 1], par, 51);
      colsum(buf(2, t), cs[1]);                              // dH2'
      __syncwarp(); if (lane == 0) mbar_arrive(&bars[kEgCs + 1]);
      mbar_wait(&bars[kEgAfree], par, 52);                   // H1 has been read back and the dW1 MMAs are done:
      if (more) load_e(t + 1);                               // buffer A(t) = S(t+1) takes the next tile's edge rows
      mbar_wait(&bars[kEgG + 2], par, 53);                   // dH1' = G0 is in buffer C (and fenced for the async proxy)
      if (ptid == 0 && !(a.ablate & 2)) {                    // G0 goes to HBM straight from the operand buffer
        const int y = int((blockIdx.x + t * gridDim.x) * kTile);
        tma_store_2d(&tm_g0, buf(3, t), 0, y);
        tma_store_2d(&tm_g0, buf(3, t) + kPanel, 64, y);
        tma_store_commit();
      }
      colsum(buf(3, t), cs[2]);
      if (ptid == 0) tma_store_wait_read<0>();               // the store has read the buffer before E1 of the next tile reuses it
      __syncwarp(); if (lane == 0) mbar_arrive(&bars[kEgCs + 2]);
      if (more) {
        // the next tile's table rows.  C(t) [= B(t+1)] and S(t) [= C(t+1)] are free once the dWe MMAs of step 5 have read them
        // (their own commit) and both producer warps have summed G0; B(t) [= A(t+1)] once the TMA store of d e has read it.
        mbar_wait(&bars[kEgWeDone], par, 54);
        fence_async_smem();                                   // our generic-proxy reads of C(t) before the async-proxy writes
        if (pw == 0) {
          if (has_pr) { mbar_wait(&bars[kEgCs + 2], par, 56); gather_tile(&tm_pr, a.receivers, t + 1, buf(2, t + 1), &bars[kEgPrFull]); }
          mbar_wait(&bars[kEgDeFree], par, 55);
          gather_tile(&tm_ps, a.senders, t + 1, buf(1, t + 1), &bars[kEgPsFull]);
        } else if (has_ga) {
          gather_tile(&tm_ga, a.receivers, t + 1, buf(3, t + 1), &bars[kEgGaFull]);
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
        mbar_spin(&bars[kEgEpi], epi_phase++ & 1, 60);
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
        mbar_wait(&bars[kEgFull], uint32_t(t) & 1, 61);
        fence_after_sync();
        stamp_g4(a, t, 0);
        if (!first) wait_epi();                                  // previous tile's last epilogue has drained the accumulator
        stamp_g4(a, t, 1);
        chain(S, sbase + kEgWe, false);  mma_commit(&bars[kEgAcc]);                                    // 0: e We^T
        stamp_g4(a, t, 2);
        wait_epi(); stamp_g4(a, t, 3); chain(A, sbase + kEgW1, false);  mma_commit(&bars[kEgAcc]);                        // 1: H1 W1^T
        wait_epi(); stamp_g4(a, t, 4); chain(B, sbase + kEgW2, false);  mma_commit(&bars[kEgAcc]);                        // 2: H2 W2^T
        // steps 3 and 4 commit after their weight-gradient MMAs: the epilogue warps spend that time on the LayerNorm vector
        // column sums anyway, and phases E3 / E4 overwrite buffers those MMAs read.  Step 5 commits right after the chain.
        wait_epi(); stamp_g4(a, t, 5); chain(C, sbase + kEgW2, true);   wgrad(dW2, C, B, first); mma_commit(&bars[kEgAcc]);   // 3: dY W2 ; dW2
        wait_epi(); stamp_g4(a, t, 6); chain(B, sbase + kEgW1, true);   wgrad(dW1, B, A, first); mma_commit(&bars[kEgAcc]);   // 4: dH2' W1 ; dW1
        wait_epi(); stamp_g4(a, t, 7); chain(C, sbase + kEgWe, true);   mma_commit(&bars[kEgAcc]); wgrad(dWe, C, S, first);   // 5: dH1' We ; dWe
        mma_commit(&bars[kEgWeDone]);                            // buffers C and S are free for the next tile's gathers
        stamp_g4(a, t, 8);
      }
      if (my_tiles > 0) mma_commit(&bars[kEgFinal]);          // every MMA of this CTA, for the accumulator drain
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
      mbar_wait(&bars[kEgAcc], acc_phase++ & 1, tag);         // try_wait (suspends): measured 2 % faster than busy polling here -- eight
                                                               // polling warps take issue slots from the column-sum warps and the MMA thread
      fence_after_sync();
    };
    // k-th column-sum hand-over of (local) tile tt: the producers have finished reading that tile from its buffer
    auto wait_cs = [&](int k, int64_t tt) { mbar_wait(&bars[kEgCs + k], uint32_t(tt) & 1, 70 + k); };
    auto done = [&](int producers_k) {   // producers_k >= 0: the tile just written is also the producers' column-sum input k
      fence_async_smem();          // generic-proxy tile writes -> visible to the tensor core's async-proxy reads
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars[kEgEpi]);
        if (producers_k >= 0) mbar_arrive(&bars[kEgG + producers_k]);
      }
    };
    const uint32_t row_off = hh * kPanel;                      // my 64 columns = panel hh of every buffer
    auto store_row = [&](uint32_t bufaddr, const uint32_t* w) {   // 64 columns (32 packed words) of row r
#pragma unroll
      for (int k = 0; k < 8; ++k) st_shared128(bufaddr + row_off + sw128_chunk(r, k), w + 4 * k);
    };
    auto tile_row = [&](int64_t tt) { return (blockIdx.x + tt * gridDim.x) * kTile + r; };
    const bool ld_on = !(a.ablate & 1);
    const bool epi_has_pr = a.proj_r != nullptr, epi_has_ga = a.grad_agg != nullptr;
    // the table rows of this thread's edge arrive by TMA in the tile buffers (producer warps); only the dense gradient row is still a
    // register load
    for (int64_t t = 0; t < my_tiles; ++t) {
      const uint32_t A = buf(1, t), B = buf(2, t), C = buf(3, t);
      // row-half exchange area (LayerNorm statistics): the first 4 KiB of buffer C, which is idle until dY is written into it
      float2* xch = reinterpret_cast<float2*>(smem + (C - sbase));
      const int64_t grow = tile_row(t);
      const bool valid = grow < rows;
      const bool has_do = valid && a.grad_out != nullptr && ld_on;
      const __nv_bfloat16* dorow = a.grad_out + grow * kD + hh * 64;
      // ---- E0: H1 = relu(e We^T + Ps[s] + Pr[r] + b0) -> A ---------------------------------------------------
      {
        // the TMA store of d e (previous tile) has read buffer A: the producers gather Ps[s] into it after that, so waiting for the
        // gathered rows covers the hand-over
        if (t > 0 && tid == 0) { tma_store_wait_read<0>(); mbar_arrive(&bars[kEgDeFree]); }
        mbar_wait(&bars[kEgPsFull], uint32_t(t) & 1, 78);
        if (epi_has_pr) mbar_wait(&bars[kEgPrFull], uint32_t(t) & 1, 79);
        wait_acc(100); if (tid == 0) stamp_g4(a, t, 10);
        uint32_t h[32];
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          uint32_t v[32], ps[16], pr[16];
          tmem_ld32(acc + cg * 32, v);
#pragma unroll
          for (int k = 0; k < 4; ++k) ld_shared128(A + row_off + sw128_chunk(r, 4 * cg + k), ps + 4 * k);       // Ps[s]: my row, in place
#pragma unroll
          for (int j = 0; j < 16; ++j) pr[j] = 0u;
          if (epi_has_pr) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ld_shared128(B + row_off + sw128_chunk(r, 4 * cg + k), pr + 4 * k);     // Pr[r]
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float2 x = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            x = __fadd2_rn(x, unpack_bf16x2(ps[j]));
            x = __fadd2_rn(x, unpack_bf16x2(pr[j]));
            x = __fadd2_rn(x, b0[cg * 16 + j]);
            h[cg * 16 + j] = cvt_relu_bf16x2(x.x, x.y);
          }
        }
        store_row(A, h);
        if (tid == 0) stamp_g4(a, t, 11);
        if (lane == 0) stamp_g4(a, t, 32 + warp);
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
        if (epi_has_ga) {                                       // grad_agg[receiver]: gathered into buffer C, idle until E2 writes dY into it
          mbar_wait(&bars[kEgGaFull], uint32_t(t) & 1, 80);
#pragma unroll
          for (int k = 0; k < 8; ++k) ld_shared128(C + row_off + sw128_chunk(r, k), dq + 32 + 4 * k);
        }
        // ---- E1: H2 = relu(H1 W1^T + b1) -> B -------------------------------------------------------------------
        if (t > 0) {
          wait_cs(2, t - 1);                                    // previous tile's dH1' column sum has left this buffer
        }
        wait_acc(101); if (tid == 0) stamp_g4(a, t, 12);
        uint32_t h[32];
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          uint32_t v[32];
          tmem_ld32(acc + cg * 32, v);
          tmem_ld_wait();
          if (tid == 0) stamp_g4(a, t, 26 + 2 * cg);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 x = __fadd2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), b1[cg * 16 + j]);
            h[cg * 16 + j] = cvt_relu_bf16x2(x.x, x.y);
          }
          if (tid == 0) stamp_g4(a, t, 27 + 2 * cg);
        }
        store_row(B, h);
        if (tid == 0) stamp_g4(a, t, 13);
        done(-1);
        // one rounding to bf16, the value every later use sees (EXPERIMENT: consumed after the phase's fence)
#pragma unroll
        for (int j = 0; j < 32; ++j) dreg[j] = add_bf16x2(dq[j], dq[32 + j]);
      }
      // ---- E2: y = H2 W2^T + b2 ; LayerNorm forward statistics and backward -> dY -> C --------------------------------
      uint32_t preg[32];                                        // dO * yhat (bf16): gamma-gradient terms, summed after the phase
      {
        wait_acc(102); if (tid == 0) stamp_g4(a, t, 14);
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
        if (tid == 0) stamp_g4(a, t, 15);
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
        wait_acc(103); if (tid == 0) stamp_g4(a, t, 16);
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
        if (tid == 0) stamp_g4(a, t, 17);
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
        wait_acc(104); if (tid == 0) stamp_g4(a, t, 18);
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
        __syncwarp(); if (lane == 0) mbar_arrive(&bars[kEgAfree]);   // this warp has read its H1 rows: the last reader of buffer A in this tile
        if (tid == 0) stamp_g4(a, t, 19);
        if (lane == 0) stamp_g4(a, t, 40 + warp);
        done(2);
      }
      // ---- E5: d e = dH1' We + dO -> HBM -------------------------------------------------------------------------------
      {
        wait_acc(105); if (tid == 0) stamp_g4(a, t, 20);
        uint32_t v0[32], v1[32];
        tmem_ld32(acc, v0);
        tmem_ld32(acc + 32, v1);
        tmem_ld_wait();
        // the accumulator is in registers: the next tile's step 0 may overwrite it while this phase does its arithmetic and
        // stores (no shared-memory writes here, so no proxy fence)
        fence_before_sync();
        __syncwarp(); if (lane == 0) mbar_arrive(&bars[kEgEpi]);
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
        __syncwarp(); if (lane == 0) mbar_arrive(&bars[kEgDe]);
        if (tid == 0) {                                         // one thread stores the staged tile once all 256 rows are in
          mbar_wait(&bars[kEgDe], uint32_t(t) & 1, 77);
          if (!(a.ablate & 2)) {
            const int y = int((blockIdx.x + t * gridDim.x) * kTile);
            tma_store_2d(&tm_de, B, 0, y);
            tma_store_2d(&tm_de, B + kPanel, 64, y);
            tma_store_commit();
          }
          stamp_g4(a, t, 21);
        }
      }
    }
    if (tid == 0) tma_store_wait<0>();
    // ---- drain the weight-gradient accumulators and the LayerNorm vector partials ---------------------------------------
    if (my_tiles > 0) {
      mbar_wait(&bars[kEgFinal], 0, 75);
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

// 2-D map over a [rows][128] bf16 table for tile::gather4: 64-column (128-byte) boxes, 128B swizzle.  The table's row count is not
// known to the launcher (the C ABI passes only the edge count), so the map spans 2^30 rows: the indices are the plan's, in range.
// Box height 1 (CuTe builds its gather4 descriptors the same way: make_tma_copy_atom composes the box to {columns, 1} and counts four
// boxes per instruction); HGN_GATHER4_BOX_ROWS overrides it for the bring-up probe.
typedef CUresult (*EncodeTiledFnG4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_gather_tensor_map(CUtensorMap* tm, const void* base) {
  static EncodeTiledFnG4 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    HGN_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (p == nullptr || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return HGN_ERR_CUDA; }
    fn = reinterpret_cast<EncodeTiledFnG4>(p);
  }
  const char* br = getenv("HGN_GATHER4_BOX_ROWS");
  const cuuint64_t gdim[2] = {cuuint64_t(kD), cuuint64_t(1) << 30};
  const cuuint64_t gstride[1] = {cuuint64_t(kD) * 2};
  const cuuint32_t box[2] = {64, cuuint32_t(br != nullptr ? atoi(br) : 1)};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult rc = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (gather4 table map) failed with %d (base %p)", int(rc), base); return HGN_ERR_CUDA; }
  return HGN_OK;
}

int edge_bwd_g4_launch(int64_t rows, int64_t tiles, int grid, const void* packed, const void* dense, const void* proj_s, const void* proj_r,
                       const int32_t* senders, const int32_t* receivers, int w0_chunks, int w0_chunk0, const void* grad_out, const void* grad_agg,
                       void* grad_dense, void* grad_pre0, float* w_partial, float* epi_colpart, float* prod_colpart, const char* name,
                       cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    uint32_t* dbg = debug_buffer_device();
    HGN_CUDA_OK(cudaMemcpyToSymbol(tc05::g_debug_words, &dbg, sizeof(dbg)));
    HGN_CUDA_OK(cudaFuncSetAttribute(edge_bwd_g4_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kEgSmem)));
    configured = true;
  }
  EdgeBwdG4Args a{};
  a.edge = static_cast<const __nv_bfloat16*>(dense);
  a.proj_s = static_cast<const __nv_bfloat16*>(proj_s);
  a.proj_r = static_cast<const __nv_bfloat16*>(proj_r);
  a.senders = senders;
  a.receivers = receivers;
  a.grad_out = static_cast<const __nv_bfloat16*>(grad_out);
  a.grad_agg = static_cast<const __nv_bfloat16*>(grad_agg);
  a.grad_edge = static_cast<__nv_bfloat16*>(grad_dense);
  a.grad_pre0 = static_cast<__nv_bfloat16*>(grad_pre0);
  a.w_partial = w_partial;
  a.epi_colpart = epi_colpart;
  a.prod_colpart = prod_colpart;
  a.w0_chunks = w0_chunks;
  a.w0_chunk0 = w0_chunk0;
  { const char* ab = getenv("HGN_TC_ABLATE"); a.ablate = ab ? atoi(ab) : 0; }
  const int64_t map_rows = rows > 0 ? rows : 1;      // rows == 0: maps over one (never accessed) row keep the encoder happy
  CUtensorMap tm_e, tm_g0, tm_de, tm_ps, tm_pr, tm_ga;
  if (int rc = make_rows_tensor_map(&tm_e, rows > 0 ? dense : w_partial, map_rows)) return rc;
  if (int rc = make_rows_tensor_map(&tm_g0, rows > 0 ? grad_pre0 : w_partial, map_rows)) return rc;
  if (int rc = make_rows_tensor_map(&tm_de, rows > 0 ? grad_dense : w_partial, map_rows)) return rc;
  if (int rc = make_gather_tensor_map(&tm_ps, proj_s != nullptr ? proj_s : w_partial)) return rc;
  if (int rc = make_gather_tensor_map(&tm_pr, proj_r != nullptr ? proj_r : w_partial)) return rc;
  if (int rc = make_gather_tensor_map(&tm_ga, grad_agg != nullptr ? grad_agg : w_partial)) return rc;
  static long long* tl_dev = nullptr;
  if (a.ablate & 64) {
    if (tl_dev == nullptr) cudaMalloc(&tl_dev, 8 * 48 * sizeof(long long));
    cudaMemsetAsync(tl_dev, 0, 8 * 48 * sizeof(long long), st);
    a.timeline = tl_dev;
  }
  {
    HGN_TIMED(name, st);
    edge_bwd_g4_tc_kernel<<<unsigned(grid), kEgThreads, kEgSmem, st>>>(rows, tiles, static_cast<const uint8_t*>(packed), a, tm_e, tm_g0, tm_de, tm_ps, tm_pr, tm_ga);
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
  return HGN_OK;
}

}  // namespace hgn
