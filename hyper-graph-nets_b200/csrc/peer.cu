// Halo exchange over NVLink peer memory (edge-cut partitioning; no reference counterpart: the reference is single
// device).  One process per GPU; every rank allocates its exchange pool with cudaMalloc, exports it with cudaIpcGetMemHandle and
// maps its peers' pools with cudaIpcOpenMemHandle (the 64-byte handles travel over torch.distributed).  A push is then ONE small
// kernel of this rank: it copies the boundary rows its peers need straight into their pools (ordinary stores through the NVLink
// mapping), fences, and the last CTA to finish raises a 4-byte flag in each peer's pool; the consumer's stream runs a one-warp
// kernel that spins on its own flags before the kernel that reads the rows.  No NCCL kernel, no host round trip, nothing that
// needs an SM of the receiving GPU while its persistent compute kernels hold all of them.
//
// Flags carry a monotonically increasing exchange number chosen by the (identical) host programs of all ranks, so they are never
// reset: a waiter proceeds as soon as flag >= its exchange number.
#include <cuda_runtime.h>
#include <string.h>

#include "common.cuh"

namespace hgn {

__global__ void halo_push_kernel(const uint4* __restrict__ table, const int32_t* __restrict__ send_index, hgn_halo_peers peers, int n_peers,
                                 int64_t total_rows, int row_u4, uint32_t epoch, unsigned int* done_counter) {
  // one warp per row: D * elem_size bytes = row_u4 16-byte pieces (128 bf16 = 16 pieces: half a warp)
  const int warps_per_block = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t i = int64_t(blockIdx.x) * warps_per_block + warp; i < total_rows; i += int64_t(gridDim.x) * warps_per_block) {
    int q = 0;
    while (q + 1 < n_peers && i >= peers.row_begin[q + 1]) ++q;
    const int64_t local = i - peers.row_begin[q];
    const uint4* src = table + int64_t(__ldg(send_index + i)) * row_u4;
    uint4* dst = reinterpret_cast<uint4*>(peers.dst[q]) + local * row_u4;
    for (int c = lane; c < row_u4; c += 32) dst[c] = __ldg(src + c);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done_counter, 1u);
    if (prev == gridDim.x - 1) {                 // every CTA's rows are fenced: raise the flags, re-arm the counter
      *done_counter = 0u;
      __threadfence_system();
      for (int q = 0; q < n_peers; ++q)
        if (peers.flag[q] != nullptr) *reinterpret_cast<volatile uint32_t*>(peers.flag[q]) = epoch;
    }
  }
}

__global__ void halo_wait_kernel(hgn_halo_flags flags, int n_flags, uint32_t epoch, uint32_t* timeout_word) {
  const int q = threadIdx.x;
  if (q < n_flags && flags.flag[q] != nullptr) {
    const volatile uint32_t* f = reinterpret_cast<const volatile uint32_t*>(flags.flag[q]);
    unsigned long long spins = 0;
    while (int32_t(*f - epoch) < 0) {            // wrap-safe "flag < epoch"
      if (++spins > (1ull << 31)) {              // seconds: a peer that never pushes must end in an error, not in a hung GPU
        if (timeout_word != nullptr) *timeout_word = 0xDEAD0000u | uint32_t(q);
        __trap();
      }
    }
  }
  __threadfence_system();
}

}  // namespace hgn

using namespace hgn;

extern "C" int hgn_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  HGN_CHECK_ARG(ptr != nullptr && handle64 != nullptr && bytes > 0, "peer_alloc: bad arguments");
  HGN_CUDA_OK(cudaMalloc(ptr, bytes));
  HGN_CUDA_OK(cudaMemset(*ptr, 0, bytes));
  cudaIpcMemHandle_t h;
  HGN_CUDA_OK(cudaIpcGetMemHandle(&h, *ptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  HGN_CUDA_OK(cudaDeviceSynchronize());
  return HGN_OK;
}

extern "C" int hgn_peer_open(const unsigned char* handle64, void** ptr) {
  HGN_CHECK_ARG(ptr != nullptr && handle64 != nullptr, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  HGN_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return HGN_OK;
}

extern "C" int hgn_peer_close(void* ptr) {
  if (ptr != nullptr) HGN_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return HGN_OK;
}

extern "C" int hgn_peer_free(void* ptr) {
  if (ptr != nullptr) HGN_CUDA_OK(cudaFree(ptr));
  return HGN_OK;
}

extern "C" int hgn_halo_push(int dtype, const void* table, const int32_t* send_index, int32_t D, const hgn_halo_peers* peers, int32_t n_peers,
                             uint32_t epoch, void* counter, void* stream) {
  HGN_CHECK_ARG(peers != nullptr && n_peers >= 0 && n_peers <= HGN_MAX_PEERS, "halo_push: n_peers=%d outside [0,%d]", n_peers, HGN_MAX_PEERS);
  HGN_CHECK_ARG(counter != nullptr, "halo_push: counter is NULL");
  const size_t elem = dtype == HGN_BF16 ? 2 : 4;
  HGN_CHECK_ARG((size_t(D) * elem) % 16 == 0, "halo_push: rows must be a multiple of 16 bytes");
  if (n_peers == 0) return HGN_OK;
  const int64_t total = peers->row_begin[n_peers];
  HGN_CHECK_ARG(total >= 0 && (total == 0 || (table != nullptr && send_index != nullptr)), "halo_push: null pointer");
  for (int q = 0; q < n_peers; ++q)
    HGN_CHECK_ARG(peers->row_begin[q] <= peers->row_begin[q + 1] && (peers->row_begin[q] == peers->row_begin[q + 1] || peers->dst[q] != nullptr),
                  "halo_push: peer %d has rows but no destination", q);
  const int warps = 8;
  int64_t blocks = (total + warps - 1) / warps;
  if (blocks < 1) blocks = 1;                    // flags are raised even when there is nothing to send (the peer still waits for them)
  if (blocks > 64) blocks = 64;                  // a few CTAs saturate one NVLink direction for these sizes; they slot in beside persistent kernels
  HGN_TIMED("halo_push", static_cast<cudaStream_t>(stream));
  halo_push_kernel<<<unsigned(blocks), warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(table), send_index, *peers, n_peers, total, int(size_t(D) * elem / 16), epoch, static_cast<unsigned int*>(counter));
  HGN_LAUNCH_OK("halo_push");
  return HGN_OK;
}

extern "C" int hgn_halo_wait(const hgn_halo_flags* flags, int32_t n_flags, uint32_t epoch, void* stream) {
  HGN_CHECK_ARG(flags != nullptr && n_flags >= 0 && n_flags <= HGN_MAX_PEERS, "halo_wait: n_flags=%d outside [0,%d]", n_flags, HGN_MAX_PEERS);
  if (n_flags == 0) return HGN_OK;
  HGN_TIMED("halo_wait", static_cast<cudaStream_t>(stream));
  halo_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(*flags, n_flags, epoch, nullptr);
  HGN_LAUNCH_OK("halo_wait");
  return HGN_OK;
}
