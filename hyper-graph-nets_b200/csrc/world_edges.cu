// World-edge radius search of the deforming-plate model -- replaces the dense N x N search of src/model/plate.py:86-110
//   world_distance_matrix = torch.cdist(world_pos, world_pos)            (4 TB of fp32 at 1 M nodes)
//   connect where distance < radius, off the diagonal, not already a mesh edge (senders, receivers),
//   sender an OBSTACLE node, receiver a NORMAL node;  world_senders, world_receivers = torch.nonzero(matrix)
// by a uniform-grid cell list: receivers are counting-sorted into cells of edge >= radius (+ the rounding margin of the
// distance formula below), every sender tests the 27 cells around its own, and the pairs leave in torch.nonzero's row-major
// order (ascending sender, then ascending receiver) as int64.  HBM-bound integer / fp32 work; no tensor cores.
//
// Bit-exact edge sets need the reference's distance arithmetic, not the textbook one.  torch.cdist (p = 2, more than 25 rows)
// evaluates  d(i, j) = sqrt(max(0, x1_[i] . x2_[j]))  with  x1_[i] = [-2 x_i, |x_i|^2, 1],  x2_[j] = [x_j, 1, |x_j|^2]  as one
// fp32 GEMM with K = 5: five fused multiply-adds in k order (checked against torch's CPU GEMM on 9 M pairs: 0 mismatches;
// the direct form sqrt(sum (x_i - x_j)^2) differs from it in most entries (65 % of 640 000 pairs in the CPU test)).  |x|^2 = (x^2 + y^2) + z^2 with every
// product and sum rounded (pow(2) and sum(-1) are separate ATen kernels).  we_distance() below is that arithmetic.
//
// Determinism: the only atomics are integer (cell counters, hash-table CAS, bounding box); the order in which receivers land
// inside a cell varies from run to run, and every sender's receiver list is sorted before it leaves, so the output does not.
#include <cub/cub.cuh>

#include "common.cuh"

namespace hgn {

namespace {

constexpr int kWeThreads = 256;
constexpr int kWeMaxDim = 1024;           // cells per axis: keeps the fp32 cell coordinate exact to ~2e-4 of a cell

struct WeHeader {
  int lo[3], hi[3];                        // receiver bounding box, order-preserving int image of the floats
  unsigned max_norm2;                      // max |x|^2 over senders and receivers (bits of a non-negative float)
  float h, inv_h, origin[3];
  int dim[3];
  int n_receivers;
  unsigned long long total;                // number of world edges
};

struct WeLayout { size_t header, cell_start, cell_cursor, cell_items, counts, offsets, table, cub, cub_bytes, total; int64_t cells, table_slots; };

int64_t pow2_at_least(int64_t x) { int64_t p = 1; while (p < x) p <<= 1; return p; }

WeLayout we_layout(int64_t n, int64_t e_mesh) {
  WeLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.cells = pow2_at_least(n * 2 < 64 ? 64 : (n * 2 > (int64_t(1) << 25) ? (int64_t(1) << 25) : n * 2));
  L.table_slots = pow2_at_least(e_mesh * 2 < 16 ? 16 : e_mesh * 2);
  L.header = take(sizeof(WeHeader));
  L.cell_start = take(size_t(L.cells + 1) * 4);
  L.cell_cursor = take(size_t(L.cells + 1) * 4);
  L.cell_items = take(size_t(n > 0 ? n : 1) * 4);
  L.counts = take(size_t(n + 1) * 4);
  L.offsets = take(size_t(n + 1) * 8);
  L.table = take(size_t(L.table_slots) * 8);
  size_t a = 0, b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, a, static_cast<int*>(nullptr), static_cast<int*>(nullptr), int(L.cells + 1));
  cub::DeviceScan::ExclusiveSum(nullptr, b, static_cast<int*>(nullptr), static_cast<long long*>(nullptr), int(n + 1));
  L.cub_bytes = a > b ? a : b;
  L.cub = take(L.cub_bytes);
  L.total = off;
  return L;
}

__device__ __forceinline__ int ordered_int(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float from_ordered_int(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ float we_norm2(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
// torch.cdist's fp32 arithmetic (see the file header): row i = sender, column j = receiver
__device__ __forceinline__ float we_distance(float xi, float yi, float zi, float ni, float xj, float yj, float zj, float nj) {
  float acc = __fmul_rn(-2.0f * xi, xj);
  acc = __fmaf_rn(-2.0f * yi, yj, acc);
  acc = __fmaf_rn(-2.0f * zi, zj, acc);
  acc = __fadd_rn(ni, acc);
  acc = __fadd_rn(acc, nj);
  return __fsqrt_rn(fmaxf(acc, 0.0f));
}
__device__ __forceinline__ bool we_finite(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

__global__ void we_init_kernel(WeHeader* hd) {
  for (int a = 0; a < 3; ++a) { hd->lo[a] = 0x7fffffff; hd->hi[a] = int(0x80000000u); }
  hd->max_norm2 = 0u;
  hd->n_receivers = 0;
  hd->total = 0ull;
}

__global__ void we_bounds_kernel(const float* __restrict__ pos, const int32_t* __restrict__ type, int64_t n, int sender_type,
                                 int receiver_type, WeHeader* hd) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {int(0x80000000u), int(0x80000000u), int(0x80000000u)};
  unsigned n2 = 0u;
  int is_recv = 0;
  if (i < n) {
    const int t = type[i];
    const float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    if ((t == sender_type || t == receiver_type) && we_finite(x, y, z)) {
      n2 = __float_as_uint(we_norm2(x, y, z));
      if (t == receiver_type) {
        is_recv = 1;
        lo[0] = hi[0] = ordered_int(x); lo[1] = hi[1] = ordered_int(y); lo[2] = hi[2] = ordered_int(z);
      }
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
    }
    n2 = max(n2, __shfl_xor_sync(0xffffffffu, n2, off));
    is_recv += __shfl_xor_sync(0xffffffffu, is_recv, off);
  }
  if ((threadIdx.x & 31) == 0) {
    if (is_recv) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { atomicMin(&hd->lo[a], lo[a]); atomicMax(&hd->hi[a], hi[a]); }
      atomicAdd(&hd->n_receivers, is_recv);
    }
    if (n2) atomicMax(&hd->max_norm2, n2);
  }
}

// one thread: cell edge and grid dimensions.  A pair the reference connects has computed distance < radius, hence a true
// distance below sqrt(radius^2 + 2 E2), E2 = 16 eps max|x|^2 bounding the rounding of the five-term dot product; the cell edge
// adds 0.1 % for the rounding of the cell coordinate itself, and grows until the grid fits `cells` cells / kWeMaxDim per axis.
__global__ void we_grid_kernel(WeHeader* hd, float radius, int64_t cells) {
  const float max_n2 = __uint_as_float(hd->max_norm2);
  const float e2 = 16.0f * 5.9604645e-8f * max_n2;
  float h = sqrtf(radius * radius + 2.0f * e2) * 1.001f + 8.0f * 5.9604645e-8f * sqrtf(max_n2);
  if (!(h > 0.0f) || !isfinite(h)) h = 1.0f;
  float lo[3], ext[3];
  for (int a = 0; a < 3; ++a) {
    if (hd->n_receivers > 0) { lo[a] = from_ordered_int(hd->lo[a]); ext[a] = from_ordered_int(hd->hi[a]) - lo[a]; }
    else { lo[a] = 0.0f; ext[a] = 0.0f; }
    hd->origin[a] = lo[a];
  }
  for (;;) {
    long long prod = 1;
    bool ok = true;
    for (int a = 0; a < 3; ++a) {
      const float cells_a = floorf(ext[a] / h) + 1.0f;
      if (!(cells_a <= float(kWeMaxDim))) { ok = false; break; }
      hd->dim[a] = int(cells_a);
      prod *= hd->dim[a];
    }
    if (ok && prod <= cells) break;
    h *= 1.25f;
  }
  hd->h = h;
  hd->inv_h = 1.0f / h;
}

__device__ __forceinline__ int we_cell_coord(float x, float origin, float inv_h, int dim) {
  // clamped to [-1, dim]: a receiver whose coordinate rounds up to `dim` is binned in cell dim - 1, so every sender at or beyond
  // `dim` has to look into that cell (a superset of its true neighbourhood; the distance test decides)
  const float t = floorf((x - origin) * inv_h);
  return int(fminf(fmaxf(t, -1.0f), float(dim)));
}

// pass 0: count receivers per cell; pass 1: place them (cell_cursor starts as a copy of cell_start)
__global__ void we_bin_kernel(const float* __restrict__ pos, const int32_t* __restrict__ type, int64_t n, int receiver_type,
                              const WeHeader* __restrict__ hd, int* __restrict__ counter, int* __restrict__ items, int pass) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || type[i] != receiver_type) return;
  const float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
  if (!we_finite(x, y, z)) return;
  const int cx = min(max(we_cell_coord(x, hd->origin[0], hd->inv_h, hd->dim[0]), 0), hd->dim[0] - 1);
  const int cy = min(max(we_cell_coord(y, hd->origin[1], hd->inv_h, hd->dim[1]), 0), hd->dim[1] - 1);
  const int cz = min(max(we_cell_coord(z, hd->origin[2], hd->inv_h, hd->dim[2]), 0), hd->dim[2] - 1);
  const int cell = (cx * hd->dim[1] + cy) * hd->dim[2] + cz;
  const int slot = atomicAdd(&counter[cell], 1);
  if (pass == 1) items[slot] = int(i);
}

__device__ __forceinline__ unsigned long long we_key(int64_t s, int64_t r) { return ((unsigned long long)(s) << 32 | (unsigned long long)(uint32_t)(r)) + 1ull; }
__device__ __forceinline__ int64_t we_hash(unsigned long long key, int64_t mask) { return int64_t((key * 0x9E3779B97F4A7C15ull) >> 20) & mask; }

// mesh edges (sender OBSTACLE, receiver NORMAL) into an open-addressing set: `world_connection_matrix[senders, receivers] = False`
__global__ void we_mesh_insert_kernel(const int64_t* __restrict__ ms, const int64_t* __restrict__ mr, int64_t e, const int32_t* __restrict__ type,
                                      int64_t n, int sender_type, int receiver_type, unsigned long long* table, int64_t mask) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= e) return;
  const int64_t s = ms[k], r = mr[k];
  if (s < 0 || s >= n || r < 0 || r >= n || type[s] != sender_type || type[r] != receiver_type) return;
  const unsigned long long key = we_key(s, r);
  int64_t slot = we_hash(key, mask);
  for (;;) {
    const unsigned long long old = atomicCAS(&table[slot], 0ull, key);
    if (old == 0ull || old == key) return;
    slot = (slot + 1) & mask;
  }
}
__device__ __forceinline__ bool we_is_mesh_edge(const unsigned long long* __restrict__ table, int64_t mask, int64_t s, int64_t r) {
  const unsigned long long key = we_key(s, r);
  int64_t slot = we_hash(key, mask);
  for (;;) {
    const unsigned long long v = table[slot];
    if (v == key) return true;
    if (v == 0ull) return false;
    slot = (slot + 1) & mask;
  }
}

// one thread per node: the sender's matches among the 27 cells around it.  kEmit = false counts them; kEmit = true writes
// them at offsets[i] and sorts the sender's receivers ascending (torch.nonzero order).
template <bool kEmit>
__global__ void we_search_kernel(const float* __restrict__ pos, const int32_t* __restrict__ type, int64_t n, int sender_type, float radius,
                                 const WeHeader* __restrict__ hd, const int* __restrict__ cell_start, const int* __restrict__ items,
                                 const unsigned long long* __restrict__ table, int64_t mask, int* __restrict__ counts,
                                 const long long* __restrict__ offsets, int64_t* __restrict__ out_s, int64_t* __restrict__ out_r) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int found = 0;
  const long long base = kEmit ? offsets[i] : 0;
  if (type[i] == sender_type && hd->n_receivers > 0) {
    const float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    if (we_finite(x, y, z)) {
      const float ni = we_norm2(x, y, z);
      const int cx = we_cell_coord(x, hd->origin[0], hd->inv_h, hd->dim[0]);
      const int cy = we_cell_coord(y, hd->origin[1], hd->inv_h, hd->dim[1]);
      const int cz = we_cell_coord(z, hd->origin[2], hd->inv_h, hd->dim[2]);
      for (int ax = max(cx - 1, 0); ax <= min(cx + 1, hd->dim[0] - 1); ++ax)
        for (int ay = max(cy - 1, 0); ay <= min(cy + 1, hd->dim[1] - 1); ++ay) {
          const int z0 = max(cz - 1, 0), z1 = min(cz + 1, hd->dim[2] - 1);
          if (z0 > z1) continue;
          const int row = (ax * hd->dim[1] + ay) * hd->dim[2];
          const int p0 = cell_start[row + z0], p1 = cell_start[row + z1 + 1];     // the z-neighbours are contiguous cells
          for (int p = p0; p < p1; ++p) {
            const int j = items[p];
            const float xj = pos[3 * j], yj = pos[3 * j + 1], zj = pos[3 * j + 2];
            const float d = we_distance(x, y, z, ni, xj, yj, zj, we_norm2(xj, yj, zj));
            if (d < radius && j != i && !we_is_mesh_edge(table, mask, i, j)) {
              if (kEmit) { out_s[base + found] = i; out_r[base + found] = j; }
              ++found;
            }
          }
        }
    }
  }
  if constexpr (!kEmit) {
    counts[i] = found;
  } else {
    for (int a = 1; a < found; ++a) {                 // insertion sort of this sender's receivers (a handful per sender)
      const int64_t v = out_r[base + a];
      int b = a - 1;
      while (b >= 0 && out_r[base + b] > v) { out_r[base + b + 1] = out_r[base + b]; --b; }
      out_r[base + b + 1] = v;
    }
  }
}

__global__ void we_total_kernel(const long long* __restrict__ offsets, int64_t n, WeHeader* hd) { hd->total = (unsigned long long)offsets[n]; }

}  // namespace

}  // namespace hgn

using namespace hgn;

extern "C" size_t hgn_world_edges_workspace_bytes(int64_t num_nodes, int64_t num_mesh_edges) {
  if (num_nodes < 0 || num_mesh_edges < 0) return 0;
  return we_layout(num_nodes, num_mesh_edges).total;
}

extern "C" int hgn_world_edges_count(const float* world_pos, const int32_t* node_type, int64_t num_nodes, const int64_t* mesh_senders,
                                     const int64_t* mesh_receivers, int64_t num_mesh_edges, float radius, int32_t sender_type,
                                     int32_t receiver_type, void* workspace, size_t workspace_bytes, int64_t* host_num_pairs, void* stream) {
  HGN_CHECK_ARG(num_nodes >= 0 && num_nodes < (int64_t(1) << 31) - 1 && num_mesh_edges >= 0, "world_edges: bad sizes");
  HGN_CHECK_ARG(workspace && host_num_pairs && (num_nodes == 0 || (world_pos && node_type)), "world_edges: null argument");
  HGN_CHECK_ARG(num_mesh_edges == 0 || (mesh_senders && mesh_receivers), "world_edges: null mesh edge list");
  HGN_CHECK_ARG(radius > 0.0f && sender_type != receiver_type, "world_edges: radius must be positive and the two node types distinct");
  const WeLayout L = we_layout(num_nodes, num_mesh_edges);
  if (workspace_bytes < L.total) { set_error("world_edges: workspace %zu < %zu", workspace_bytes, L.total); return HGN_ERR_WORKSPACE; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  WeHeader* hd = reinterpret_cast<WeHeader*>(ws + L.header);
  int* cell_start = reinterpret_cast<int*>(ws + L.cell_start);
  int* cell_cursor = reinterpret_cast<int*>(ws + L.cell_cursor);
  int* items = reinterpret_cast<int*>(ws + L.cell_items);
  int* counts = reinterpret_cast<int*>(ws + L.counts);
  long long* offsets = reinterpret_cast<long long*>(ws + L.offsets);
  unsigned long long* table = reinterpret_cast<unsigned long long*>(ws + L.table);
  const unsigned blocks = unsigned(ceil_div(num_nodes > 0 ? num_nodes : 1, kWeThreads));
  HGN_TIMED("world_edges_count", st);
  we_init_kernel<<<1, 1, 0, st>>>(hd);
  HGN_CUDA_OK(cudaMemsetAsync(cell_cursor, 0, size_t(L.cells + 1) * 4, st));
  HGN_CUDA_OK(cudaMemsetAsync(cell_start, 0, size_t(L.cells + 1) * 4, st));
  HGN_CUDA_OK(cudaMemsetAsync(counts, 0, size_t(num_nodes + 1) * 4, st));
  HGN_CUDA_OK(cudaMemsetAsync(table, 0, size_t(L.table_slots) * 8, st));
  if (num_nodes > 0) {
    we_bounds_kernel<<<blocks, kWeThreads, 0, st>>>(world_pos, node_type, num_nodes, sender_type, receiver_type, hd);
    we_grid_kernel<<<1, 1, 0, st>>>(hd, radius, L.cells);
    we_bin_kernel<<<blocks, kWeThreads, 0, st>>>(world_pos, node_type, num_nodes, receiver_type, hd, cell_cursor, items, 0);
    size_t cub_bytes = L.cub_bytes;
    HGN_CUDA_OK(cub::DeviceScan::ExclusiveSum(ws + L.cub, cub_bytes, cell_cursor, cell_start, int(L.cells + 1), st));
    HGN_CUDA_OK(cudaMemcpyAsync(cell_cursor, cell_start, size_t(L.cells) * 4, cudaMemcpyDeviceToDevice, st));
    we_bin_kernel<<<blocks, kWeThreads, 0, st>>>(world_pos, node_type, num_nodes, receiver_type, hd, cell_cursor, items, 1);
    if (num_mesh_edges > 0)
      we_mesh_insert_kernel<<<unsigned(ceil_div(num_mesh_edges, kWeThreads)), kWeThreads, 0, st>>>(mesh_senders, mesh_receivers, num_mesh_edges, node_type,
                                                                                               num_nodes, sender_type, receiver_type, table, L.table_slots - 1);
    we_search_kernel<false><<<blocks, kWeThreads, 0, st>>>(world_pos, node_type, num_nodes, sender_type, radius, hd, cell_start, items, table,
                                                          L.table_slots - 1, counts, nullptr, nullptr, nullptr);
  }
  {
    size_t cub_bytes = L.cub_bytes;
    HGN_CUDA_OK(cub::DeviceScan::ExclusiveSum(ws + L.cub, cub_bytes, counts, offsets, int(num_nodes + 1), st));
    we_total_kernel<<<1, 1, 0, st>>>(offsets, num_nodes, hd);
  }
  HGN_LAUNCH_OK("world_edges_count");
  unsigned long long total = 0;
  HGN_CUDA_OK(cudaMemcpyAsync(&total, &hd->total, sizeof(total), cudaMemcpyDeviceToHost, st));
  HGN_CUDA_OK(cudaStreamSynchronize(st));
  *host_num_pairs = int64_t(total);
  return HGN_OK;
}

extern "C" int hgn_world_edges_emit(const float* world_pos, const int32_t* node_type, int64_t num_nodes, int64_t num_mesh_edges, float radius,
                                    int32_t sender_type, const void* workspace, size_t workspace_bytes, int64_t* senders_out,
                                    int64_t* receivers_out, int64_t capacity, void* stream) {
  HGN_CHECK_ARG(num_nodes >= 0 && num_mesh_edges >= 0 && workspace, "world_edges_emit: bad arguments");
  const WeLayout L = we_layout(num_nodes, num_mesh_edges);
  if (workspace_bytes < L.total) { set_error("world_edges_emit: workspace %zu < %zu", workspace_bytes, L.total); return HGN_ERR_WORKSPACE; }
  if (num_nodes == 0 || capacity == 0) return HGN_OK;
  HGN_CHECK_ARG(world_pos && node_type && senders_out && receivers_out, "world_edges_emit: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const char* ws = static_cast<const char*>(workspace);
  const WeHeader* hd = reinterpret_cast<const WeHeader*>(ws + L.header);
  unsigned long long total = 0;                       // the count pass left it in the workspace: the caller's capacity must cover it
  HGN_CUDA_OK(cudaMemcpyAsync(&total, &hd->total, sizeof(total), cudaMemcpyDeviceToHost, st));
  HGN_CUDA_OK(cudaStreamSynchronize(st));
  HGN_CHECK_ARG(int64_t(total) <= capacity, "world_edges_emit: capacity %lld < %llu pairs", (long long)capacity, total);
  HGN_TIMED("world_edges_emit", st);
  we_search_kernel<true><<<unsigned(ceil_div(num_nodes, kWeThreads)), kWeThreads, 0, st>>>(
      world_pos, node_type, num_nodes, sender_type, radius, hd, reinterpret_cast<const int*>(ws + L.cell_start),
      reinterpret_cast<const int*>(ws + L.cell_items), reinterpret_cast<const unsigned long long*>(ws + L.table), L.table_slots - 1, nullptr,
      reinterpret_cast<const long long*>(ws + L.offsets), senders_out, receivers_out);
  HGN_LAUNCH_OK("world_edges_emit");
  return HGN_OK;
}
