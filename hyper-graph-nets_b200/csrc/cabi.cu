// C-ABI entry points of libhgn_b200.so that are not defined next to their kernels: error state,
// device probe, MLP-tile dispatch (fp32 FFMA parity path vs bf16 tcgen05 path) and host<->device copies.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace hgn {

static thread_local char g_error[512] = "";

static uint32_t* g_debug_host = nullptr;
static uint32_t* g_debug_dev = nullptr;

uint32_t* debug_buffer_device() {
  if (g_debug_dev == nullptr) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&g_debug_host), 256, cudaHostAllocMapped) == cudaSuccess) {
      memset(g_debug_host, 0, 256);
      if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_debug_dev), g_debug_host, 0) != cudaSuccess) g_debug_dev = nullptr;
    }
    cudaGetLastError();
  }
  return g_debug_dev;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  int n = vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  if (g_debug_host != nullptr && g_debug_host[0] == 0xB200DEADu && n > 0 && n < int(sizeof(g_error)) - 160) {
    const uint32_t* w = g_debug_host;
    snprintf(g_error + n, sizeof(g_error) - n, " [device stall record: tag=%u block=%u thread=%u words=%08x %08x %08x %08x %08x %08x]", w[1], w[2],
             w[3], w[4], w[5], w[6], w[7], w[8], w[9]);
  }
}

// ---- kernel timing registry ---------------------------------------------------------------------
struct TimedLaunch { const char* name; cudaEvent_t start, stop; };
static bool g_profile_on = false;
static std::vector<TimedLaunch> g_launches;
static std::mutex g_profile_mutex;

KernelTimer::KernelTimer(const char* name, cudaStream_t st) : name_(name), st_(st), start_(nullptr), on_(g_profile_on) {
  if (on_) {
    cudaEventCreate(&start_);
    cudaEventRecord(start_, st_);
  }
}
KernelTimer::~KernelTimer() {
  if (on_) {
    cudaEvent_t stop;
    cudaEventCreate(&stop);
    cudaEventRecord(stop, st_);
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    g_launches.push_back({name_, start_, stop});
  }
}

// fp32 path (mlp_f32.cu)
size_t mlp_f32_packed_bytes(int n_chunks);
int mlp_f32_pack(int n_chunks, const float* W0, const float* b0, const float* W1, const float* b1, const float* W2,
                 const float* b2, const float* gamma, const float* beta, void* packed, cudaStream_t st);
int mlp_f32_forward(int64_t rows, const hgn_chunks* ch, const void* packed, const void* resid, int64_t resid_off, void* out,
                    cudaStream_t st);
size_t mlp_f32_backward_workspace_bytes(int64_t rows, int n_chunks);
int mlp_f32_backward(int64_t rows, const hgn_chunks* ch, const void* packed, const void* grad_out, int resid_chunk, void* const* grad_chunk,
                     float* gW0, float* gb0, float* gW1, float* gb1, float* gW2, float* gb2, float* ggamma, float* gbeta,
                     void* workspace, size_t workspace_bytes, cudaStream_t st);
// bf16 tcgen05 path (mlp_tc.cu)
size_t mlp_tc_packed_bytes(int n_chunks);
int mlp_tc_pack(int n_chunks, const float* W0, const float* b0, const float* W1, const float* b1, const float* W2,
                const float* b2, const float* gamma, const float* beta, void* packed, cudaStream_t st);
int mlp_tc_forward(int64_t rows, const hgn_chunks* ch, const void* packed, const void* resid, int64_t resid_off, void* out,
                   cudaStream_t st);
size_t mlp_tc_backward_workspace_bytes(int64_t rows, int n_chunks);
int mlp_tc_backward(int64_t rows, const hgn_chunks* ch, const void* packed, const void* grad_out, int resid_chunk, void* const* grad_chunk,
                    float* gW0, float* gb0, float* gW1, float* gb1, float* gW2, float* gb2, float* ggamma, float* gbeta,
                    void* workspace, size_t workspace_bytes, cudaStream_t st);

// projected edge update (edge_tc.cu)
int edge_project_forward_tc(int64_t num_nodes, const void* v, const void* packed, void* proj_s, void* proj_r, cudaStream_t st);
size_t edge_project_backward_workspace_tc(int64_t num_nodes);
int edge_project_backward_tc(int64_t num_nodes, const void* v, const void* packed, const void* grad_s, const void* grad_r, const void* grad_v_add, void* grad_v,
                             float* grad_W0, void* workspace, size_t workspace_bytes, cudaStream_t st);
int edge_update_forward_tc(int64_t num_edges, const void* edge, const void* proj_s, const void* proj_r, const int32_t* senders,
                           const int32_t* receivers, const void* packed, void* out, cudaStream_t st);
size_t edge_update_backward_workspace_tc(int64_t num_edges);
int edge_update_backward_tc(int64_t num_edges, const void* edge, const void* proj_s, const void* proj_r, const int32_t* senders,
                            const int32_t* receivers, const void* packed, const void* grad_out,
                            const void* grad_agg, void* grad_edge, void* grad_pre0, float* gW0, float* gb0, float* gW1, float* gb1, float* gW2,
                            float* gb2, float* ggamma, float* gbeta, void* workspace, size_t workspace_bytes, cudaStream_t st);

int node_update_forward_tc(int64_t num_nodes, const void* v, int n_agg, const void* const* aggs, const void* packed, void* q1, void* q2,
                           void* out, cudaStream_t st);
size_t node_update_backward_workspace_tc(int64_t num_nodes);
int node_update_backward_tc(int64_t num_nodes, const void* v, int n_agg, const void* const* aggs, const void* q1, const void* q2,
                            const void* packed, const void* grad_out, void* grad_v, void* const* grad_aggs,
                            float* gW0, float* gb0, float* gW1, float* gb1, float* gW2, float* gb2, float* ggamma, float* gbeta,
                            void* workspace, size_t workspace_bytes, cudaStream_t st);

// An empty row set (rows == 0: an edge set without edges, e.g. deforming_plate world_edges in a frame without contact, plate.py:86-110)
// comes with NULL sources -- an empty device tensor has no storage -- so the per-chunk pointer check applies to rows > 0 only.
static int check_chunks(const hgn_chunks* ch, int64_t rows, const char* who) {
  HGN_CHECK_ARG(ch != nullptr, "%s: chunks is NULL", who);
  HGN_CHECK_ARG(ch->n_chunks >= 1 && ch->n_chunks <= HGN_MAX_CHUNKS, "%s: n_chunks=%d outside [1,%d]", who, ch->n_chunks, HGN_MAX_CHUNKS);
  if (rows > 0)
    for (int c = 0; c < ch->n_chunks; ++c) HGN_CHECK_ARG(ch->src[c] != nullptr, "%s: chunk %d has no source", who, c);
  return HGN_OK;
}

}  // namespace hgn

using namespace hgn;

extern "C" int hgn_abi_version(void) { return HGN_B200_ABI_VERSION; }
extern "C" const char* hgn_last_error(void) { return g_error; }

extern "C" int hgn_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return major == 10 ? 1 : 0;
}

extern "C" size_t hgn_mlp_packed_bytes(int dtype, int32_t n_chunks) {
  if (n_chunks < 1 || n_chunks > HGN_MAX_CHUNKS) return 0;
  return dtype == HGN_BF16 ? mlp_tc_packed_bytes(n_chunks) : mlp_f32_packed_bytes(n_chunks);
}

extern "C" int hgn_mlp_pack(int dtype, int32_t n_chunks, const float* W0, const float* b0, const float* W1, const float* b1,
                            const float* W2, const float* b2, const float* gamma, const float* beta, void* packed, void* stream) {
  HGN_CHECK_ARG(n_chunks >= 1 && n_chunks <= HGN_MAX_CHUNKS, "mlp_pack: n_chunks=%d", n_chunks);
  HGN_CHECK_ARG(W0 && b0 && W1 && b1 && W2 && b2 && gamma && beta && packed, "mlp_pack: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == HGN_F32) return mlp_f32_pack(n_chunks, W0, b0, W1, b1, W2, b2, gamma, beta, packed, st);
  if (dtype == HGN_BF16) return mlp_tc_pack(n_chunks, W0, b0, W1, b1, W2, b2, gamma, beta, packed, st);
  set_error("mlp_pack: unknown dtype %d", dtype);
  return HGN_ERR_INVALID_ARGUMENT;
}

extern "C" int hgn_mlp_forward(int dtype, int64_t rows, const hgn_chunks* chunks, const void* packed, const void* resid,
                               int64_t resid_row_offset, void* out, void* stream) {
  HGN_CHECK_ARG(rows >= 0 && rows < (int64_t(1) << 31), "mlp_forward: rows=%lld", (long long)rows);
  if (int rc = check_chunks(chunks, rows, "mlp_forward")) return rc;
  if (rows == 0) return HGN_OK;
  HGN_CHECK_ARG(packed && resid && out, "mlp_forward: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == HGN_F32) return mlp_f32_forward(rows, chunks, packed, resid, resid_row_offset, out, st);
  if (dtype == HGN_BF16) return mlp_tc_forward(rows, chunks, packed, resid, resid_row_offset, out, st);
  set_error("mlp_forward: unknown dtype %d", dtype);
  return HGN_ERR_INVALID_ARGUMENT;
}

extern "C" size_t hgn_mlp_backward_workspace_bytes(int dtype, int64_t rows, int32_t n_chunks) {
  if (n_chunks < 1 || n_chunks > HGN_MAX_CHUNKS || rows < 0) return 0;
  return dtype == HGN_BF16 ? mlp_tc_backward_workspace_bytes(rows, n_chunks) : mlp_f32_backward_workspace_bytes(rows, n_chunks);
}

extern "C" int hgn_mlp_backward(int dtype, int64_t rows, const hgn_chunks* chunks, const void* packed, const void* grad_out,
                                int32_t resid_chunk, void* const* grad_chunk, float* grad_W0, float* grad_b0, float* grad_W1, float* grad_b1,
                                float* grad_W2, float* grad_b2, float* grad_gamma, float* grad_beta, void* workspace,
                                size_t workspace_bytes, void* stream) {
  HGN_CHECK_ARG(rows >= 0 && rows < (int64_t(1) << 31), "mlp_backward: rows=%lld", (long long)rows);
  if (int rc = check_chunks(chunks, rows, "mlp_backward")) return rc;
  HGN_CHECK_ARG(packed && workspace && grad_W0 && grad_b0 && grad_W1 && grad_b1 && grad_W2 && grad_b2 && grad_gamma && grad_beta,
                "mlp_backward: null pointer");
  HGN_CHECK_ARG(rows == 0 || grad_out, "mlp_backward: grad_out is NULL");
  HGN_CHECK_ARG(resid_chunk >= -1 && resid_chunk < chunks->n_chunks, "mlp_backward: resid_chunk=%d", resid_chunk);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == HGN_F32)
    return mlp_f32_backward(rows, chunks, packed, grad_out, resid_chunk, grad_chunk, grad_W0, grad_b0, grad_W1, grad_b1, grad_W2, grad_b2,
                            grad_gamma, grad_beta, workspace, workspace_bytes, st);
  if (dtype == HGN_BF16)
    return mlp_tc_backward(rows, chunks, packed, grad_out, resid_chunk, grad_chunk, grad_W0, grad_b0, grad_W1, grad_b1, grad_W2, grad_b2,
                           grad_gamma, grad_beta, workspace, workspace_bytes, st);
  set_error("mlp_backward: unknown dtype %d", dtype);
  return HGN_ERR_INVALID_ARGUMENT;
}

#define HGN_BF16_ONLY(who)                                                                                   \
  do {                                                                                                        \
    if (dtype != HGN_BF16) {                                                                                  \
      set_error("%s: only HGN_BF16 is implemented (fp32 parity mode uses hgn_mlp_*)", who);                    \
      return HGN_ERR_UNSUPPORTED;                                                                             \
    }                                                                                                         \
  } while (0)

extern "C" int hgn_edge_project_forward(int dtype, int64_t num_nodes, const void* v, const void* packed, void* proj_s, void* proj_r,
                                        void* stream) {
  HGN_BF16_ONLY("edge_project_forward");
  HGN_CHECK_ARG(num_nodes >= 0 && num_nodes < (int64_t(1) << 31), "edge_project_forward: num_nodes=%lld", (long long)num_nodes);
  if (num_nodes == 0) return HGN_OK;
  HGN_CHECK_ARG(v && packed && proj_s && proj_r, "edge_project_forward: null pointer");
  return edge_project_forward_tc(num_nodes, v, packed, proj_s, proj_r, static_cast<cudaStream_t>(stream));
}

extern "C" size_t hgn_edge_project_backward_workspace_bytes(int dtype, int64_t num_nodes) {
  if (dtype != HGN_BF16 || num_nodes < 0) return 0;
  return edge_project_backward_workspace_tc(num_nodes);
}

extern "C" int hgn_edge_project_backward(int dtype, int64_t num_nodes, const void* v, const void* packed, const void* grad_s,
                                         const void* grad_r, const void* grad_v_add, void* grad_v, float* grad_W0, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  HGN_BF16_ONLY("edge_project_backward");
  HGN_CHECK_ARG(num_nodes >= 0 && num_nodes < (int64_t(1) << 31), "edge_project_backward: num_nodes=%lld", (long long)num_nodes);
  HGN_CHECK_ARG(packed && grad_W0 && workspace, "edge_project_backward: null pointer");
  HGN_CHECK_ARG(num_nodes == 0 || (v && grad_s && grad_r && grad_v), "edge_project_backward: null pointer");
  return edge_project_backward_tc(num_nodes, v, packed, grad_s, grad_r, grad_v_add, grad_v, grad_W0, workspace, workspace_bytes,
                                  static_cast<cudaStream_t>(stream));
}

extern "C" int hgn_edge_update_forward(int dtype, int64_t num_edges, const void* edge, const void* proj_s, const void* proj_r,
                                       const int32_t* senders, const int32_t* receivers, const void* packed, void* out,
                                       void* stream) {
  HGN_BF16_ONLY("edge_update_forward");
  HGN_CHECK_ARG(num_edges >= 0 && num_edges < (int64_t(1) << 31), "edge_update_forward: num_edges=%lld", (long long)num_edges);
  if (num_edges == 0) return HGN_OK;
  HGN_CHECK_ARG(edge && proj_s && proj_r && senders && receivers && packed && out, "edge_update_forward: null pointer");
  return edge_update_forward_tc(num_edges, edge, proj_s, proj_r, senders, receivers, packed, out, static_cast<cudaStream_t>(stream));
}

extern "C" size_t hgn_edge_update_backward_workspace_bytes(int dtype, int64_t num_edges) {
  if (dtype != HGN_BF16 || num_edges < 0) return 0;
  return edge_update_backward_workspace_tc(num_edges);
}

extern "C" int hgn_edge_update_backward(int dtype, int64_t num_edges, const void* edge, const void* proj_s, const void* proj_r,
                                        const int32_t* senders, const int32_t* receivers,
                                        const void* packed, const void* grad_out,
                                        const void* grad_agg, void* grad_edge, void* grad_pre0, float* grad_W0, float* grad_b0,
                                        float* grad_W1, float* grad_b1, float* grad_W2, float* grad_b2, float* grad_gamma,
                                        float* grad_beta, void* workspace, size_t workspace_bytes, void* stream) {
  HGN_BF16_ONLY("edge_update_backward");
  HGN_CHECK_ARG(num_edges >= 0 && num_edges < (int64_t(1) << 31), "edge_update_backward: num_edges=%lld", (long long)num_edges);
  HGN_CHECK_ARG(packed && workspace && grad_W0 && grad_b0 && grad_W1 && grad_b1 && grad_W2 && grad_b2 && grad_gamma && grad_beta,
                "edge_update_backward: null pointer");
  HGN_CHECK_ARG(num_edges == 0 || (edge && receivers && grad_edge && grad_pre0 && proj_s && proj_r && senders),
                "edge_update_backward: null pointer");
  return edge_update_backward_tc(num_edges, edge, proj_s, proj_r, senders, receivers, packed, grad_out, grad_agg, grad_edge, grad_pre0,
                                 grad_W0, grad_b0, grad_W1, grad_b1, grad_W2, grad_b2, grad_gamma, grad_beta, workspace, workspace_bytes,
                                 static_cast<cudaStream_t>(stream));
}

extern "C" int hgn_node_update_forward(int dtype, int64_t num_nodes, const void* v, int32_t n_agg, const void* const* aggs, const void* packed,
                                       void* q1, void* q2, void* out, void* stream) {
  HGN_BF16_ONLY("node_update_forward");
  HGN_CHECK_ARG(num_nodes >= 0 && num_nodes < (int64_t(1) << 31), "node_update_forward: num_nodes=%lld", (long long)num_nodes);
  HGN_CHECK_ARG(n_agg >= 1 && n_agg <= 4 && aggs != nullptr, "node_update_forward: n_agg=%d outside [1,4]", n_agg);
  if (num_nodes == 0) return HGN_OK;
  for (int j = 0; j < n_agg; ++j) HGN_CHECK_ARG(aggs[j] != nullptr, "node_update_forward: aggregate %d is NULL", j);
  HGN_CHECK_ARG(v && packed && q1 && out && (n_agg <= 2 || q2), "node_update_forward: null pointer");
  return node_update_forward_tc(num_nodes, v, n_agg, aggs, packed, q1, q2, out, static_cast<cudaStream_t>(stream));
}

extern "C" size_t hgn_node_update_backward_workspace_bytes(int dtype, int64_t num_nodes) {
  if (dtype != HGN_BF16 || num_nodes < 0) return 0;
  return node_update_backward_workspace_tc(num_nodes);
}

extern "C" int hgn_node_update_backward(int dtype, int64_t num_nodes, const void* v, int32_t n_agg, const void* const* aggs, const void* q1,
                                        const void* q2, const void* packed, const void* grad_out, void* grad_v,
                                        void* const* grad_aggs, float* grad_W0, float* grad_b0, float* grad_W1, float* grad_b1, float* grad_W2,
                                        float* grad_b2, float* grad_gamma, float* grad_beta, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  HGN_BF16_ONLY("node_update_backward");
  HGN_CHECK_ARG(num_nodes > 0 && num_nodes < (int64_t(1) << 31), "node_update_backward: num_nodes=%lld", (long long)num_nodes);
  HGN_CHECK_ARG(n_agg >= 1 && n_agg <= 4 && aggs != nullptr && grad_aggs != nullptr, "node_update_backward: n_agg=%d outside [1,4]", n_agg);
  for (int j = 0; j < n_agg; ++j) HGN_CHECK_ARG(aggs[j] != nullptr && grad_aggs[j] != nullptr, "node_update_backward: aggregate %d is NULL", j);
  HGN_CHECK_ARG(v && packed && grad_out && grad_v && workspace, "node_update_backward: null pointer");
  HGN_CHECK_ARG(q1 && (n_agg <= 2 || q2), "node_update_backward: needs the q tables of the forward");
  HGN_CHECK_ARG(grad_W0 && grad_b0 && grad_W1 && grad_b1 && grad_W2 && grad_b2 && grad_gamma && grad_beta, "node_update_backward: null pointer");
  return node_update_backward_tc(num_nodes, v, n_agg, aggs, q1, q2, packed, grad_out, grad_v, grad_aggs, grad_W0, grad_b0, grad_W1,
                                 grad_b1, grad_W2, grad_b2, grad_gamma, grad_beta, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int hgn_profile_enable(int on) { g_profile_on = on != 0; return HGN_OK; }

extern "C" int hgn_profile_reset(void) {
  std::lock_guard<std::mutex> lock(g_profile_mutex);
  for (auto& l : g_launches) { cudaEventDestroy(l.start); cudaEventDestroy(l.stop); }
  g_launches.clear();
  return HGN_OK;
}

extern "C" size_t hgn_profile_report(char* buf, size_t buf_bytes) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<long, double>> agg;
  std::vector<std::string> order;
  {
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    for (auto& l : g_launches) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, l.start, l.stop) != cudaSuccess) { cudaGetLastError(); continue; }
      auto it = agg.find(l.name);
      if (it == agg.end()) { order.push_back(l.name); agg[l.name] = {1, ms}; }
      else { it->second.first += 1; it->second.second += ms; }
    }
  }
  std::string out = "[";
  for (size_t i = 0; i < order.size(); ++i) {
    char line[256];
    snprintf(line, sizeof(line), "%s{\"name\": \"%s\", \"launches\": %ld, \"ms\": %.6f}", i ? ", " : "", order[i].c_str(),
             agg[order[i]].first, agg[order[i]].second);
    out += line;
  }
  out += "]";
  if (buf && buf_bytes) {
    const size_t n = out.size() < buf_bytes - 1 ? out.size() : buf_bytes - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return out.size() + 1;
}

extern "C" int hgn_copy_h2d(void* dst_device, const void* src_host, size_t bytes, void* stream) {
  HGN_CUDA_OK(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  return HGN_OK;
}
extern "C" int hgn_copy_d2h(void* dst_host, const void* src_device, size_t bytes, void* stream) {
  HGN_CUDA_OK(cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  return HGN_OK;
}
