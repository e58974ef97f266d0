// placeholder until the tcgen05 path lands
#include "common.cuh"
namespace hgn {
size_t mlp_tc_packed_bytes(int) { return 256; }
int mlp_tc_pack(int, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const float*, void*, cudaStream_t) { set_error("bf16 path not built"); return HGN_ERR_UNSUPPORTED; }
int mlp_tc_forward(int64_t, const hgn_chunks*, const void*, const void*, int64_t, void*, cudaStream_t) { set_error("bf16 path not built"); return HGN_ERR_UNSUPPORTED; }
size_t mlp_tc_backward_workspace_bytes(int64_t, int) { return 256; }
int mlp_tc_backward(int64_t, const hgn_chunks*, const void*, const void*, int, void* const*, float*, float*, float*, float*, float*, float*, float*, float*, void*, size_t, cudaStream_t) { set_error("bf16 path not built"); return HGN_ERR_UNSUPPORTED; }
}
