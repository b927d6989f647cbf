// placeholder until the tcgen05 kernels land
#include "common.cuh"
namespace nerfca {
int tc_supported(const nerfca_field_t&) { set_error("bf16 tcgen05 path not built yet"); return NERFCA_E_UNSUPPORTED; }
size_t tc_stash_bytes(const nerfca_field_t&, long long) { return 0; }
size_t tc_workspace_bytes(const nerfca_field_t&, long long, int) { return 0; }
int tc_field_forward(const nerfca_field_t&, const nerfca_samples_t&, float*, void*, void*, cudaStream_t) { return NERFCA_E_UNSUPPORTED; }
int tc_field_backward(const nerfca_field_t&, const nerfca_samples_t&, const float*, const void*, void*, const nerfca_field_grads_t&, cudaStream_t) { return NERFCA_E_UNSUPPORTED; }
}
