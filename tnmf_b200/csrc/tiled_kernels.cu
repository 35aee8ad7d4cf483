// placeholder: replaced by the register-tiled kernels
#include "common.cuh"
namespace tnmf {
bool tiled_supported(const Geo &, int) { return false; }
size_t tiled_workspace_bytes(const Geo &) { return 0; }
int tiled_reconstruct(const Geo &, const float *, const float *, float *, const float *, double *, int *, cudaStream_t) { return TNMF_EUNSUPPORTED; }
int tiled_gradient_h(const Geo &, const float *, const float *, const float *, float *, float *, float *, double, const float *, double, const float *, double, cudaStream_t) { return TNMF_EUNSUPPORTED; }
int tiled_gradient_w(const Geo &, const float *, const float *, const float *, float *, float *, void *, size_t, cudaStream_t) { return TNMF_EUNSUPPORTED; }
}
