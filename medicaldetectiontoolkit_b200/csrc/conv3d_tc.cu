// tcgen05 implicit-GEMM conv3d — placeholder until the kernels land (returns "unsupported" so `auto` resolves to SIMT).
#include "conv3d_common.cuh"
namespace mdt {
bool conv_tc_supported(const ConvGeom &, int) { return false; }
size_t conv_tc_workspace_bytes(const ConvGeom &, int, int) { return 0; }
int conv_tc_fprop(const ConvGeom &, const float *, const float *, const float *, const float *, float *, int, int, void *, size_t, cudaStream_t) { return MDT_EUNSUPPORTED; }
int conv_tc_dgrad(const ConvGeom &, const float *, const float *, float *, int, void *, size_t, cudaStream_t) { return MDT_EUNSUPPORTED; }
int conv_tc_wgrad(const ConvGeom &, const float *, const float *, float *, float *, int, void *, size_t, cudaStream_t) { return MDT_EUNSUPPORTED; }
}  // namespace mdt
