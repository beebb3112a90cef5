// S1b — tensor-core (tcgen05 / TMEM) path of the fused PointNet shared-MLP + max.
// Placeholder until the kernel lands: the entry points exist so the ABI is stable, and report
// B200BEV_ERR_UNSUPPORTED instead of silently computing something else.
#include "common.cuh"

namespace b200bev {

int pointnet_encode_tc(const float*, int, int, int, const float*, const int32_t*, int, const int32_t*, const int32_t*,
                       int, const void*, float*, float*, cudaStream_t) {
  return B200BEV_ERR_UNSUPPORTED;
}

}  // namespace b200bev

extern "C" B200BEV_API size_t b200bev_pointnet_pack_bf16_bytes(const int32_t*, int) { return 0; }

extern "C" B200BEV_API int b200bev_pointnet_pack_bf16(const float*, const int32_t*, int, void*, size_t, void*) {
  return B200BEV_ERR_UNSUPPORTED;
}
