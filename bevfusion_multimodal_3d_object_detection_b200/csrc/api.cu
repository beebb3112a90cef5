// C-ABI entry points that are not tied to one kernel file: version/error strings, device query and
// the precision dispatch of b200bev_pointnet_encode.
#include "common.cuh"

namespace b200bev {
int pointnet_encode_f32(const float* points, int B, int N, int C, const float* params, const int32_t* dims,
                        int n_layers, const int32_t* perm, const int32_t* offsets, int n_cells, float* out_global,
                        float* out_canvas, cudaStream_t st);
int pointnet_encode_tc(const float* points, int B, int N, int C, const float* params, const int32_t* dims,
                       int n_layers, const int32_t* perm, const int32_t* offsets, int n_cells, const void* tc_params,
                       float* out_global, float* out_canvas, cudaStream_t st);
int pointnet_encode_split(const float* points, int B, int N, int C, const int32_t* dims, int n_layers, const int32_t* perm,
                          const int32_t* offsets, int n_cells, const void* image, float* out_global, float* out_canvas,
                          void* workspace, size_t workspace_bytes, cudaStream_t st);
}  // namespace b200bev

using namespace b200bev;

extern "C" B200BEV_API int b200bev_abi_version(void) { return B200BEV_ABI_VERSION; }

extern "C" B200BEV_API const char* b200bev_error_string(int status) {
  switch (status) {
    case B200BEV_OK: return "ok";
    case B200BEV_ERR_INVALID_ARGUMENT: return "invalid argument (null pointer, non-positive size or misaligned buffer)";
    case B200BEV_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
    case B200BEV_ERR_K_OUT_OF_RANGE: return "selected index k out of range";
    case B200BEV_ERR_WORKSPACE: return "workspace too small or misaligned";
    default: break;
  }
  if (status >= B200BEV_ERR_CUDA) return cudaGetErrorString((cudaError_t)(status - B200BEV_ERR_CUDA));
  return "unknown b200bev status";
}

extern "C" B200BEV_API int b200bev_device_info(int* sm_count_out, int* cc_major, int* cc_minor) {
  int dev = 0;
  B200BEV_CUDA_TRY(cudaGetDevice(&dev));
  int v = 0;
  if (sm_count_out) {
    B200BEV_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    *sm_count_out = v;
  }
  if (cc_major) {
    B200BEV_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
    *cc_major = v;
  }
  if (cc_minor) {
    B200BEV_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
    *cc_minor = v;
  }
  return B200BEV_OK;
}

extern "C" B200BEV_API int b200bev_pointnet_encode(const float* points, int B, int N, int C, const float* params,
                                       const int32_t* dims, int n_layers, const int32_t* perm, const int32_t* offsets,
                                       int n_cells, int precision, const void* tc_params, float* out_global,
                                       float* out_canvas, void* stream) {
  if (!points || !params || !dims || B <= 0 || C <= 0) return B200BEV_ERR_INVALID_ARGUMENT;
  if (!out_global && !out_canvas) return B200BEV_ERR_INVALID_ARGUMENT;
  if (N <= 0) return B200BEV_ERR_INVALID_ARGUMENT;  // torch.max over an empty point axis raises
  if ((perm == nullptr) != (offsets == nullptr)) return B200BEV_ERR_INVALID_ARGUMENT;
  if (out_canvas && (!perm || n_cells <= 0)) return B200BEV_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == B200BEV_F32)
    return pointnet_encode_f32(points, B, N, C, params, dims, n_layers, perm, offsets, n_cells, out_global, out_canvas, st);
  if (precision == B200BEV_BF16_TENSOR) {
    if (!tc_params) return B200BEV_ERR_INVALID_ARGUMENT;
    return pointnet_encode_tc(points, B, N, C, params, dims, n_layers, perm, offsets, n_cells, tc_params, out_global,
                              out_canvas, st);
  }
  return B200BEV_ERR_INVALID_ARGUMENT;
}

extern "C" B200BEV_API int b200bev_pointnet_encode_split(const float* points, int B, int N, int C, const int32_t* dims, int n_layers,
                                             const int32_t* perm, const int32_t* offsets, int n_cells, const void* image,
                                             float* out_global, float* out_canvas, void* workspace, size_t workspace_bytes,
                                             void* stream) {
  if (!points || !dims || !image || !workspace || B <= 0 || C <= 0 || N <= 0) return B200BEV_ERR_INVALID_ARGUMENT;
  if (!out_global && !out_canvas) return B200BEV_ERR_INVALID_ARGUMENT;
  if ((perm == nullptr) != (offsets == nullptr)) return B200BEV_ERR_INVALID_ARGUMENT;
  if (out_canvas && (!perm || n_cells <= 0)) return B200BEV_ERR_INVALID_ARGUMENT;
  return pointnet_encode_split(points, B, N, C, dims, n_layers, perm, offsets, n_cells, image, out_global, out_canvas, workspace,
                               workspace_bytes, (cudaStream_t)stream);
}
