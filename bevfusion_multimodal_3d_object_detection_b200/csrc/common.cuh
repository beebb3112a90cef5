// Shared helpers for the b200bev kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "b200bev.h"

#ifndef __CUDA_ARCH_LIST__
#define __CUDA_ARCH_LIST__ 1000
#endif

namespace b200bev {

constexpr unsigned FULL_MASK = 0xffffffffu;

// Launch-error → status code; never throws across the ABI.
inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? B200BEV_OK : B200BEV_ERR_CUDA + (int)e; }

#define B200BEV_CUDA_TRY(expr)                                   \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return B200BEV_ERR_CUDA + (int)_e;    \
  } while (0)

inline int launch_status() { return cuda_status(cudaGetLastError()); }

// Experiment and trace switches (B200BEV_TC_TRACE, B200BEV_DECODE_TRACE, B200BEV_CONV_IMPL, ...) exist only in a library
// built with -DB200BEV_DEBUG_ENV (`python -m bevfusion_multimodal_3d_object_detection_b200.build --debug-env`).  In the
// release library this is a constant null: no getenv on a launch path, and the trace branches — which allocate and
// synchronise, i.e. cannot run under CUDA-graph capture — are compiled out.
#ifdef B200BEV_DEBUG_ENV
inline const char* debug_env(const char* name) { return getenv(name); }
#else
inline const char* debug_env(const char*) { return nullptr; }
#endif

// SM count of the current device, cached per thread (re-entrant: no shared mutable global).
inline int sm_count() {
  thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Order-preserving float <-> uint32 key (larger float -> larger key). -0.0 is folded onto +0.0 so that
// the two compare equal, as they do for torch.topk.
__device__ __forceinline__ uint32_t float_to_key(float f) {
  uint32_t b = __float_as_uint(f);
  if (b == 0x80000000u) b = 0u;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(FULL_MASK, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// Split-fp16 tensor-core paths (fp32 accuracy from three fp16 products): 2^k with (tensor maximum) * 2^k in [2^13, 2^14), so
// that fp16 holds the scaled hi part (< 65504) and the lo part of anything that matters (lo is subnormal only below 2^-17
// of the maximum).  `max_bits` = the float bits of max|x|, reduced on the device; 1 for an all-zero tensor.
__host__ __device__ inline float activation_scale(uint32_t max_bits) {
  const int e = (int)((max_bits >> 23) & 0xff);
  if (e == 0 || e == 0xff) return 1.f;
  const uint32_t bits = (uint32_t)(127 + 14 - (e - 126)) << 23;   // max = f * 2^(e-126), f in [0.5, 1)
#ifdef __CUDA_ARCH__
  return __uint_as_float(bits);
#else
  union { uint32_t u; float f; } c;
  c.u = bits;
  return c.f;
#endif
}

// Streaming 128-bit load that does not allocate in L1 (data touched once).
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace b200bev
