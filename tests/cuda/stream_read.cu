// Stand-alone probe (test infrastructure): how fast can N CTAs x T threads read 164 MB from HBM with plain 128-bit loads,
// U of them in flight per thread?  Decides whether a 148-CTA x 8-warp kernel can be HBM-bound at all.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_read stream_read.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int U>
__global__ void reader(const float4* __restrict__ p, long long n, float* sink) {
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p + i + u * stride));
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) *sink = acc;
}

int main() {
  const long long bytes = 164ll << 20, n = bytes / 16;
  float4* d; float* sink; char* flush;
  cudaMalloc(&d, bytes); cudaMalloc(&sink, 4); cudaMalloc(&flush, 512 << 20);
  cudaMemset(d, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grids[] = {148, 296, 592, 2368}, threads[] = {256, 512, 1024};
  for (int g : grids) for (int t : threads) for (int u : {4, 8, 16}) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaMemset(flush, rep, 512 << 20);
      cudaEventRecord(e0);
      if (u == 4) reader<4><<<g, t>>>(d, n, sink); else if (u == 8) reader<8><<<g, t>>>(d, n, sink); else reader<16><<<g, t>>>(d, n, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("grid %4d x %4d threads, %2d loads in flight: %7.1f us  %6.0f GB/s\n", g, t, u, best * 1e3, bytes / best / 1e6);
  }
  return 0;
}
