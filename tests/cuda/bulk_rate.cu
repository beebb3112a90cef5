// Stand-alone probe (test infrastructure): how fast can one SM pull an L2-resident weight image into shared
// memory?  cp.async.bulk (1-D TMA) with P producer threads, copy size S and W copies in flight per producer,
// against plain ld.global.v4 + st.shared by 128..512 threads.  Prints bytes per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o bulk_rate bulk_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

constexpr int kRingBytes = 192 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

struct Cfg {
  int producers;  // threads issuing bulk copies (one per warp); 0 = ld.global/st.shared with all threads
  int copy_bytes;
  int window;     // copies in flight per producer
  int n_copies;   // per producer
  int hint;       // 1: L2 evict_last cache hint on the copies
};

__global__ void __launch_bounds__(512, 1) bulk_kernel(const uint8_t* img, size_t img_bytes, Cfg c, unsigned long long* out_clk) {
  extern __shared__ __align__(1024) uint8_t ring[];
  __shared__ uint64_t bars[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 64; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned long long t0 = clock64();
  if (c.producers > 0) {
    if (warp < c.producers && lane == 0) {
      uint64_t policy = 0;
      if (c.hint) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
      uint64_t* mybars = bars + warp * c.window;
      uint8_t* myring = ring + (size_t)warp * c.window * c.copy_bytes;
      size_t src = ((size_t)blockIdx.x * 4096 + (size_t)warp * 65536) % img_bytes;
      uint32_t phase_bits = 0;
      auto issue = [&](int i) {
        if (src + c.copy_bytes > img_bytes) src = 0;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mybars[i])), "r"(c.copy_bytes) : "memory");
        if (c.hint)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                           smem_u32(myring + (size_t)i * c.copy_bytes)),
                       "l"(img + src), "r"(c.copy_bytes), "r"(smem_u32(&mybars[i])), "l"(policy)
                       : "memory");
        else
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(myring + (size_t)i * c.copy_bytes)),
                       "l"(img + src), "r"(c.copy_bytes), "r"(smem_u32(&mybars[i]))
                       : "memory");
        src += c.copy_bytes;
      };
      for (int i = 0; i < c.window; ++i) issue(i);
      int i = 0;
      for (int n = c.window; n < c.n_copies; ++n) {
        mbar_wait(&mybars[i], (phase_bits >> i) & 1);
        phase_bits ^= 1u << i;
        issue(i);
        if (++i == c.window) i = 0;
      }
      for (int k = 0; k < c.window; ++k) {
        mbar_wait(&mybars[i], (phase_bits >> i) & 1);
        phase_bits ^= 1u << i;
        if (++i == c.window) i = 0;
      }
    }
  } else {
    // all threads: 16-byte loads, 4 in flight per thread, stored to shared memory
    const uint4* g = reinterpret_cast<const uint4*>(img);
    const size_t n16 = img_bytes / 16;
    uint4* s = reinterpret_cast<uint4*>(ring);
    size_t pos = ((size_t)blockIdx.x * 256) % n16;
    const int per_iter = blockDim.x * 4;
    const long long total16 = (long long)c.n_copies * c.copy_bytes / 16;
    for (long long done = 0; done < total16; done += per_iter) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        size_t j = pos + (size_t)k * blockDim.x + tid;
        if (j >= n16) j -= n16;
        v[k] = __ldg(g + j);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) s[(k * blockDim.x + tid) % (kRingBytes / 16)] = v[k];
      pos += per_iter;
      if (pos >= n16) pos -= n16;
    }
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (tid == 0) out_clk[blockIdx.x] = t1 - t0;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t img_bytes = 85 * 16384;
  uint8_t* img;
  CK(cudaMalloc(&img, img_bytes));
  CK(cudaMemset(img, 1, img_bytes));
  unsigned long long* d_clk;
  CK(cudaMalloc(&d_clk, sms * 8));
  CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes));
  struct Case { const char* name; Cfg c; int grid; int threads; };
  const Case cases[] = {
      {"bulk 1 producer 16 KB x 8     ", {1, 16384, 8, 2048, 0}, 0, 64},
      {"bulk 1 producer 16 KB x 8, 1 CTA", {1, 16384, 8, 2048, 0}, 1, 64},
      {"bulk 1 producer 16 KB x 2     ", {1, 16384, 2, 2048, 0}, 0, 64},
      {"bulk 1 producer 32 KB x 4     ", {1, 32768, 4, 1024, 0}, 0, 64},
      {"bulk 1 producer  4 KB x 16    ", {1, 4096, 16, 8192, 0}, 0, 64},
      {"bulk 1 producer  2 KB x 16    ", {1, 2048, 16, 16384, 0}, 0, 64},
      {"bulk 2 producers 16 KB x 4    ", {2, 16384, 4, 1024, 0}, 0, 64},
      {"bulk 4 producers 16 KB x 2    ", {4, 16384, 2, 512, 0}, 0, 128},
      {"bulk 4 producers  4 KB x 8    ", {4, 4096, 8, 2048, 0}, 0, 128},
      {"bulk 1 producer 16 KB x 8 evict_last", {1, 16384, 8, 2048, 1}, 0, 64},
      {"ldg+sts 128 threads           ", {0, 16384, 0, 2048, 0}, 0, 128},
      {"ldg+sts 512 threads           ", {0, 16384, 0, 2048, 0}, 0, 512},
      {"ldg+sts 512 threads, 1 CTA    ", {0, 16384, 0, 2048, 0}, 1, 512},
  };
  std::vector<unsigned long long> clk(sms);
  for (const Case& cs : cases) {
    const int grid = cs.grid > 0 ? cs.grid : sms;
    for (int rep = 0; rep < 2; ++rep) {
      bulk_kernel<<<grid, cs.threads, kRingBytes>>>(img, img_bytes, cs.c, d_clk);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(clk.data(), d_clk, grid * 8, cudaMemcpyDeviceToHost));
    double worst = 0, sum = 0;
    for (int i = 0; i < grid; ++i) { worst = std::max(worst, (double)clk[i]); sum += (double)clk[i]; }
    const double bytes = (double)std::max(cs.c.producers, 1) * cs.c.n_copies * cs.c.copy_bytes;
    printf("%s grid %3d: %.1f B/clk/SM (avg), %.1f (slowest CTA)\n", cs.name, grid, bytes / (sum / grid), bytes / worst);
  }
  return 0;
}
