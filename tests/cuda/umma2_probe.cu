// Stand-alone probe (test infrastructure) for the NEXT step of csrc/pointnet_mlp_tc.cu: tcgen05.mma.cta_group::2.
// A CTA pair (cluster of 2) computes D_r[128 x 128] = A_r[128 x 64] . B[128 x 64]^T for r = 0, 1 with ONE instruction
// stream issued by the leader CTA: each CTA keeps its own A (TMEM) and accumulator (TMEM) and only HALF of B
// (64 of the 128 rows) in its shared memory.  The probe checks the conventions (which half of B lives where, the
// M = 256 instruction descriptor, alloc / commit / dealloc with cta_group::2) against the CPU and then measures the
// sustained clocks per instruction.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma2_probe umma2_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

constexpr int M = 128, N = 128, K = 64;   // per CTA: 128 rows of A; N = 128 split 64 + 64 over the pair

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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;   // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}

// mode 0: correctness (one 256 x 128 x 64 product); mode 1: rate (n_mma instructions back to back)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe2(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int mode, int n_mma, unsigned long long* out_clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* btile = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // this CTA's half of B: 64 rows x 64 k, 8 KB
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;

  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // B half of this CTA: rows n = rank * 64 + r, r = tid (threads 0..63), 8 chunks of 8 bf16, 128-byte swizzle
  if (tid < 64) {
    const int n = (int)rank * 64 + tid;
    for (int c = 0; c < 8; ++c) {
      const uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)n * K + c * 8);
      *reinterpret_cast<uint4*>(btile + tid * 128 + ((c ^ (tid & 7)) * 16)) = v;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  // A rows of this CTA: global row = (pair * 2 + rank) * 128 + tid -> TMEM lane tid, columns [0, 32)
  {
    const __nv_bfloat16* a = A + ((size_t)(pair * 2 + rank) * M + tid) * K;
    uint32_t r[32];
    for (int j = 0; j < 32; ++j) {
      const uint16_t e = __bfloat16_as_ushort(a[2 * j]);
      const uint16_t o = __bfloat16_as_ushort(a[2 * j + 1]);
      r[j] = (uint32_t)o << 16 | e;
    }
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();   // both CTAs' operands are in place before the leader issues
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  unsigned long long t0 = 0;
  if (rank == 0 && warp == 0) {
    // instruction descriptor: D=f32 (bit 4), A=bf16 (bit 7), B=bf16 (bit 10), K-major both, N>>3 at 17, M>>4 at 24 with M = 256
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t bdesc = make_b_desc(smem_u32(btile));
    const uint32_t d_addr = tmem + 128;
    const int reps = mode == 0 ? 1 : n_mma / 4;
    t0 = clock64();
    for (int it = 0; it < reps; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < K / 16; ++s) {
          const uint32_t acc = (mode == 1) || s > 0;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_addr),
              "r"(tmem + s * 8), "l"(bdesc + (uint64_t)(s * 2)), "r"(idesc), "r"(acc)
              : "memory");
        }
      }
      __syncwarp();
    }
    if (elect_one()) {
      // completion to the barrier at this offset in BOTH CTAs
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                       smem_u32(&bar_mma)),
                   "h"((uint16_t)3)
                   : "memory");
    }
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (mode == 1 && rank == 0 && tid == 0) out_clk[pair] = clock64() - t0;
  if (mode == 0) {
    float* d = D + ((size_t)(pair * 2 + rank) * M + tid) * N;
    for (int q = 0; q < 4; ++q) {
      uint32_t r[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 128 + q * 32;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
          "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) d[q * 32 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();   // nobody frees tensor memory while the peer may still use the pair's allocation
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
  }
}

int main() {
  const int pairs_check = 1;
  std::vector<__nv_bfloat16> hA(2 * M * K), hB(N * K);
  std::vector<float> fA(2 * M * K), fB(N * K), ref(2 * M * N), out(2 * M * N);
  srand(11);
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 500.0f); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 700.0f); fB[i] = __bfloat162float(hB[i]); }
  for (int m = 0; m < 2 * M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
      ref[m * N + n] = (float)s;
    }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  unsigned long long* dClk;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int pairs_rate = sms / 2;
  CK(cudaMalloc(&dA, (size_t)pairs_rate * 2 * M * K * 2));
  CK(cudaMemset(dA, 0, (size_t)pairs_rate * 2 * M * K * 2));
  CK(cudaMalloc(&dB, N * K * 2));
  CK(cudaMalloc(&dD, (size_t)2 * M * N * 4));
  CK(cudaMalloc(&dClk, pairs_rate * 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, (size_t)2 * M * N * 4));
  CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  probe2<<<2 * pairs_check, 128, 10240>>>(dA, dB, dD, 0, 0, dClk);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr[2] = {0, 0}, maxref = 0;
  for (int m = 0; m < 2 * M; ++m)
    for (int n = 0; n < N; ++n) {
      maxerr[m / M] = fmax(maxerr[m / M], fabs(out[m * N + n] - ref[m * N + n]));
      maxref = fmax(maxref, fabs(ref[m * N + n]));
    }
  // which columns match, per CTA: tells how B halves map to N if the convention is different from the assumed one
  int col_ok[2][2] = {{0, 0}, {0, 0}};
  for (int r = 0; r < 2; ++r)
    for (int h = 0; h < 2; ++h) {
      double e = 0;
      for (int m = 0; m < M; ++m)
        for (int n = h * 64; n < h * 64 + 64; ++n) e = fmax(e, fabs(out[(r * M + m) * N + n] - ref[(r * M + m) * N + n]));
      col_ok[r][h] = e < 1e-3 * maxref;
    }
  printf("cta_group::2 256x128x64: max|err| CTA0 %.3e CTA1 %.3e of max|ref| %.3e -> %s  (column halves ok: CTA0 %d%d, CTA1 %d%d)\n",
         maxerr[0], maxerr[1], maxref, (maxerr[0] < 1e-3 * maxref && maxerr[1] < 1e-3 * maxref) ? "MATCH" : "mismatch",
         col_ok[0][0], col_ok[0][1], col_ok[1][0], col_ok[1][1]);
  const int n_mma = 8192;
  for (int rep = 0; rep < 2; ++rep) {
    probe2<<<2 * pairs_rate, 128, 10240>>>(dA, dB, dD, 1, n_mma, dClk);
    CK(cudaDeviceSynchronize());
  }
  std::vector<unsigned long long> clk(pairs_rate);
  CK(cudaMemcpy(clk.data(), dClk, pairs_rate * 8, cudaMemcpyDeviceToHost));
  double sum = 0;
  for (int i = 0; i < pairs_rate; ++i) sum += (double)clk[i] / n_mma;
  printf("rate: %.1f clk per cta_group::2 instruction (M=256 over the pair, N=128, K=16) on %d pairs; per SM that is the work of one 128x128x16 MMA (64 clk ideal)\n",
         sum / pairs_rate, pairs_rate);
  return 0;
}
