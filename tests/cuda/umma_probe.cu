// Stand-alone probe (test infrastructure): one 128x128x64 bf16 tcgen05.mma with the A operand in
// TMEM and the B operand in shared memory (K-major, 128-byte swizzle), checked against the CPU.
// It pins down, on real hardware, the conventions csrc/pointnet_mlp_tc.cu relies on:
//   * bf16 pairs packed into a 32-bit TMEM column: low half = even k
//   * the shared-memory matrix descriptor (SBO 1024 B, version 1, SWIZZLE_128B) and the +32 B
//     advance per 16-element k step
//   * the instruction descriptor bits for M=128, N=128, bf16 x bf16 -> f32
//   * cp.async.bulk global->shared completing on an mbarrier, tcgen05.commit, tcgen05.ld/st lane mapping
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

constexpr int M = 128, N = 128, K = 64;

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
        : "r"(smem_u32(bar)), "r"(parity));
  }
}

__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);   // start address
  d |= (uint64_t)1 << 16;                    // LBO (ignored for swizzled K-major), CUTLASS sets 1
  d |= (uint64_t)(1024 >> 4) << 32;          // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                    // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

// variant bit0: 0 = low half holds even k, 1 = low half holds odd k
// variant bit1: 0 = threads write B into smem, 1 = cp.async.bulk of a pre-swizzled global image
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, const uint8_t* Bimg,
                                            float* D, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // 16 KB tile; SWIZZLE_128B wants the tile base 1024-byte aligned in the shared window
  uint8_t* btile = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  __shared__ uint64_t bar_mma, bar_tma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_tma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;

  if (variant & 2) {
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tma)), "r"(16384));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(btile)),
                   "l"(Bimg), "r"(16384), "r"(smem_u32(&bar_tma))
                   : "memory");
    }
  } else {
    // row n = tid: 8 chunks of 8 bf16, chunk c lands at ((c ^ (n & 7)) * 16)
    for (int c = 0; c < 8; ++c) {
      const uint4 v = *reinterpret_cast<const uint4*>(B + tid * K + c * 8);
      *reinterpret_cast<uint4*>(btile + tid * 128 + ((c ^ (tid & 7)) * 16)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;");
  }

  // A row m = tid -> TMEM lane tid, columns [0, 32): two bf16 per column
  {
    uint32_t r[32];
    for (int j = 0; j < 32; ++j) {
      const uint16_t e = __bfloat16_as_ushort(A[tid * K + 2 * j]);
      const uint16_t o = __bfloat16_as_ushort(A[tid * K + 2 * j + 1]);
      r[j] = (variant & 1) ? ((uint32_t)e << 16 | o) : ((uint32_t)o << 16 | e);
    }
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]));
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();

  if (tid == 0) {
    if (variant & 2) mbar_wait(&bar_tma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    // instruction descriptor: D=f32 (bit4), A=bf16 (bit7), B=bf16 (bit10), K-major both, N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t bdesc = make_b_desc(smem_u32(btile));
    for (int s = 0; s < K / 16; ++s) {
      const uint32_t a_addr = tmem + s * 8;      // 16 bf16 = 8 columns
      const uint64_t b = bdesc + (uint64_t)((s * 32) >> 4);
      const uint32_t d_addr = tmem + 128;
      const uint32_t acc = s > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_addr),
          "r"(a_addr), "l"(b), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)));
  }
  mbar_wait(&bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
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
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 32; ++j) D[tid * N + q * 32 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

int main() {
  std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
  srand(7);
  for (int i = 0; i < M * K; ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 500.0f); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 700.0f); fB[i] = __bfloat162float(hB[i]); }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
      ref[m * N + n] = (float)s;
    }
  // pre-swizzled image of B as the bulk copy expects it in shared memory
  std::vector<uint8_t> img(16384);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < 8; ++c) memcpy(&img[n * 128 + ((c ^ (n & 7)) * 16)], &hB[n * K + c * 8], 16);
  __nv_bfloat16 *dA, *dB;
  uint8_t* dImg;
  float* dD;
  CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dImg, 16384)); CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dImg, img.data(), 16384, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  int ok_variants = 0;
  for (int v = 0; v < 4; ++v) {
    CK(cudaMemset(dD, 0, M * N * 4));
    probe<<<1, 128, 17408>>>(dA, dB, dImg, dD, v);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs(out[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
    printf("variant %d (pack %s, B via %s): max|err| %.3e of max|ref| %.3e -> %s\n", v, (v & 1) ? "low=odd" : "low=even",
           (v & 2) ? "cp.async.bulk" : "st.shared", maxerr, maxref, maxerr < 1e-3 * maxref ? "MATCH" : "mismatch");
    if (maxerr < 1e-3 * maxref) ok_variants |= 1 << v;
  }
  printf("RESULT mask=%d\n", ok_variants);
  return 0;
}
