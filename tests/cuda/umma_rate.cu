// Stand-alone probe (test infrastructure): sustained issue/execute rate of tcgen05.mma on sm_100a under the
// conditions csrc/pointnet_mlp_tc.cu creates — A operand in tensor memory, B operand in shared memory
// (K-major, 128-byte swizzle), one CTA per SM — as a function of N, of the accumulator pattern and of a
// concurrent cp.async.bulk weight stream into the same shared memory.  Prints clk per MMA; the ideal is
// N/2 clk for M=128 (128 x N x 16 MACs at 4096 MAC/clk/SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate umma_rate.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

constexpr int kStageBytes = 16384;
constexpr int kStages = 12;

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
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
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
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Mode {
  int n;          // MMA N (64, 128, 256)
  int n_acc;      // accumulators cycled through (1, 2 or 4); each n columns wide
  int acc_run;    // consecutive MMAs into one accumulator before moving to the next
  int ss;         // 1: A from shared memory instead of tensor memory
  int stream;     // 1: a producer thread streams 16 KB stages into the ring while the MMAs run
  int b_stages;   // how many ring stages the MMAs read from (cycled every 4 MMAs)
  int epi;        // 1: four warps hammer shared memory with st.shared/ld.shared meanwhile (epilogue transposes)
  int fence;      // 1: tcgen05.fence::after_thread_sync before every group of 4 MMAs (as after a barrier wait)
  int poll;       // 1: an mbarrier try_wait on an already completed barrier before every group
  int ldtm;       // 1: the four other warps read the accumulators with tcgen05.ld all the time (layer-5 epilogue)
  int a_walk;     // 1: the A operand walks over 256 TMEM columns (K = 512) instead of staying on 32
};

__global__ void __launch_bounds__(192, 1) rate_kernel(const uint8_t* wimg, int n_img_stages, Mode m, int n_mma,
                                                      unsigned long long* out_clk, unsigned long long* out_stream) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* scratch = ring + (size_t)kStages * kStageBytes;   // 16 KB for the epi traffic
  __shared__ uint64_t bar_done, bar_full[kStages], bar_dummy, bar_ready;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop_flag;
  __shared__ volatile int mma_progress;   // stages consumed by the MMA thread: the stream is rate-matched to it
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&bar_done, 1);
    mbar_init(&bar_ready, 1);
    mbar_init(&bar_dummy, (1 << 20) - 1);
    for (int i = 0; i < kStages; ++i) mbar_init(&bar_full[i], 1);
    stop_flag = 0;
    mma_progress = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // something finite in the ring
  for (int i = tid; i < kStages * kStageBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ring)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_ready)) : "memory");   // phase 0 of bar_ready is complete
  __syncthreads();

  if (warp == 0) {
    // producer: keeps `depth` bulk copies in flight into ring stages [b_stages, kStages)
    if (lane == 0 && m.stream) {
      const int first = m.b_stages, nfree = kStages - m.b_stages;
      unsigned long long copies = 0;
      uint32_t phase_bits = 0;
      int issued = 0;
      // prime
      for (int i = 0; i < nfree; ++i) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_full[first + i])), "r"(kStageBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(ring + (size_t)(first + i) * kStageBytes)),
                     "l"(wimg + (size_t)((issued++) % n_img_stages) * kStageBytes), "r"(kStageBytes), "r"(smem_u32(&bar_full[first + i]))
                     : "memory");
      }
      int i = 0;
      while (!stop_flag) {
        if (m.stream == 1 && issued > mma_progress + nfree) continue;   // one stage per 4 MMAs, as in the real kernel
        mbar_wait(&bar_full[first + i], (phase_bits >> i) & 1);
        phase_bits ^= 1u << i;
        ++copies;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_full[first + i])), "r"(kStageBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(ring + (size_t)(first + i) * kStageBytes)),
                     "l"(wimg + (size_t)((issued++) % n_img_stages) * kStageBytes), "r"(kStageBytes), "r"(smem_u32(&bar_full[first + i]))
                     : "memory");
        if (++i == nfree) i = 0;
      }
      // drain what is in flight before the CTA exits
      for (int k = 0; k < nfree; ++k) {
        mbar_wait(&bar_full[first + i], (phase_bits >> i) & 1);
        if (++i == nfree) i = 0;
      }
      out_stream[blockIdx.x] = copies;
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(m.n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t ring_u = __shfl_sync(0xffffffffu, smem_u32(ring), 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    // accumulators at the top of TMEM, A operand columns [0, 32)
    // The whole warp runs the loop and one elected lane issues (warp-uniform operands live in uniform
    // registers; a loop run by a single divergent lane pays an R2UR waterfall per MMA — measured 156 clk).
    unsigned long long t0 = clock64(), t1 = 0;
    int acc = 0, run = 0, bst = 0;
    for (int i = 0; i < n_mma; i += 4) {
      const uint32_t d = tmem_u + 512 - (uint32_t)(acc + 1) * m.n;
      const uint64_t bdesc = make_desc(ring_u + bst * kStageBytes);
      const uint64_t adesc = make_desc(ring_u + ((bst + 1) % m.b_stages) * kStageBytes);
      if (m.poll) mbar_wait(&bar_ready, 0);
      if (m.fence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_col = m.a_walk ? (uint32_t)((i & 63) * 8 / 4 * 4) % 256u : 0u;
      if (elect_one()) {
        if (m.ss) {
#pragma unroll
          for (int s = 0; s < 4; ++s) umma_ss(d, adesc + (uint64_t)(s * 2), bdesc + (uint64_t)(s * 2), idesc, 1u);
        } else {
#pragma unroll
          for (int s = 0; s < 4; ++s) umma_ts(d, tmem_u + a_col + s * 8, bdesc + (uint64_t)(s * 2), idesc, 1u);
        }
        tc_commit(&bar_dummy);
        mma_progress = (i + 4) >> 2;
      }
      __syncwarp();
      if (++bst == m.b_stages) bst = 0;
      run += 4;
      if (run >= m.acc_run) {
        run = 0;
        if (++acc == m.n_acc) acc = 0;
      }
    }
    if (elect_one()) tc_commit(&bar_done);
    __syncwarp();
    mbar_wait(&bar_done, 0);
    t1 = clock64();
    if (lane == 0) {
      out_clk[blockIdx.x] = t1 - t0;
      stop_flag = 1;
    }
  } else if (m.ldtm) {
    // warps 2..5 <-> TMEM lane quadrants 2,3,0,1: read 32 accumulator columns over and over
    const uint32_t tm = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sink = 0;
    while (!stop_flag) {
#pragma unroll 1
      for (int q = 0; q < 8; ++q) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
            "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tm + 256 + q * 32)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sink ^= r[lane & 31];
      }
    }
    if (sink == 0x12345678u) out_clk[1] = 0;
  } else if (m.epi) {
    // 4 warps: conflict-free 128-bit st.shared + ld.shared on a 16 KB scratch until told to stop
    float4* s4 = reinterpret_cast<float4*>(scratch);
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    const int t = tid - 64;
    while (!stop_flag) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s4[t + k * 128] = v;
        const float4 r = s4[(t + 37) % 128 + k * 128];
        v.x += r.y;
      }
    }
    if (v.x == 123.456f) out_clk[0] = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int n_img = 85;
  uint8_t* wimg;
  CK(cudaMalloc(&wimg, (size_t)n_img * kStageBytes));
  CK(cudaMemset(wimg, 0x3c, (size_t)n_img * kStageBytes));
  unsigned long long *d_clk, *d_str;
  CK(cudaMalloc(&d_clk, sms * 8));
  CK(cudaMalloc(&d_str, sms * 8));
  const size_t smem = 1024 + (size_t)kStages * kStageBytes + 16384;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  struct Case { const char* name; Mode m; int grid; };
  const Case cases[] = {
      {"TS N=128 1 acc            ", {128, 1, 4, 0, 0, 8, 0}},
      {"TS N=128 2 acc alternating", {128, 2, 4, 0, 0, 8, 0}},
      {"TS N=128 2 acc runs of 32 ", {128, 2, 32, 0, 0, 8, 0}},
      {"TS N=256 1 acc            ", {256, 1, 1, 0, 0, 8, 0}},
      {"TS N=64  2 acc runs of 32 ", {64, 2, 32, 0, 0, 8, 0}},
      {"SS N=128 1 acc            ", {128, 1, 1, 1, 0, 8, 0}},
      {"TS N=128 1 acc + stream   ", {128, 1, 1, 0, 1, 6, 0}},
      {"TS N=128 2x32 + stream    ", {128, 2, 32, 0, 1, 6, 0}},
      {"TS N=256 1 acc + stream   ", {256, 1, 1, 0, 1, 6, 0}},
      {"TS N=128 1 acc + epi smem ", {128, 1, 1, 0, 0, 8, 1}},
      {"TS N=128 1 acc + stream+epi", {128, 1, 1, 0, 1, 6, 1}},
      {"TS N=128 1 acc, 1 B stage ", {128, 1, 1, 0, 0, 1, 0}},
      {"TS N=128 1 acc + free stream", {128, 1, 1, 0, 2, 6, 0}},
      {"tiny MMAs: N=16 + free stream", {16, 1, 1, 0, 2, 6, 0}},
      {"TS N=128 2x32 + fence per 4   ", {128, 2, 32, 0, 0, 8, 0, 1, 0, 0, 0}, 0},
      {"TS N=128 2x32 + poll+fence    ", {128, 2, 32, 0, 0, 8, 0, 1, 1, 0, 0}, 0},
      {"TS N=128 2x32 + A walks 256col", {128, 2, 32, 0, 0, 8, 0, 0, 0, 0, 1}, 0},
      {"TS N=128 2x32 + tcgen05.ld    ", {128, 2, 32, 0, 0, 8, 0, 0, 0, 1, 0}, 0},
      {"TS N=128 2x32 + ld+stream+walk+poll+fence", {128, 2, 32, 0, 1, 6, 0, 1, 1, 1, 1}, 0},
      {"N=128 free stream window 11, 148 CTAs", {128, 1, 4, 0, 2, 1, 0}, 0},
      {"N=128 free stream window 8,  148 CTAs", {128, 1, 4, 0, 2, 4, 0}, 0},
      {"N=128 free stream window 4,  148 CTAs", {128, 1, 4, 0, 2, 8, 0}, 0},
      {"N=128 free stream window 2,  148 CTAs", {128, 1, 4, 0, 2, 10, 0}, 0},
      {"N=128 free stream window 11, 74 CTAs ", {128, 1, 4, 0, 2, 1, 0}, 74},
      {"N=128 free stream window 11, 16 CTAs ", {128, 1, 4, 0, 2, 1, 0}, 16},
      {"N=128 free stream window 4,  16 CTAs ", {128, 1, 4, 0, 2, 8, 0}, 16},
      {"N=128 free stream window 11, 1 CTA   ", {128, 1, 4, 0, 2, 1, 0}, 1},
  };
  const int n_mma = 8192;
  std::vector<unsigned long long> clk(sms), str(sms);
  for (const Case& c0 : cases) {
    Case c = c0;
    if (c.grid <= 0) c.grid = sms;
    CK(cudaMemset(d_str, 0, sms * 8));
    for (int rep = 0; rep < 2; ++rep) {
      rate_kernel<<<c.grid, 192, smem>>>(wimg, n_img, c.m, n_mma, d_clk, d_str);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(clk.data(), d_clk, sms * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(str.data(), d_str, sms * 8, cudaMemcpyDeviceToHost));
    double sum = 0, mx = 0, mn = 1e30, ssum = 0;
    for (int i = 0; i < c.grid; ++i) {
      const double v = (double)clk[i] / n_mma;
      sum += v; mx = std::max(mx, v); mn = std::min(mn, v);
      ssum += (double)str[i];
    }
    printf("%s clk/MMA avg %.1f min %.1f max %.1f (ideal %d)  stream: %.2f stages per 4 MMAs\n", c.name, sum / c.grid, mn, mx, c.m.n / 2,
           ssum / c.grid / (n_mma / 4.0));
  }
  return 0;
}
