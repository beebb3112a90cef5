// Stand-alone probe (test infrastructure) for the cell-mode epilogue of csrc/pointnet_mlp_tc.cu: which (lane, column) of
// tensor memory does register j of thread t receive under tcgen05.ld.16x256b, and does a round trip
//   ld.16x256b -> st.32x32b (in place) -> ld.16x256b -> one shuffle level
// leave thread = column, register = lane (a 32 x 32 transposition that never touches shared memory)?
// Part 1 dumps the raw fragment layout; part 2 runs the transposition as the kernel would and checks it; part 3 times it
// against the shared-memory transposition (private padded tile, 32 STS + 8 LDS.128) in clocks per 32 x 32 block.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_shapes_probe tmem_shapes_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD_32x32b_X32(r, taddr)                                                                                            \
  asm volatile(                                                                                                            \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"      \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                           \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),        \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),             \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),            \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                          \
      : "r"(taddr)                                                                                                         \
      : "memory")

#define ST_32x32b_X32(taddr, r)                                                                                            \
  asm volatile(                                                                                                            \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"   \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),                                                \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),        \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),          \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),          \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                                       \
      : "memory")

// 16 lanes x 256 bit, four repeats: 32 columns of 16 lanes -> 16 registers
#define LD_16x256b_X4(r, o, taddr)                                                                                         \
  asm volatile(                                                                                                            \
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"            \
      : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]),    \
        "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]),                 \
        "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15])                                                                  \
      : "r"(taddr)                                                                                                         \
      : "memory")

__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// The transposition under test.  In: tensor memory [32 lanes of this warp's quadrant][32 columns at `taddr`].
// Out: thread t holds column chan_of(t), out[s] = the value of lane row_of_slot(s).
// Derivation (hypothesis checked by part 1): ld.16x256b gives thread (r2 r1 r0 c2 c1) the registers (half = r4, +8 = r3,
// repeat = c4 c3, element = c0).
__device__ __forceinline__ int chan_of(int t) {   // thread bits (t4 t3 t2 t1 t0) = (c0, c2, c1, c4, c3)
  return ((t >> 4) & 1) | (((t >> 2) & 3) << 1) | ((t & 3) << 3);
}

__device__ __forceinline__ void tmem_transpose(uint32_t taddr, uint32_t* out, int lane) {
  uint32_t a[32];
  LD_16x256b_X4(a, 0, taddr);                    // lanes 0..15
  LD_16x256b_X4(a, 16, taddr + (16u << 16));     // lanes 16..31
  wait_ld();
  // a[h*16 + x*4 + g*2 + e]: lane = h*16 + g*8 + (t >> 2), column = x*8 + (t & 3)*2 + e
  // store so that column' = (b4 b3 b2 b1 b0) with (b2 b1) = (c4 c3) = x and (b4 b3 b0) = (h g e):
  uint32_t s[32];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int e = 0; e < 2; ++e) s[h * 16 + g * 8 + x * 2 + e] = a[h * 16 + x * 4 + g * 2 + e];
  ST_32x32b_X32(taddr, s);
  wait_st();
  uint32_t b[32];
  LD_16x256b_X4(b, 0, taddr);
  LD_16x256b_X4(b, 16, taddr + (16u << 16));
  wait_ld();
  // now: row' = old thread = (r2 r1 r0 c2 c1), column' = (h g x1 x0 e) = (r4 r3 c4 c3 c0)
  // b[H*16 + X*4 + G*2 + E]: row' = H*16 + G*8 + (t >> 2), column' = X*8 + (t & 3)*2 + E
  //   -> r2 = H, r1 = G, (r0 c2 c1) = t >> 2, (r4 r3) = X, (c4 c3) = t & 3, c0 = E
  // thread = (r0 c2 c1 c4 c3); exchange r0 with c0 across lane ^ 16
  const bool up = (lane & 16) != 0;
#pragma unroll
  for (int H = 0; H < 2; ++H)
#pragma unroll
    for (int G = 0; G < 2; ++G)
#pragma unroll
      for (int X = 0; X < 4; ++X) {
        const uint32_t v0 = b[H * 16 + X * 4 + G * 2 + 0], v1 = b[H * 16 + X * 4 + G * 2 + 1];
        const uint32_t send = up ? v0 : v1;
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 16);
        const int row_hi = X * 8 + H * 4 + G * 2;           // (r4 r3 r2 r1 0)
        out[row_hi + 0] = up ? recv : v0;                   // r0 = 0
        out[row_hi + 1] = up ? v1 : recv;                   // r0 = 1
      }
}

__global__ void __launch_bounds__(128, 1) probe(uint32_t* raw, uint32_t* tr, unsigned long long* clk, int iters) {
  __shared__ uint32_t slot;
  __shared__ __align__(16) float tile[4][32 * 36];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot + ((uint32_t)(warp * 32) << 16);
  uint32_t v[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) v[c] = ((uint32_t)(warp * 32 + lane) << 8) | (uint32_t)c;   // (lane of the CTA, column)
  ST_32x32b_X32(tm, v);
  wait_st();
  // part 1: raw layout of ld.16x256b.x4 (both halves)
  uint32_t a[32];
  LD_16x256b_X4(a, 0, tm);
  LD_16x256b_X4(a, 16, tm + (16u << 16));
  wait_ld();
#pragma unroll
  for (int j = 0; j < 32; ++j) raw[(threadIdx.x) * 32 + j] = a[j];
  // part 2: the transposition
  uint32_t o[32];
  tmem_transpose(tm, o, lane);
#pragma unroll
  for (int j = 0; j < 32; ++j) tr[(threadIdx.x) * 32 + j] = o[j];
  __syncthreads();
  // part 3: timing, all four warps at once.  (a) tensor-memory round trip, (b) shared-memory tile
  ST_32x32b_X32(tm, v);
  wait_st();
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    tmem_transpose(tm, o, lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += o[j];
  }
  long long t1 = clock64();
  float* tp = tile[warp];
  for (int i = 0; i < iters; ++i) {
    uint32_t r[32];
    LD_32x32b_X32(r, tm);
    wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) tp[j * 36 + lane] = __uint_as_float(r[j]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 f = *reinterpret_cast<const float4*>(tp + lane * 36 + q * 4);
      acc += __float_as_uint(f.x) + __float_as_uint(f.y) + __float_as_uint(f.z) + __float_as_uint(f.w);
    }
    __syncwarp();
  }
  long long t2 = clock64();
  if (lane == 0) {
    clk[warp * 2] = (unsigned long long)(t1 - t0);
    clk[warp * 2 + 1] = (unsigned long long)(t2 - t1);
  }
  if (acc == 0x12345678u) raw[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(slot) : "memory");
}

int main() {
  uint32_t *d_raw, *d_tr;
  unsigned long long* d_clk;
  const int iters = 2000;
  CK(cudaMalloc(&d_raw, 128 * 32 * 4));
  CK(cudaMalloc(&d_tr, 128 * 32 * 4));
  CK(cudaMalloc(&d_clk, 8 * 8));
  probe<<<1, 128>>>(d_raw, d_tr, d_clk, iters);
  CK(cudaDeviceSynchronize());
  std::vector<uint32_t> raw(128 * 32), tr(128 * 32);
  unsigned long long clk[8];
  CK(cudaMemcpy(raw.data(), d_raw, raw.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(tr.data(), d_tr, tr.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(clk, d_clk, sizeof(clk), cudaMemcpyDeviceToHost));
  // part 1: print warp 1's layout (so that a quadrant offset shows) for a few threads, and check the hypothesis for all
  int bad1 = 0;
  for (int t = 0; t < 128; ++t) {
    const int w = t / 32, l = t % 32;
    for (int j = 0; j < 32; ++j) {
      const int h = j / 16, x = (j % 16) / 4, g = (j % 4) / 2, e = j % 2;
      const int row = w * 32 + h * 16 + g * 8 + (l >> 2), col = x * 8 + (l & 3) * 2 + e;
      if (raw[t * 32 + j] != (((uint32_t)row << 8) | (uint32_t)col)) ++bad1;
    }
  }
  printf("part 1: ld.16x256b.x4 layout hypothesis: %d mismatches of %d\n", bad1, 128 * 32);
  for (int t : {32, 33, 36, 63}) {
    printf("  thread %3d:", t);
    for (int j = 0; j < 32; ++j) printf(" (%u,%u)", raw[t * 32 + j] >> 8, raw[t * 32 + j] & 255);
    printf("\n");
  }
  int bad2 = 0;
  for (int t = 0; t < 128; ++t) {
    const int w = t / 32, l = t % 32;
    const int ch = ((l >> 4) & 1) | (((l >> 2) & 3) << 1) | ((l & 3) << 3);
    for (int s = 0; s < 32; ++s)
      if (tr[t * 32 + s] != (((uint32_t)(w * 32 + s) << 8) | (uint32_t)ch)) ++bad2;
  }
  printf("part 2: transposition through tensor memory: %d mismatches of %d\n", bad2, 128 * 32);
  if (bad2) {
    for (int t : {0, 1, 4, 16, 17}) {
      printf("  thread %3d:", t);
      for (int j = 0; j < 32; ++j) printf(" (%u,%u)", tr[t * 32 + j] >> 8, tr[t * 32 + j] & 255);
      printf("\n");
    }
  }
  for (int w = 0; w < 4; ++w)
    printf("part 3: warp %d: tensor-memory round trip %.1f clk per 32x32 block, shared-memory tile %.1f clk\n", w,
           (double)clk[w * 2] / iters, (double)clk[w * 2 + 1] / iters);
  printf(bad1 == 0 && bad2 == 0 ? "TMEM-SHAPES-PROBE-OK\n" : "TMEM-SHAPES-PROBE-MISMATCH\n");
  return 0;
}
