// S1b — the PointNet shared MLP + max at FP32 ACCURACY on the tensor cores: the `f32` precision of b200bev_pointnet_encode
// when the layer widths are C-64-128-256-512-1024 (configs/base.yaml:176-185); any other widths take the FFMA kernel.
//
// Reference arithmetic: PointNetLiDAREncoder.forward, src/encoders.py:289-298 (eval mode, BatchNorm folded), fp32.  Parity
// bound 1e-5 of max|ref| (north_star).  A single bf16 or TF32 product misses that by three orders of magnitude (SURVEY §7:
// 4.5e-3 / 3.5e-3); here every fp32 product a.w is THREE fp16 tensor-core products
//     a.w ~= a_hi.w_hi + a_lo.w_hi + a_hi.w_lo,      x_hi = fp16(x), x_lo = fp16(x - x_hi)          (22 mantissa bits each)
// accumulated in fp32 in tensor memory; the dropped a_lo.w_lo term is 2^-22 of the product.  Measured on the golden frame
// (tests): 7e-7 of max|ref| after five layers; three bf16 products give 1.2e-5 (not enough), bf16-hi + fp16-lo 4.3e-6.
// fp16 has a narrow exponent range, so both operands are scaled by exact powers of two:
//   weights      per output channel, 2^k with max_k |w| in [1,2) after scaling (pack time; undone in the epilogue)
//   activations  per layer, 2^k with the layer's maximum in [2^13, 2^14): the maximum of every layer's output is reduced by
//                the kernel that PRODUCES the layer (atomicMax over the whole batch), and the kernel that consumes it
//                reads the word — exact, data-driven, no calibration pass and no overflow for any input.
// That is also why the network runs LAYER BY LAYER (five launches: layer 1 on CUDA cores, K = 4; layers 2-5 as GEMMs)
// instead of one fused kernel like the bf16 path: the fp32 activations go through HBM/L2 between layers (3.8 KB per point,
// 1.4 ms of HBM time per 32 x 35,000 points against 4.7 TFLOP of fp16 tensor work), and the split costs no extra bytes —
// the consumer's producer warps read fp32 and split into hi/lo on their way into shared memory.
//
// GEMM kernel (layers 2..5), D[co][px] = sum_k W[co][k] X[px][k]:  M = 128 output channels = the TMEM lanes, N = 256 points
// = the columns (so the max over points in layer 5 is a per-thread reduction, and a per-cell run walk needs no transpose),
// both operands from shared memory, K-major, 128-byte swizzle.  Per 64-wide k block: x_hi, x_lo (2 x 32 KB, ring of 2),
// w_hi, w_lo (2 x 16 KB, ring of 3, one 32 KB bulk copy), twelve 128x256x16 MMAs.  Warp-specialised, everything meets at
// mbarriers: 16 producer warps (fp32 rows -> registers -> scale, split, 16-byte swizzled stores), 1 weight-copy warp,
// 1 MMA-issue warp, 4 epilogue warps; accumulator double-buffered in tensor memory (2 x 256 columns).
#include <cuda_fp16.h>

#include "async_copy.cuh"
#include "common.cuh"

namespace b200bev {
namespace {

constexpr int kPx = 256;                   // points per tile = MMA N
constexpr int kCo = 128;                   // output channels per tile = MMA M
constexpr int kKB = 64;                    // k per stage: 128-byte fp16 rows
constexpr int kXHalf = kPx * kKB * 2;      // 32 KB: the hi (or lo) block of a stage
constexpr int kXStage = 2 * kXHalf;
constexpr int kWHalf = kCo * kKB * 2;      // 16 KB
constexpr int kWStage = 2 * kWHalf;
constexpr int kXRing = 2, kWRing = 3;
constexpr int kProducerWarps = 16;
// warps kProducerWarps .. +3 are the epilogue: warp % 4 == the TMEM lane quadrant a warp may read
constexpr int kMmaWarp = kProducerWarps + 4, kWeightWarp = kProducerWarps + 5;
constexpr int kThreads = (kProducerWarps + 6) * 32;
constexpr int kSmem = kXRing * kXStage + kWRing * kWStage + 1024 /*alignment*/ + 256 /*barriers: 19 + the tensor-memory slot*/;
constexpr int kMaxCin = 16;
__host__ __device__ constexpr int cin_of(int l) { return 64 << l; }      // layers 2..5 = l 0..3
__host__ __device__ constexpr int cout_of(int l) { return 128 << l; }
constexpr size_t kPassPoints = 1u << 20;   // points per pass the default workspace is sized for (~3.2 GB of activations)

// ---- weight image ------------------------------------------------------------------------------------------------------
// [W1^T (C x 64) f32][b1 (64) f32] pad to 1 KB | per layer 2..5: stages [(co tile, k block)] of {w_hi 16 KB, w_lo 16 KB},
// then inv_scale (Cout) f32, bias (Cout) f32, pad to 1 KB.
struct SplitLayout {
  size_t stages[4], tail[4], total;
};
__host__ __device__ inline SplitLayout split_layout(int C) {
  SplitLayout L{};
  size_t off = (((size_t)C * 64 + 64) * sizeof(float) + 1023) & ~(size_t)1023;
  for (int l = 0; l < 4; ++l) {
    L.stages[l] = off;
    off += (size_t)(cout_of(l) / kCo) * (cin_of(l) / kKB) * kWStage;
    L.tail[l] = off;
    off = (off + 2 * (size_t)cout_of(l) * sizeof(float) + 1023) & ~(size_t)1023;
  }
  L.total = off;
  return L;
}

bool split_dims_supported(const int32_t* dims, int n_layers) {
  return dims && n_layers == 5 && dims[0] >= 1 && dims[0] <= kMaxCin && dims[1] == 64 && dims[2] == 128 && dims[3] == 256 &&
         dims[4] == 512 && dims[5] == 1024;
}

// per output channel: 2^k that brings max_k |w| into [1,2) (1 for an all-zero row); its inverse and the bias go to the tail
__global__ void __launch_bounds__(256) split_scale_kernel(const float* __restrict__ params, int C, uint8_t* __restrict__ img) {
  const SplitLayout L = split_layout(C);
  const int dims[6] = {C, 64, 128, 256, 512, 1024};
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  size_t w_off = (size_t)C * 64 + 64;
  int base = 0;
  for (int l = 0; l < 4; ++l) {
    const int K = dims[l + 1], N = dims[l + 2];
    if (warp >= base && warp < base + N) {
      const int n = warp - base;
      float m = 0.f;
      for (int k = lane; k < K; k += 32) m = fmaxf(m, fabsf(params[w_off + (size_t)k * N + n]));
#pragma unroll
      for (int d = 16; d; d >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, d));
      if (lane == 0) {
        int ex = 0;
        float inv = 1.f;
        if (m > 0.f && isfinite(m)) {
          frexpf(m, &ex);                 // m = f * 2^ex, f in [0.5, 1): 2^(1-ex) brings it into [1, 2)
          inv = ldexpf(1.f, ex - 1);
        }
        float* tail = reinterpret_cast<float*>(img + L.tail[l]);
        tail[n] = inv;
        tail[N + n] = params[w_off + (size_t)K * N + n];
      }
    }
    base += N;
    w_off += (size_t)K * N + N;
  }
}

// stage (layer, co tile, k block), row r, 16-byte chunk ch: eight weights scaled, split, swizzled as SWIZZLE_128B reads them
__global__ void __launch_bounds__(256) split_pack_kernel(const float* __restrict__ params, int C, uint8_t* __restrict__ img) {
  const SplitLayout L = split_layout(C);
  const int dims[6] = {C, 64, 128, 256, 512, 1024};
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx < (long long)C * 64 + 64) reinterpret_cast<float*>(img)[idx] = params[idx];   // W1^T, b1
  size_t w_off = (size_t)C * 64 + 64;
  for (int l = 0; l < 4; ++l) {
    const int K = dims[l + 1], N = dims[l + 2];
    const int nkb = K / kKB;
    const long long n_chunks = (long long)(N / kCo) * nkb * kCo * 8;
    if (idx < n_chunks) {
      const int ch = (int)(idx & 7), r = (int)((idx >> 3) & 127);
      const long long stage = idx >> 10;
      const int kb = (int)(stage % nkb), ct = (int)(stage / nkb);
      const int n = ct * kCo + r;
      const float sc = 1.f / reinterpret_cast<const float*>(img + L.tail[l])[n];   // exact: a power of two
      __align__(16) __half hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float w = params[w_off + (size_t)(kb * kKB + ch * 8 + j) * N + n] * sc;
        hi[j] = __float2half_rn(w);
        lo[j] = __float2half_rn(w - __half2float(hi[j]));
      }
      uint8_t* st = img + L.stages[l] + (size_t)stage * kWStage + r * 128 + ((ch ^ (r & 7)) << 4);
      *reinterpret_cast<uint4*>(st) = *reinterpret_cast<const uint4*>(hi);
      *reinterpret_cast<uint4*>(st + kWHalf) = *reinterpret_cast<const uint4*>(lo);
      return;
    }
    idx -= n_chunks;
    w_off += (size_t)K * N + N;
  }
}

// ---- layer 1 (K = C_in = 4) on CUDA cores, the gather by `perm`, the cell id of every sorted slot ---------------------------
struct L1Args {
  const float* pts;        // (B, N, C)
  const uint8_t* img;
  const int32_t* perm;     // (B, N) or null
  const int32_t* offsets;  // (B, n_cells + 1) or null
  int n_cells;
  int B0, nB, N, Npad, C;  // frames [B0, B0 + nB) of this pass
  float* act1;             // (nB * Npad, 64)
  int32_t* cid;            // (nB * Npad) or null
  uint32_t* stat;          // [0]: max of act1 (float bits)
};

__global__ void __launch_bounds__(256) split_layer1_kernel(L1Args a) {
  __shared__ float w_s[kMaxCin * 64 + 64];
  const float* w1 = reinterpret_cast<const float*>(a.img);
  for (int i = threadIdx.x; i < a.C * 64 + 64; i += 256) w_s[i] = __ldg(w1 + i);
  __syncthreads();
  const long long slot = blockIdx.x * 64ll + (threadIdx.x >> 2);
  const int part = threadIdx.x & 3;
  float vmax = 0.f;
  if (slot < (long long)a.nB * a.Npad) {
    const int fl = (int)(slot / a.Npad), s = (int)(slot - (long long)fl * a.Npad), f = a.B0 + fl;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    if (s < a.N) {
      const int p = a.perm ? __ldg(a.perm + (size_t)f * a.N + s) : s;
      const float* src = a.pts + ((size_t)f * a.N + p) * a.C;
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = w_s[a.C * 64 + part * 16 + j];
      for (int k = 0; k < a.C; ++k) {
        const float xk = __ldg(src + k);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fmaf(w_s[k * 64 + part * 16 + j], xk, acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        acc[j] = fmaxf(acc[j], 0.f);
        vmax = fmaxf(vmax, acc[j]);
      }
    }
    float4* dst = reinterpret_cast<float4*>(a.act1 + slot * 64 + part * 16);   // rows past the frame's end are zeros
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    if (a.cid && part == 0) {
      int c = -1;
      const int32_t* off = a.offsets + (size_t)f * (a.n_cells + 1);
      if (s < a.N && s < __ldg(off + a.n_cells)) {       // in-grid points come first in perm; cid = largest c with off[c] <= s
        int lo = 0, hi = a.n_cells;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(off + mid) <= s) lo = mid; else hi = mid;
        }
        c = lo;
      }
      a.cid[slot] = c;
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(FULL_MASK, vmax, d));
  if ((threadIdx.x & 31) == 0 && vmax > 0.f) atomicMax(a.stat, __float_as_uint(vmax));
}

// ---- layers 2..5 ---------------------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* x;            // (rows, Cin) fp32, rows = nB * Npad
  float* y;                  // (rows, Cout) fp32                                  [hidden layers]
  const uint8_t* stages;     // this layer's weight stages
  const float* tail;         // inv_scale (Cout), bias (Cout)
  const uint32_t* stat_in;   // max of x (float bits)
  uint32_t* stat_out;        // max of y                                           [hidden layers]
  int Cin, Cout, nB, N, Npad;
  float* out_global;         // (B, 1024) rows of this pass, or null               [last layer]
  float* out_canvas;         // (B, n_cells, 1024) rows of this pass, or null      [last layer]
  const int32_t* cid;        // cell id per sorted slot (rows)                     [last layer, cell mode]
  int n_cells;
};

__device__ __forceinline__ void tcs_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcs_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcs_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {   // K-major, 128-byte swizzle, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- CTA-pair (cta_group::2) helpers, as in pointnet_mlp_tc.cu / conv_tc.cu ----
__device__ __forceinline__ uint32_t split_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void split_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void split_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_addr(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void split_arrive_cluster_relaxed(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_addr(bar)),
      "r"(rank)
      : "memory");
}
template <int CG>
__device__ __forceinline__ void tcs_commit_cg(uint64_t* bar) {
  if constexpr (CG == 1) {
    tcs_commit(bar);
  } else {   // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_addr(bar)),
                 "h"((uint16_t)3)
                 : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void umma_f16_ss_cg(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1) {
    umma_f16_ss(d, adesc, bdesc, idesc, accumulate);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

#define SPLIT_LD32(r, taddr)                                                                                               \
  asm volatile(                                                                                                            \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"      \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                           \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),        \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),             \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),            \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                          \
      : "r"(taddr)                                                                                                         \
      : "memory")

// CG = 2: a CTA pair (thread-block cluster of two, tcgen05.mma.cta_group::2) per tile of 256 output channels x 256 points.  Each
// CTA streams its own co tile's weight stages and converts HALF of the points (128 rows of x_hi / x_lo): the fp32 activations
// cross L2 -> SM and the producers' split once for two co tiles, and the tensor core reads 8 KB of shared memory per MMA and SM
// instead of 12 (shared-memory bandwidth held the single-CTA kernel at 0.6-0.7 of the fp16 peak).  The leader
// (cluster rank 0) issues every MMA; the follower's producer warps arrive on the leader's `peer_x`, its otherwise idle MMA warp
// relays its weight stages' completions to `peer_w`; tcgen05.commit arrives in both CTAs; both CTAs' epilogue warps release the
// accumulator halves at the leader.
template <bool FINAL, int CG>
__global__ void __launch_bounds__(kThreads, 1) split_gemm_kernel(GemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* xring = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  uint8_t* wring = xring + kXRing * kXStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wring + kWRing * kWStage);
  uint64_t* full_x = bars;                         // [2] one arrival per producer warp
  uint64_t* empty_x = bars + kXRing;               // [2] tcgen05.commit
  uint64_t* full_w = bars + 2 * kXRing;            // [3] bulk-copy bytes
  uint64_t* empty_w = full_w + kWRing;             // [3] tcgen05.commit
  uint64_t* acc_full = empty_w + kWRing;           // [2]
  uint64_t* acc_empty = acc_full + 2;              // [2] one arrival per epilogue warp of the pair, at the leader
  uint64_t* peer_x = acc_empty + 2;                // [2] CG = 2, leader: the follower's half of the x stage is in place
  uint64_t* peer_w = peer_x + kXRing;              // [3] CG = 2, leader: the follower's weight stage has landed (relayed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_w + kWRing);
  const uint32_t rank = CG == 2 ? split_cluster_ctarank() : 0u;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == kMmaWarp) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  if (tid == 0) {
    for (int s = 0; s < kXRing; ++s) {
      mbarrier_init(&full_x[s], kProducerWarps);
      mbarrier_init(&empty_x[s], 1);
      mbarrier_init(&peer_x[s], kProducerWarps);
    }
    for (int s = 0; s < kWRing; ++s) {
      mbarrier_init(&full_w[s], 1);
      mbarrier_init(&empty_w[s], 1);
      mbarrier_init(&peer_w[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbarrier_init(&acc_full[s], 1);
      mbarrier_init(&acc_empty[s], 4 * CG);
    }
    mbarrier_init_fence();
  }
  tcs_fence_before();
  __syncthreads();
  if constexpr (CG == 2) split_cluster_sync();   // the peer's barriers are initialised before anyone signals them
  tcs_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int n_co = a.Cout / kCo, nkb = a.Cin / kKB;
  const int tiles_per_frame = a.Npad / kPx;
  // a tile is CG co tiles x 256 points, walked by a cluster of CG CTAs in lock step (host: n_co % CG == 0); the co groups of one
  // point tile are neighbours in the grid (shared x in L2)
  const int n_cg = n_co / CG;
  const int n_tiles = a.nB * tiles_per_frame * n_cg;
  const int first_tile = (int)blockIdx.x / CG, tile_step = (int)gridDim.x / CG;
  const int my_tiles = (n_tiles - first_tile + tile_step - 1) / tile_step;
  const int total = my_tiles * nkb;
  constexpr int kRowsJ = 4 / CG;             // this CTA's rows of an x stage: 256 / CG points, 64 per step of j

  if (warp < kProducerWarps) {
    // ---- producers: fp32 rows -> scale by the layer's power of two -> fp16 hi / lo -> swizzled shared memory ----
    const float S = activation_scale(__ldg(a.stat_in));
    const int row0 = tid >> 3, chunk = tid & 7;                 // rows row0 + 64 j, j = 0..3; 8 floats of the 64-wide k block
    const uint32_t dst_off = (uint32_t)(row0 * 128 + ((chunk ^ (row0 & 7)) << 4));
    float4 raw[2 * kRowsJ];
    int tile = first_tile, kb = 0;
    auto load = [&](int t, int k) {
      const long long px0 = (long long)(t / n_cg) * kPx + (long long)rank * (kPx / CG);
      const float* src = a.x + (px0 + row0) * a.Cin + k * kKB + chunk * 8;
#pragma unroll
      for (int j = 0; j < kRowsJ; ++j) {
        raw[2 * j] = __ldg(reinterpret_cast<const float4*>(src + (size_t)j * 64 * a.Cin));
        raw[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(src + (size_t)j * 64 * a.Cin + 4));
      }
    };
    if (total > 0) load(tile, 0);
    for (int it = 0; it < total; ++it) {
      uint4 hi[kRowsJ], lo[kRowsJ];
#pragma unroll
      for (int j = 0; j < kRowsJ; ++j) {
        const float v[8] = {raw[2 * j].x * S,     raw[2 * j].y * S,     raw[2 * j].z * S,     raw[2 * j].w * S,
                            raw[2 * j + 1].x * S, raw[2 * j + 1].y * S, raw[2 * j + 1].z * S, raw[2 * j + 1].w * S};
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __half2 hh = __floats2half2_rn(v[2 * e], v[2 * e + 1]);          // element with the even k in the low half
          const float2 back = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(v[2 * e] - back.x, v[2 * e + 1] - back.y);
          h[e] = *reinterpret_cast<const uint32_t*>(&hh);
          l[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        hi[j] = make_uint4(h[0], h[1], h[2], h[3]);
        lo[j] = make_uint4(l[0], l[1], l[2], l[3]);
      }
      // the next stage's rows are in flight while this one waits for its slot and is stored
      if (++kb == nkb) {
        kb = 0;
        tile += tile_step;
      }
      if (it + 1 < total) load(tile, kb);
      const uint32_t slot = (uint32_t)it % kXRing;
      if (it >= kXRing) mbarrier_wait(&empty_x[slot], (((uint32_t)it / kXRing) - 1) & 1);
      uint8_t* dst = xring + slot * kXStage + dst_off;
#pragma unroll
      for (int j = 0; j < kRowsJ; ++j) {
        *reinterpret_cast<uint4*>(dst + j * 64 * 128) = hi[j];
        *reinterpret_cast<uint4*>(dst + j * 64 * 128 + kXHalf) = lo[j];
      }
      fence_proxy_async_shared();        // generic-proxy stores ordered before the tensor core's asynchronous-proxy reads
      __syncwarp();
      if (lane == 0) {
        if (CG == 1 || rank == 0) mbarrier_arrive(&full_x[slot]);
        else split_arrive_cluster(&peer_x[slot], 0);       // the leader issues the MMAs that read these rows
      }
    }
  } else if (warp == kWeightWarp) {
    // ---- weight stages, (tile, k block) in the order the MMAs use them: one 32 KB bulk copy each ----
    int tile = first_tile, kb = 0;
    for (int g = 0; g < total; ++g) {
      const uint32_t slot = (uint32_t)g % kWRing;
      if (g >= kWRing) mbarrier_wait(&empty_w[slot], (((uint32_t)g / kWRing) - 1) & 1);
      if (elect_one()) {
        mbarrier_expect_tx(&full_w[slot], kWStage);
        bulk_copy_global_to_shared(wring + slot * kWStage, a.stages + ((size_t)((tile % n_cg) * CG + (int)rank) * nkb + kb) * kWStage, kWStage,
                                   &full_w[slot]);
      }
      __syncwarp();
      if (++kb == nkb) {
        kb = 0;
        tile += tile_step;
      }
    }
  } else if (warp == kMmaWarp && CG == 2 && rank != 0) {
    // ---- follower of a pair: forward the completion of every weight stage to the leader ----
    for (int c = 0; c < total; ++c) {
      const uint32_t ws = (uint32_t)c % kWRing;
      mbarrier_wait(&full_w[ws], ((uint32_t)c / kWRing) & 1);
      if (lane == 0) split_arrive_cluster_relaxed(&peer_w[ws], 0);
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {
    // ---- MMA issuer: 12 x (128 x 256 x 16) per k block: hi.hi, w_hi.x_lo, w_lo.x_hi ----
    // instruction descriptor: D f32 (bit 4), A and B fp16 (format 0), both K-major, N = 256, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kPx >> 3) << 17) | ((uint32_t)((kCo * CG) >> 4) << 24);
    int kb = 0, tile_seq = 0;
    for (int c = 0; c < total; ++c) {
      const uint32_t xs = (uint32_t)c % kXRing, ws = (uint32_t)c % kWRing, buf = tile_seq & 1;
      if (kb == 0 && tile_seq >= 2) mbarrier_wait(&acc_empty[buf], ((tile_seq >> 1) - 1) & 1);
      mbarrier_wait(&full_x[xs], ((uint32_t)c / kXRing) & 1);
      mbarrier_wait(&full_w[ws], ((uint32_t)c / kWRing) & 1);
      if constexpr (CG == 2) {
        mbarrier_wait(&peer_x[xs], ((uint32_t)c / kXRing) & 1);
        mbarrier_wait(&peer_w[ws], ((uint32_t)c / kWRing) & 1);
      }
      tcs_fence_after();
      if (elect_one()) {
        const uint32_t x_hi = smem_addr(xring + xs * kXStage), x_lo = x_hi + kXHalf;
        const uint32_t w_hi = smem_addr(wring + ws * kWStage), w_lo = w_hi + kWHalf;
        const uint32_t d = tmem + buf * kPx;
#pragma unroll
        for (int s = 0; s < 4; ++s) umma_f16_ss_cg<CG>(d, sw128_desc(w_hi + s * 32), sw128_desc(x_hi + s * 32), idesc, !(kb == 0 && s == 0));
#pragma unroll
        for (int s = 0; s < 4; ++s) umma_f16_ss_cg<CG>(d, sw128_desc(w_hi + s * 32), sw128_desc(x_lo + s * 32), idesc, 1u);
#pragma unroll
        for (int s = 0; s < 4; ++s) umma_f16_ss_cg<CG>(d, sw128_desc(w_lo + s * 32), sw128_desc(x_hi + s * 32), idesc, 1u);
        tcs_commit_cg<CG>(&empty_x[xs]);
        tcs_commit_cg<CG>(&empty_w[ws]);
        if (kb == nkb - 1) tcs_commit_cg<CG>(&acc_full[buf]);
      }
      __syncwarp();
      if (++kb == nkb) {
        kb = 0;
        ++tile_seq;
      }
    }
  } else {
    // ---- epilogue: warp kProducerWarps + q owns TMEM lanes [32q, 32q + 32) = 32 output channels, a lane is one channel ----
    const int quad = warp & 3;
    const float inv_S = 1.f / activation_scale(__ldg(a.stat_in));   // exact: a power of two
    float layer_max = 0.f;
    int tile = first_tile;
    for (int tile_seq = 0; tile_seq < my_tiles; ++tile_seq, tile += tile_step) {
      const int co_tile = (tile % n_cg) * CG + (int)rank, px_tile = tile / n_cg;
      const int fl = px_tile / tiles_per_frame, s0 = (px_tile - fl * tiles_per_frame) * kPx;
      const long long row_base = (long long)px_tile * kPx;          // = fl * Npad + s0
      const uint32_t buf = tile_seq & 1;
      const int co = co_tile * kCo + quad * 32 + lane;
      const float unscale = __ldg(a.tail + co) * inv_S, bias = __ldg(a.tail + a.Cout + co);
      int c_cur = -1;
      if (FINAL && a.out_canvas != nullptr) c_cur = __ldg(a.cid + row_base + lane);
      mbarrier_wait(&acc_full[buf], (tile_seq >> 1) & 1);
      tcs_fence_after();
      const uint32_t tbase = tmem + ((uint32_t)(quad * 32) << 16) + buf * kPx;
      if (!FINAL) {
#pragma unroll 1
        for (int col0 = 0; col0 < kPx; col0 += 32) {
          uint32_t r[32];
          SPLIT_LD32(r, tbase + (uint32_t)col0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float* dst = a.y + (row_base + col0) * a.Cout + co;       // a store instruction: 32 channels of one point, 128 B
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = fmaxf(fmaf(__uint_as_float(r[j]), unscale, bias), 0.f);
            layer_max = fmaxf(layer_max, v);
            dst[(size_t)j * a.Cout] = v;
          }
        }
      } else {
        // max over the tile's points; per-cell runs (slots are in cell order) go to the canvas
        const int n_valid = a.N - s0 < kPx ? a.N - s0 : kPx;        // slots past the frame's end are padding, not points
        float gmax = -INFINITY, m = -INFINITY;
        bool first_run = true;
        float* canvas = a.out_canvas ? a.out_canvas + (size_t)fl * a.n_cells * a.Cout + co : nullptr;
#pragma unroll 1
        for (int col0 = 0; col0 < kPx; col0 += 32) {
          uint32_t r[32];
          SPLIT_LD32(r, tbase + (uint32_t)col0);
          // the cell ids of the NEXT 32 slots travel while this group is walked (the first group's were fetched before the
          // wait for the accumulator): a dependent global load per group was a third of the cell-mode epilogue
          int c_nxt = INT_MIN;
          if (canvas != nullptr && col0 + 32 < kPx) c_nxt = __ldg(a.cid + row_base + col0 + 32 + lane);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (col0 < n_valid) {                                       // warp-uniform
            if (col0 + 32 > n_valid) {                                // the frame's last, ragged group: padding never counts
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j >= n_valid) r[j] = 0xff800000u;          // -inf
            }
            if (canvas == nullptr) {
#pragma unroll
              for (int j = 0; j < 32; ++j) gmax = fmaxf(gmax, __uint_as_float(r[j]));
            } else {
              int c_n = __shfl_down_sync(FULL_MASK, c_cur, 1);
              const int c_first_next = __shfl_sync(FULL_MASK, c_nxt, 0);
              if (lane == 31) c_n = col0 + 32 < kPx ? c_first_next : INT_MIN;   // the tile's last slot ends a run
              // the same word in every lane: broadcast through a shuffle so that the branches on its bits are provably warp-uniform
              const unsigned ends = __shfl_sync(FULL_MASK, __ballot_sync(FULL_MASK, c_cur != c_n), 0);
#pragma unroll
              for (int j4 = 0; j4 < 32; j4 += 4) {
                if (((ends >> j4) & 0xfu) == 0u) {                    // no run ends in these four slots: most groups
                  m = fmaxf(fmaxf(m, fmaxf(__uint_as_float(r[j4]), __uint_as_float(r[j4 + 1]))),
                            fmaxf(__uint_as_float(r[j4 + 2]), __uint_as_float(r[j4 + 3])));
                  continue;
                }
#pragma unroll
                for (int j = j4; j < j4 + 4; ++j) {
                  m = fmaxf(m, __uint_as_float(r[j]));
                  if (ends & (1u << j)) {
                    const int cj = __shfl_sync(FULL_MASK, c_cur, j);
                    if (cj >= 0) {
                      // zeros are stored like any value (the canvas starts at zero): a store that depends on the value would be a
                      // divergent branch per run end
                      const float val = fmaxf(fmaf(m, unscale, bias), 0.f);
                      float* dst = canvas + (size_t)cj * a.Cout;      // 32 lanes: 128 contiguous bytes of the cell's row
                      // a run touching the tile's first or last slot may go on in a neighbouring tile: atomic; else the only writer
                      if (first_run || col0 + j == kPx - 1) atomicMax(reinterpret_cast<int*>(dst), __float_as_int(val));
                      else *dst = val;
                    }
                    gmax = fmaxf(gmax, m);                            // the global maximum is the maximum of the runs' maxima
                    m = -INFINITY;
                    first_run = false;
                  }
                }
              }
            }
          }
          c_cur = c_nxt;
        }
        gmax = fmaxf(gmax, m);
        if (a.out_global) {
          const float val = fmaxf(fmaf(gmax, unscale, bias), 0.f);    // bias + ReLU commute with the max (unscale > 0)
          if (val > 0.f) atomicMax(reinterpret_cast<int*>(a.out_global + (size_t)fl * a.Cout + co), __float_as_int(val));
        }
      }
      tcs_fence_before();    // the tensor-memory loads above are complete (wait::ld) before the half is handed back
      __syncwarp();
      if (lane == 0) {
        if (CG == 1 || rank == 0) mbarrier_arrive(&acc_empty[buf]);
        else split_arrive_cluster(&acc_empty[buf], 0);
      }
    }
    if (!FINAL) {
#pragma unroll
      for (int d = 16; d; d >>= 1) layer_max = fmaxf(layer_max, __shfl_xor_sync(FULL_MASK, layer_max, d));
      if (lane == 0 && layer_max > 0.f) atomicMax(a.stat_out, __float_as_uint(layer_max));
    }
  }

  tcs_fence_before();
  __syncthreads();
  if constexpr (CG == 2) split_cluster_sync();   // no CTA leaves (or frees tensor memory) while its peer may still signal or compute
  if (warp == kMmaWarp) {
    tcs_fence_after();
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// workspace: [stats: 16 words][cid: rows i32][bufA: rows x 256 f32 (act1, act3)][bufB: rows x 512 f32 (act2, act4)]
struct SplitWorkspace {
  size_t cid, buf_a, buf_b, total;
};
inline SplitWorkspace split_workspace(size_t rows) {
  SplitWorkspace w{};
  w.cid = 256;
  w.buf_a = (w.cid + rows * sizeof(int32_t) + 1023) & ~(size_t)1023;
  w.buf_b = w.buf_a + rows * 256 * sizeof(float);
  w.total = w.buf_b + rows * 512 * sizeof(float);
  return w;
}
inline int padded_points(int N) { return ceil_div(N, kPx) * kPx; }

}  // namespace

int pointnet_encode_split(const float* points, int B, int N, int C, const int32_t* dims, int n_layers, const int32_t* perm,
                          const int32_t* offsets, int n_cells, const void* image, float* out_global, float* out_canvas,
                          void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (!split_dims_supported(dims, n_layers) || dims[0] != C) return B200BEV_ERR_UNSUPPORTED;
  if (!image || !workspace || ((reinterpret_cast<uintptr_t>(image) | reinterpret_cast<uintptr_t>(workspace)) & 255) != 0)
    return B200BEV_ERR_INVALID_ARGUMENT;
  const int Npad = padded_points(N);
  // frames per pass: as many as the workspace holds
  int per_pass = 0;
  for (int f = B; f >= 1; f = (f + 1) / 2 == f ? f - 1 : (f + 1) / 2) {
    if (split_workspace((size_t)f * Npad).total <= workspace_bytes) { per_pass = f; break; }
    if (f == 1) break;
  }
  if (per_pass == 0) return B200BEV_ERR_WORKSPACE;
  const bool cell = out_canvas != nullptr;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(image);
  const SplitLayout L = split_layout(C);
  if (out_global) B200BEV_CUDA_TRY(cudaMemsetAsync(out_global, 0, (size_t)B * 1024 * sizeof(float), st));
  if (out_canvas) B200BEV_CUDA_TRY(cudaMemsetAsync(out_canvas, 0, (size_t)B * n_cells * 1024 * sizeof(float), st));
  B200BEV_CUDA_TRY(cudaFuncSetAttribute(split_gemm_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  B200BEV_CUDA_TRY(cudaFuncSetAttribute(split_gemm_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  B200BEV_CUDA_TRY(cudaFuncSetAttribute(split_gemm_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  B200BEV_CUDA_TRY(cudaFuncSetAttribute(split_gemm_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  // layers with at least two co tiles run as CTA pairs (cta_group::2)
  auto launch_gemm = [&](const GemmArgs& g, bool final_layer) -> int {
    const long long tiles = (long long)g.nB * (g.Npad / kPx) * (g.Cout / kCo);
    const bool pair = (g.Cout / kCo) % 2 == 0 && !debug_env("B200BEV_SPLIT_SINGLE");
    if (!pair) {
      const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
      if (final_layer) split_gemm_kernel<true, 1><<<grid, kThreads, kSmem, st>>>(g);
      else split_gemm_kernel<false, 1><<<grid, kThreads, kSmem, st>>>(g);
      return launch_status();
    }
    const long long pair_tiles = tiles / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * (pair_tiles < sm_count() / 2 ? pair_tiles : sm_count() / 2)));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (final_layer) B200BEV_CUDA_TRY(cudaLaunchKernelEx(&cfg, split_gemm_kernel<true, 2>, g));
    else B200BEV_CUDA_TRY(cudaLaunchKernelEx(&cfg, split_gemm_kernel<false, 2>, g));
    return launch_status();
  };
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  for (int b0 = 0; b0 < B; b0 += per_pass) {
    const int nB = B - b0 < per_pass ? B - b0 : per_pass;
    const size_t rows = (size_t)nB * Npad;
    const SplitWorkspace W = split_workspace((size_t)per_pass * Npad);
    uint32_t* stat = reinterpret_cast<uint32_t*>(ws);
    int32_t* cid = reinterpret_cast<int32_t*>(ws + W.cid);
    float* buf_a = reinterpret_cast<float*>(ws + W.buf_a);
    float* buf_b = reinterpret_cast<float*>(ws + W.buf_b);
    B200BEV_CUDA_TRY(cudaMemsetAsync(stat, 0, 64, st));
    L1Args l1{points, img, perm, offsets, n_cells, b0, nB, N, Npad, C, buf_a, cell ? cid : nullptr, stat};
    split_layer1_kernel<<<(unsigned)((rows + 63) / 64), 256, 0, st>>>(l1);
    B200BEV_CUDA_TRY(cudaGetLastError());
    const float* x = buf_a;
    for (int l = 0; l < 4; ++l) {
      float* y = (l & 1) ? buf_a : buf_b;      // act2 -> B, act3 -> A, act4 -> B
      GemmArgs g{};
      g.x = x; g.y = l < 3 ? y : nullptr;
      g.stages = img + L.stages[l];
      g.tail = reinterpret_cast<const float*>(img + L.tail[l]);
      g.stat_in = stat + l; g.stat_out = l < 3 ? stat + l + 1 : nullptr;
      g.Cin = cin_of(l); g.Cout = cout_of(l); g.nB = nB; g.N = N; g.Npad = Npad;
      if (l == 3) {
        g.out_global = out_global ? out_global + (size_t)b0 * 1024 : nullptr;
        g.out_canvas = out_canvas ? out_canvas + (size_t)b0 * n_cells * 1024 : nullptr;
        g.cid = cell ? cid : nullptr;
        g.n_cells = n_cells;
      }
      const int rc = launch_gemm(g, l == 3);
      if (rc) return rc;
      x = y;
    }
  }
  return launch_status();
}

}  // namespace b200bev

using namespace b200bev;

extern "C" B200BEV_API size_t b200bev_pointnet_pack_split_bytes(const int32_t* dims, int n_layers) {
  if (!split_dims_supported(dims, n_layers)) return 0;
  return split_layout(dims[0]).total;
}

extern "C" B200BEV_API int b200bev_pointnet_pack_split(const float* params, const int32_t* dims, int n_layers, void* image,
                                                       size_t image_bytes, void* stream) {
  if (!params || !image) return B200BEV_ERR_INVALID_ARGUMENT;
  if (!split_dims_supported(dims, n_layers)) return B200BEV_ERR_UNSUPPORTED;
  const SplitLayout L = split_layout(dims[0]);
  if (image_bytes < L.total || (reinterpret_cast<uintptr_t>(image) & 255) != 0) return B200BEV_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int channels = 128 + 256 + 512 + 1024;
  split_scale_kernel<<<(channels * 32 + 255) / 256, 256, 0, st>>>(params, dims[0], reinterpret_cast<uint8_t*>(image));
  B200BEV_CUDA_TRY(cudaGetLastError());
  long long chunks = 0;
  for (int l = 0; l < 4; ++l) chunks += (long long)(cout_of(l) / kCo) * (cin_of(l) / kKB) * kCo * 8;
  split_pack_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(params, dims[0], reinterpret_cast<uint8_t*>(image));
  return launch_status();
}

extern "C" B200BEV_API size_t b200bev_pointnet_split_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  const size_t npad = (size_t)padded_points(N);
  size_t frames = kPassPoints / npad;
  if (frames < 1) frames = 1;
  if (frames > (size_t)B) frames = (size_t)B;
  return split_workspace(frames * npad).total;
}
