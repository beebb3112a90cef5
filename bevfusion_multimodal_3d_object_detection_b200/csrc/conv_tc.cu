// N1 — the convolutions between the hot-path kernels, as one implicit-GEMM tcgen05 kernel.
//
// Replaces, in eval mode and when the bf16 path is enabled (parity 1e-2): every Conv2d(k=3,pad=1 | k=1) + BatchNorm2d +
// ReLU block of FlexibleBEVFusion — camera_proj src/fusion.py:126-133, lidar_upsample :151-166, radar_refine :176-183,
// bev_fusion :199-207 — and the CenterNetHead convolutions src/fusion.py:822-854 (the five 3x3 convs as one 256->320
// conv, the five 1x1 convs as one block-diagonal 320->19 conv).  BatchNorm is folded into weights and bias on the host.
//
// GEMM view: D[co][px] = sum over (tap, ci) W[co][tap][ci] * X[px + tap][ci].  The OUTPUT CHANNELS are the M = 128 rows of
// the MMA (the 128 lanes of tensor memory) and the PIXELS its N = 256 columns, so an epilogue thread owns one output
// channel and 32 consecutive pixels at a time — the NCHW fp32 layout the reference's next module expects is written
// directly, 128 contiguous bytes per thread, bias + ReLU on the way out.  Both operands come from shared memory in the
// K-major 128-byte-swizzled layout tests/cuda/umma_probe.cu pinned down:
//   A = weights: pre-packed once per weight update (b200bev_conv_pack_bf16) into 16 KB stages [128 co][64 ci], ordered as
//       the kernel consumes them ((co tile, tap, ci chunk)); one cp.async.bulk per stage.
//   B = pixels:  the input lives channels-last in bf16 (b200bev_nchw_to_nhwc_bf16 writes it, straight into the channel
//       slice of a concatenated input, so torch.cat disappears).  For a tap (dy,dx) row n of the stage is pixel
//       (y+dy, x+dx) of output pixel n — 128 contiguous bytes of global memory, or zeros outside the image — moved with
//       16-byte cp.async (zero-fill form) by all 256 threads straight to their swizzled place: im2col never exists in
//       memory.
// One persistent CTA per SM walks (pixel tile, co tile) pairs; warp-specialised (producers / MMA issuer / epilogue meet at
// mbarriers only), stages released by tcgen05.commit, fp32 accumulation double-buffered in TMEM (2 x 256 columns).  Two
// kernels: the 3x3 form in which the nine taps share the pixel rows in shared memory, and the per-tap form (1x1, wide images).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include <cuda.h>   // CUtensorMap and its enums only: the encoder is fetched through cudaGetDriverEntryPoint (no link to libcuda)

#include "async_copy.cuh"
#include "common.cuh"

namespace b200bev {
namespace {

constexpr int kTileCo = 128;   // M
constexpr int kTilePx = 256;   // N
constexpr int kKC = 64;        // bf16 k per stage (128-byte rows)
constexpr int kStageA = kTileCo * kKC * 2;   // 16 KB
constexpr int kStageB = kTilePx * kKC * 2;   // 32 KB
constexpr int kStage = kStageA + kStageB;
constexpr int kRing = 4;
constexpr int kEpiTile = 4 * 32 * 33 * 4;   // per epilogue warp: a 32 x 32 transposing tile
constexpr int kConvSmem = kRing * kStage + kEpiTile + 1024 /*alignment slack*/ + 256 /*barriers*/;

struct ConvArgs {
  const __nv_bfloat16* x;     // (B, H, W, Cin) channels-last
  const uint8_t* wimg;        // packed stages
  const float* bias;          // (Cout) or null
  float* out;                 // (B, Cout, H, W) fp32, or null
  int B, H, W, Cin, Cout, taps, relu;
  __nv_bfloat16* out_nhwc;    // (B, H, W, out_ct) bf16 channels-last, channels [out_coff, out_coff + Cout), or null:
  int out_ct, out_coff;       // the input layout of the next convolution, written without a separate layout pass
  // fp32-accuracy mode (three fp16 products per fp32 product, see pointnet_mlp_split.cu): x is (B,H,W,[hi Cin | lo Cin]) fp16,
  // scaled by activation_scale(*x_stat); the weight stages come in (hi, hi, lo) triples per 64-channel chunk, scaled per output
  // channel by a power of two whose inverse is unscale[co]
  int split;                  // 0: bf16 operands; 1: split fp16 operands
  int x_pitch;                // elements per pixel of x: Cin, or 2 Cin in split mode
  const float* unscale;       // (Cout) or null
  const uint32_t* x_stat;     // float bits of max|x| or null
};

// k chunks of a (tap): Cin/64, or three times that in split mode — terms (w_hi, x_hi), (w_hi, x_lo), (w_lo, x_hi)
__device__ __forceinline__ int conv_chunks(const ConvArgs& a) { return (a.Cin / kKC) * (a.split ? 3 : 1); }
// first input channel (element offset inside a pixel of x) of k chunk j
__device__ __forceinline__ int conv_chunk_channel(const ConvArgs& a, int j) {
  const int ncc = a.Cin / kKC;
  return a.split ? ((j / ncc == 1 ? a.Cin : 0) + (j % ncc) * kKC) : j * kKC;
}
// k chunks per accumulator chain of the 3x3 kernel: all of them in bf16 mode; three (3 x 9 taps x 4 = 108 MMAs) in
// fp32-accuracy mode, where the partial sums are added on the CUDA cores (see the epilogue)
__device__ __forceinline__ int halo_segment_chunks(const ConvArgs& a, int ncc) { return a.split ? 3 : ncc; }
__device__ __forceinline__ uint32_t conv_idesc(const ConvArgs& a, int n_cols) {
  // D f32 (bit 4); A / B formats (bits 7-9 / 10-12): 1 = bf16, 0 = fp16; both K-major; N, M
  return (1u << 4) | (a.split ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(kTileCo >> 4) << 24);
}
// accumulator -> value: acc * (per-channel weight unscale / activation scale) + bias; 1 in bf16 mode (fmaf(acc, 1, b) == acc + b)
__device__ __forceinline__ float conv_out_scale(const ConvArgs& a, int co) {
  if (!a.split || co >= a.Cout) return 1.f;
  return __ldg(a.unscale + co) / activation_scale(__ldg(a.x_stat));
}

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool real) {
  const int n = real ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit_to(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// K-major, 128-byte swizzle, 8-row groups 1024 B apart (the layout of both operands)
__device__ __forceinline__ uint64_t kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D[tmem] (+)= A[smem] . B[smem]^T, 128 x 256 x 16, bf16 -> f32
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

#define CONV_TC_LD32(r, taddr)                                                                                             \
  asm volatile(                                                                                                            \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"      \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                           \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),        \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),             \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),            \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                          \
      : "r"(taddr)                                                                                                         \
      : "memory")

#define CONV_TC_ST32(taddr, r)                                                                                             \
  asm volatile(                                                                                                            \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"  \
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),                                                    \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),     \
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),     \
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),     \
      "r"(r[31])                                                                                                           \
      : "memory")

// ---- per-tap form: 1x1 convolutions, and 3x3 when an image row does not fit a tile ----------------------------------------
// 256 flat pixels per tile, one 48 KB stage (weight stage + the tap's pixel rows) per (tap, 64-channel chunk) in a ring of
// four.  (The first version ran producers, MMA issue and epilogue in ONE instruction stream: every stage paid the chain
// "MMA(c-2) done -> copies issued -> wait -> fence -> block barrier -> MMA(c) issued", ~800 clk against 512 clk of tensor
// work, the pipeline drained at every tile boundary and the accumulator was read out while the tensor pipe idled —
// 57 % of the sustained bf16 peak on 768->512, 37 % on the five-head conv.)  Three groups of warps that meet only at
// mbarriers:
//   warps 0-7   producers: im2col pixel rows by cp.async, the weight stage by one bulk copy; they run ahead over tile
//               boundaries, held back only by the ring (`empty`, released by tcgen05.commit)
//   warp 12     MMA issuer: waits for a stage's rows (`full_b`, one arrival per producer warp) and weights (`full_a`),
//               issues four 128x256x16 MMAs, commits the stage; the accumulator alternates between two 256-column
//               halves of tensor memory
//   warps 8-11  epilogue: one TMEM lane quadrant each; tile t is read out, biased, clamped and stored while the MMAs
//               of tile t+1 run into the other half (`acc_full` / `acc_empty`)
// ---- CTA-pair (cta_group::2) helpers, as in pointnet_mlp_tc.cu ----
__device__ __forceinline__ uint32_t cluster_ctarank_conv() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_conv() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster (release: what this thread wrote before is
// visible to whoever the barrier lets through)
__device__ __forceinline__ void mbarrier_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_addr(bar)),
      "r"(rank)
      : "memory");
}
// the same without release semantics: the relay forwards the completion of a bulk copy and has written nothing itself
__device__ __forceinline__ void mbarrier_arrive_cluster_relaxed(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_addr(bar)),
      "r"(rank)
      : "memory");
}
template <int CG>
__device__ __forceinline__ void tc_commit_cg(uint64_t* bar) {
  if constexpr (CG == 1) {
    tc_commit_to(bar);
  } else {   // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_addr(bar)),
                 "h"((uint16_t)3)
                 : "memory");
  }
}
// CG = 2: one instruction of the leader drives both CTAs: M = 256, each CTA its own 128 rows of A and of D and half of B's rows
template <int CG>
__device__ __forceinline__ void umma_ss_cg(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1) {
    umma_ss(d, adesc, bdesc, idesc, accumulate);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

constexpr int kWsThreads = 416;
constexpr int kLag = 3;   // a producer hands stage c over (waits for its own copies of it) three stages after issuing them

// Channels-last bf16 output of one pixel from a warp's transposing tile (tp[channel][33], lane = pixel): the lane gathers its
// pixel's 32 channels (conflict-free: bank = channel + lane), rounds to bf16 and writes 64 contiguous bytes as four 16-byte
// stores.  (With lane = channel, as the accumulator comes out of tensor memory, a store instruction wrote 2 bytes per lane: 32
// such instructions per 32 pixels paced the epilogue of every block that feeds another convolution.)
__device__ __forceinline__ void store_pixel_nhwc_bf16(const float* tp, int lane, __nv_bfloat16* dst, int n_ch, bool vec_ok) {
  if (vec_ok && n_ch == 32) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      __align__(16) __nv_bfloat162 w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) w[e] = __floats2bfloat162_rn(tp[(g * 8 + 2 * e) * 33 + lane], tp[(g * 8 + 2 * e + 1) * 33 + lane]);
      *reinterpret_cast<uint4*>(dst + g * 8) = *reinterpret_cast<const uint4*>(w);
    }
  } else {
    for (int c = 0; c < n_ch; ++c) dst[c] = __float2bfloat16_rn(tp[c * 33 + lane]);
  }
}

// CG = 2: a CTA pair per tile of 256 output channels x 256 pixels (tcgen05.mma.cta_group::2), as in the 3x3 kernel below: each
// CTA streams its own co tile's weight stage and HALF of the tap's pixel rows (128 of 256: every pixel then crosses L2 -> SM once
// for both co tiles), the leader issues, the follower's producers report to the leader's `peer_b`, its idle MMA warp relays the
// weight stages' completions to `peer_a`, commits arrive in both CTAs.
template <int CG>
__global__ void __launch_bounds__(kWsThreads, 1) conv_tc_ws_kernel(ConvArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  float* epi = reinterpret_cast<float*>(ring + kRing * kStage);   // [4][32][33]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kRing * kStage + kEpiTile);
  uint64_t* full_a = bars;
  uint64_t* full_b = bars + kRing;
  uint64_t* empty = bars + 2 * kRing;
  uint64_t* acc_full = bars + 3 * kRing;       // [2]
  uint64_t* acc_empty = bars + 3 * kRing + 2;  // [2] one arrival per epilogue warp of the pair, at the leader
  uint64_t* peer_a = bars + 3 * kRing + 4;     // [kRing] CG = 2, leader: the follower's weight stage has landed (relayed)
  uint64_t* peer_b = bars + 4 * kRing + 4;     // [kRing] CG = 2, leader: the follower's pixel rows have landed
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = CG == 2 ? cluster_ctarank_conv() : 0u;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 12) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_base_s)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_base_s)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  if (tid == 0) {
    for (int s = 0; s < kRing; ++s) {
      mbarrier_init(&full_a[s], 1);
      mbarrier_init(&full_b[s], 8);
      mbarrier_init(&empty[s], 1);
      mbarrier_init(&peer_a[s], 1);
      mbarrier_init(&peer_b[s], 8);
    }
    for (int s = 0; s < 2; ++s) {
      mbarrier_init(&acc_full[s], 1);
      mbarrier_init(&acc_empty[s], 4 * CG);
    }
    mbarrier_init_fence();
  }
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_conv();   // the peer's barriers are initialised before anyone signals them
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  const int HW = a.H * a.W;
  const long long n_px = (long long)a.B * HW;
  const int n_px_tiles = (int)((n_px + kTilePx - 1) / kTilePx);
  const int n_co_tiles = (a.Cout + kTileCo - 1) / kTileCo;
  const int ncc = conv_chunks(a);
  const int n_k = a.taps * ncc;
  // a tile is CG co tiles x 256 pixels, walked by a cluster of CG CTAs in lock step (host: n_co_tiles % CG == 0)
  const int n_co_groups = n_co_tiles / CG;
  const int n_tiles = n_px_tiles * n_co_groups;
  const int first_tile = (int)blockIdx.x / CG, tile_step = (int)gridDim.x / CG;
  const int my_tiles = (n_tiles - first_tile + tile_step - 1) / tile_step;
  const int total = my_tiles * n_k;   // stages this CTA streams
  constexpr int kRowsJ = 8 / CG;      // this CTA's pixel rows of a stage: 256 / CG, 32 per step of j

  if (warp < 8) {
    // ---- producers ----
    const int row0 = tid >> 3, chunk = tid & 7;
    const uint32_t dst_off = (uint32_t)(row0 * 128 + ((chunk ^ (row0 & 7)) << 4));
    int py[8], pxx[8];
    long long poff[8];
    int tile = first_tile, i = 0;   // the stage being produced: k-stage i of `tile`
    const uint8_t* wtile = nullptr;
    for (int it = 0; it < total + kLag; ++it) {
      // first hand over the oldest stage in flight (nothing below may delay the MMAs: the ring wait of the new stage is
      // a wait for an MMA to FINISH, and put in front of this arrival it serialised the tensor pipe — 2.9 ms vs 2.3)
      const int c = it - kLag;
      if (c >= 0) {
        cp_async_wait_group<kLag - 1>();   // this thread's rows of stage c have landed (groups it-kLag+1 .. it-1 may be pending)
        fence_proxy_async_shared();        // ... and are ordered before the tensor core's asynchronous-proxy reads
        __syncwarp();
        if (lane == 0) {
          if (CG == 1 || rank == 0) mbarrier_arrive(&full_b[(uint32_t)c % kRing]);
          else mbarrier_arrive_cluster(&peer_b[(uint32_t)c % kRing], 0);   // the leader issues the MMAs that read these rows
        }
      }
      if (it < total) {
        if (i == 0) {
          const int co_tile = (tile % n_co_groups) * CG + (int)rank, px_tile = tile / n_co_groups;
          const long long px0 = (long long)px_tile * kTilePx + (long long)rank * (kTilePx / CG);
#pragma unroll
          for (int j = 0; j < kRowsJ; ++j) {
            const long long n = px0 + row0 + 32 * j;
            if (n < n_px) {
              const int p = (int)(n % HW);
              py[j] = p / a.W;
              pxx[j] = p - py[j] * a.W;
              poff[j] = n * a.x_pitch + chunk * 8;
            } else {
              py[j] = -100000;
              pxx[j] = 0;
              poff[j] = 0;
            }
          }
          wtile = a.wimg + (size_t)co_tile * n_k * kStageA;
        }
        const uint32_t slot = (uint32_t)it % kRing;
        if (it >= kRing) mbarrier_wait(&empty[slot], (((uint32_t)it / kRing) - 1) & 1);
        const int tap = i / ncc, cc = i - tap * ncc;
        const int dy = a.taps == 9 ? tap / 3 - 1 : 0, dx = a.taps == 9 ? tap % 3 - 1 : 0;
        const long long shift = ((long long)dy * a.W + dx) * a.x_pitch + conv_chunk_channel(a, cc);
        const uint32_t bdst = smem_addr(ring + slot * kStage + kStageA) + dst_off;
#pragma unroll
        for (int j = 0; j < kRowsJ; ++j) {
          const int yy = py[j] + dy, xx = pxx[j] + dx;
          const bool ok = yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
          cp_async16_zfill(bdst + j * 4096, a.x + (ok ? poff[j] + shift : 0), ok);
        }
        if (warp == 0 && elect_one()) {
          mbarrier_expect_tx(&full_a[slot], kStageA);
          bulk_copy_global_to_shared(ring + slot * kStage, wtile + (size_t)i * kStageA, kStageA, &full_a[slot]);
        }
        if (++i == n_k) {
          i = 0;
          tile += tile_step;
        }
      }
      cp_async_commit_group();
    }
  } else if (warp == 12 && CG == 2 && rank != 0) {
    // ---- follower of a pair: forward the completion of every weight stage to the leader ----
    for (int c = 0; c < total; ++c) {
      const uint32_t slot = (uint32_t)c % kRing;
      mbarrier_wait(&full_a[slot], ((uint32_t)c / kRing) & 1);
      if (lane == 0) mbarrier_arrive_cluster_relaxed(&peer_a[slot], 0);
      __syncwarp();
    }
  } else if (warp == 12) {
    // ---- MMA issuer ----
    const uint32_t idesc = conv_idesc(a, kTilePx) + (CG == 2 ? ((uint32_t)(kTileCo >> 4) << 24) : 0u);   // M = 128 CG
    int i = 0, tile_seq = 0;
    for (int c = 0; c < total; ++c) {
      const uint32_t slot = (uint32_t)c % kRing, use = (uint32_t)c / kRing;
      const uint32_t buf = tile_seq & 1;
      if (i == 0 && tile_seq >= 2) mbarrier_wait(&acc_empty[buf], ((tile_seq >> 1) - 1) & 1);
      mbarrier_wait(&full_b[slot], use & 1);
      mbarrier_wait(&full_a[slot], use & 1);
      if constexpr (CG == 2) {
        mbarrier_wait(&peer_b[slot], use & 1);
        mbarrier_wait(&peer_a[slot], use & 1);
      }
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t a_addr = smem_addr(ring + slot * kStage), b_addr = a_addr + kStageA;
        const uint32_t d = tmem + buf * kTilePx;
#pragma unroll
        for (int s = 0; s < kKC / 16; ++s)
          umma_ss_cg<CG>(d, kmajor_sw128_desc(a_addr + s * 32), kmajor_sw128_desc(b_addr + s * 32), idesc, !(i == 0 && s == 0));
        tc_commit_cg<CG>(&empty[slot]);
        if (i == n_k - 1) tc_commit_cg<CG>(&acc_full[buf]);
      }
      __syncwarp();
      if (++i == n_k) {
        i = 0;
        ++tile_seq;
      }
    }
  } else {
    // ---- epilogue: warp 8+q owns TMEM lanes [32q, 32q+32) = 32 output channels; each 32 x 32 piece is turned through a
    // private shared-memory tile so that a store instruction writes contiguous pixels of ONE channel plane (see the 3x3
    // kernel below) ----
    const bool nhwc_vec = (a.out_ct & 7) == 0 && (a.out_coff & 7) == 0 && (reinterpret_cast<uintptr_t>(a.out_nhwc) & 15) == 0;
    const int quad = warp & 3;
    float* tp = epi + quad * 32 * 33;
    int tile = first_tile;
    for (int tile_seq = 0; tile_seq < my_tiles; ++tile_seq, tile += tile_step) {
      const int co_tile = (tile % n_co_groups) * CG + (int)rank, px_tile = tile / n_co_groups;
      const long long px0 = (long long)px_tile * kTilePx;
      const uint32_t buf = tile_seq & 1;
      const int co0 = co_tile * kTileCo + quad * 32;
      const int co = co0 + lane;
      const float bias = (a.bias && co < a.Cout) ? __ldg(a.bias + co) : 0.f;
      const float oscale = conv_out_scale(a, co);
      const int n_ch = a.Cout - co0 < 32 ? a.Cout - co0 : 32;
      mbarrier_wait(&acc_full[buf], (tile_seq >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll 1
      for (int q = 0; q < kTilePx / 32; ++q) {
        uint32_t r[32];
        const int col0 = q * 32;
        CONV_TC_LD32(r, tmem + ((uint32_t)(quad * 32) << 16) + buf * kTilePx + (uint32_t)col0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const long long n0 = px0 + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = fmaf(__uint_as_float(r[j]), oscale, bias);
          if (a.relu) v = fmaxf(v, 0.f);
          r[j] = __float_as_uint(v);
          tp[lane * 33 + j] = v;
        }
        __syncwarp();
        if (a.out_nhwc && n0 + lane < n_px && n_ch > 0)
          store_pixel_nhwc_bf16(tp, lane, a.out_nhwc + (size_t)(n0 + lane) * a.out_ct + a.out_coff + co0, n_ch, nhwc_vec);
        if (a.out) {
          const long long n = n0 + lane;   // this lane's pixel
          const bool px_ok = n < n_px;
          const int b = px_ok ? (int)(n / HW) : 0, p = px_ok ? (int)(n - (long long)b * HW) : 0;
          float* dst = a.out + ((size_t)b * a.Cout + co0) * HW + p;
          for (int c = 0; c < n_ch; ++c) {
            const float v = tp[c * 33 + lane];
            if (px_ok) dst[(size_t)c * HW] = v;
          }
        }
        __syncwarp();
      }
      tc_fence_before_sync();   // the tensor-memory loads above are complete (wait::ld) before the half is handed back
      __syncwarp();
      if (lane == 0) {
        if (CG == 1 || rank == 0) mbarrier_arrive(&acc_empty[buf]);
        else mbarrier_arrive_cluster(&acc_empty[buf], 0);
      }
    }
  }

  cp_async_wait_group<0>();
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_conv();   // no CTA leaves (or frees tensor memory) while its peer may still signal or compute
  if (warp == 12) {
    tc_fence_after_sync();
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem));
  }
}

// ---- 1x1 form fed by TMA tensor loads ------------------------------------------------------------------------------------------
// The per-tap kernel above moves a stage's 256 pixel rows with 2,048 sixteen-byte cp.async per CTA, and a 1x1 convolution has
// only four MMAs (512 clk) per stage to hide them behind: measured, the kernel is paced by its ring turn-around (2,400 clk per
// stage on camera_proj.3, 0.29 of the bf16 peak; neither resident weights, nor CTA pairs, nor an L2 prefetch hint changed that —
// the copies' own latency did: ~8,000 clk from issue to hand-over with ~6,000 sixteen-byte requests in flight per SM).  Here the
// input is described ONCE as a 2-D tensor (rows = pixels, columns = channels, host: cuTensorMapEncodeTiled) and a stage's pixel
// rows are ONE instruction: cp.async.bulk.tensor.2d with a 64-channel x 256-pixel box, 128-byte swizzle — the layout the MMA
// descriptor reads — and zero fill past the last pixel.  No producer warps: one elected lane issues the tensor load and the
// weight stage's bulk copy onto the same barrier.  192 threads: warp 0 copies, warp 1 MMA issue, warps 2-5 epilogue.
constexpr int kTmaThreads = 192;

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_addr(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_addr(bar))
               : "memory");
}

__global__ void __launch_bounds__(kTmaThreads, 1) conv1x1_tma_kernel(ConvArgs a, const __grid_constant__ CUtensorMap xmap) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  float* epi = reinterpret_cast<float*>(ring + kRing * kStage);   // [4][32][33]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kRing * kStage + kEpiTile);
  uint64_t* full = bars;                       // [kRing] transaction bytes: weight stage + pixel box
  uint64_t* empty = bars + kRing;              // [kRing] tcgen05.commit
  uint64_t* acc_full = bars + 2 * kRing;       // [2]
  uint64_t* acc_empty = bars + 2 * kRing + 2;  // [2] one arrival per epilogue warp
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < kRing; ++s) {
      mbarrier_init(&full[s], 1);
      mbarrier_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbarrier_init(&acc_full[s], 1);
      mbarrier_init(&acc_empty[s], 4);
    }
    mbarrier_init_fence();
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&xmap)) : "memory");
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  const int HW = a.H * a.W;
  const long long n_px = (long long)a.B * HW;
  const int n_px_tiles = (int)((n_px + kTilePx - 1) / kTilePx);
  const int n_co_tiles = (a.Cout + kTileCo - 1) / kTileCo;
  const int ncc = conv_chunks(a);              // taps == 1: one stage per k chunk
  const int n_tiles = n_px_tiles * n_co_tiles;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = my_tiles * ncc;

  if (warp == 0) {
    // ---- copies: per stage one tensor load (the tap's 256 pixel rows x 64 channels) and one bulk copy (the weight stage) ----
    int tile = blockIdx.x, i = 0;
    for (int c = 0; c < total; ++c) {
      const uint32_t slot = (uint32_t)c % kRing;
      if (c >= kRing) mbarrier_wait(&empty[slot], (((uint32_t)c / kRing) - 1) & 1);
      if (elect_one()) {
        const int co_tile = tile % n_co_tiles, px_tile = tile / n_co_tiles;
        mbarrier_expect_tx(&full[slot], kStage);
        bulk_copy_global_to_shared(ring + slot * kStage, a.wimg + ((size_t)co_tile * ncc + i) * kStageA, kStageA, &full[slot]);
        tma_load_2d(ring + slot * kStage + kStageA, &xmap, conv_chunk_channel(a, i), px_tile * kTilePx, &full[slot]);
      }
      __syncwarp();
      if (++i == ncc) {
        i = 0;
        tile += gridDim.x;
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    const uint32_t idesc = conv_idesc(a, kTilePx);
    int i = 0, tile_seq = 0;
    for (int c = 0; c < total; ++c) {
      const uint32_t slot = (uint32_t)c % kRing, use = (uint32_t)c / kRing;
      const uint32_t buf = tile_seq & 1;
      if (i == 0 && tile_seq >= 2) mbarrier_wait(&acc_empty[buf], ((tile_seq >> 1) - 1) & 1);
      mbarrier_wait(&full[slot], use & 1);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t a_addr = smem_addr(ring + slot * kStage), b_addr = a_addr + kStageA;
        const uint32_t d = tmem + buf * kTilePx;
#pragma unroll
        for (int s = 0; s < kKC / 16; ++s)
          umma_ss(d, kmajor_sw128_desc(a_addr + s * 32), kmajor_sw128_desc(b_addr + s * 32), idesc, !(i == 0 && s == 0));
        tc_commit_to(&empty[slot]);
        if (i == ncc - 1) tc_commit_to(&acc_full[buf]);
      }
      __syncwarp();
      if (++i == ncc) {
        i = 0;
        ++tile_seq;
      }
    }
  } else {
    // ---- epilogue (as in conv_tc_ws_kernel): warp w owns TMEM lanes [32 (w % 4), +32) ----
    const bool nhwc_vec = (a.out_ct & 7) == 0 && (a.out_coff & 7) == 0 && (reinterpret_cast<uintptr_t>(a.out_nhwc) & 15) == 0;
    const int quad = warp & 3;
    float* tp = epi + quad * 32 * 33;
    int tile = blockIdx.x;
    for (int tile_seq = 0; tile_seq < my_tiles; ++tile_seq, tile += gridDim.x) {
      const int co_tile = tile % n_co_tiles, px_tile = tile / n_co_tiles;
      const long long px0 = (long long)px_tile * kTilePx;
      const uint32_t buf = tile_seq & 1;
      const int co0 = co_tile * kTileCo + quad * 32;
      const int co = co0 + lane;
      const float bias = (a.bias && co < a.Cout) ? __ldg(a.bias + co) : 0.f;
      const float oscale = conv_out_scale(a, co);
      const int n_ch = a.Cout - co0 < 32 ? a.Cout - co0 : 32;
      mbarrier_wait(&acc_full[buf], (tile_seq >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll 1
      for (int q = 0; q < kTilePx / 32; ++q) {
        uint32_t r[32];
        const int col0 = q * 32;
        CONV_TC_LD32(r, tmem + ((uint32_t)(quad * 32) << 16) + buf * kTilePx + (uint32_t)col0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const long long n0 = px0 + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = fmaf(__uint_as_float(r[j]), oscale, bias);
          if (a.relu) v = fmaxf(v, 0.f);
          tp[lane * 33 + j] = v;
        }
        __syncwarp();
        if (a.out_nhwc && n0 + lane < n_px && n_ch > 0)
          store_pixel_nhwc_bf16(tp, lane, a.out_nhwc + (size_t)(n0 + lane) * a.out_ct + a.out_coff + co0, n_ch, nhwc_vec);
        if (a.out) {
          const long long n = n0 + lane;   // this lane's pixel
          const bool px_ok = n < n_px;
          const int b = px_ok ? (int)(n / HW) : 0, p = px_ok ? (int)(n - (long long)b * HW) : 0;
          float* dst = a.out + ((size_t)b * a.Cout + co0) * HW + p;
          for (int c = 0; c < n_ch; ++c) {
            const float v = tp[c * 33 + lane];
            if (px_ok) dst[(size_t)c * HW] = v;
          }
        }
        __syncwarp();
      }
      tc_fence_before_sync();   // the tensor-memory loads above are complete (wait::ld) before the half is handed back
      __syncwarp();
      if (lane == 0) mbarrier_arrive(&acc_empty[buf]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point table: libb200bev.so keeps loading where there is no driver
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// (n_px, x_pitch) elements of 2 bytes, rows x_pitch apart, as boxes of 256 rows x 64 columns, 128-byte swizzle, zeros out of bounds
inline bool make_pixel_map(CUtensorMap* map, const void* x, long long n_px, int x_pitch, bool fp16) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)x_pitch, (cuuint64_t)n_px};
  const cuuint64_t strides[1] = {(cuuint64_t)x_pitch * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)kTilePx};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- 3x3 form with the pixel rows shared by the nine taps (default for 3x3 when an image row fits) ---------------------
// In the kernels above every tap re-reads its 256 pixel rows from L2: 48 KB per 512 clk of tensor work, 96 B/clk per SM,
// and with ~2,000 clk from "slot free" to "MMA issued" the four-slot ring cannot keep that in flight (ncu: tensor pipe
// active 49 % on the 768->512 block).  A 3x3 tap is the SAME pixels displaced by (dy, dx).  Here a frame is walked in its
// padded row-major order — flat index F = y (W + 1) + x with ONE zero pixel (x = W) between consecutive image rows — and a
// tile is N consecutive flat indices starting at n0 (any n0: tiles need not begin at an image row), laid out in shared
// memory from flat index n0 - (W + 1) - 1 on: one image row and one pixel of halo in front, the same behind.
// Output column n (flat index n0 + n; x = W is a dummy column) then needs, for tap (dy, dx), block row
// n + (1 + dy)(W + 1) + (1 + dx): one block of N + 2 (W + 1) + 2 rows per 64 input channels serves all nine taps, each tap
// being the same shared-memory block entered `off` rows later (the 128-byte swizzle is a function of the shared-memory
// address, so a row-displaced descriptor reads every row with the phase it was written with).  Zero padding of the
// convolution comes from the pad pixels and from halo rows outside the frame, zero-filled by the copy.  Pixel traffic
// drops ninefold (41 B/clk per SM with the weights), and the copies of a block have nine taps of MMAs to hide behind.
//   warps 0-7 pixel blocks (cp.async, two buffers) | warp 13 weight stages (bulk copies, ring of 4) | warp 12 MMA issue
//   | warps 8-11 epilogue (accumulator double-buffered in tensor memory), all meeting at mbarriers only.
constexpr int kHaloThreads = 448;
constexpr int kHaloMaxRows = 480;                       // 60 KB per pixel block
constexpr int kHaloBlock = kHaloMaxRows * 128;
constexpr int kHaloEpi = 4 * 32 * 33 * 4;                // per epilogue warp: a 32 x 32 transposing tile
constexpr int kHaloSmem = 2 * kHaloBlock + kRing * kStageA + kHaloEpi + 1024 + 256;
constexpr int kHaloRowsPerThread = kHaloMaxRows / 32;   // 15

// A frame is walked in its PADDED row-major order: flat index F = y (W + 1) + x, x = W being the zero pixel between image rows.
// A tile is N consecutive flat indices, wherever they start — not whole image rows: a 100-pixel-wide map gets tiles of 256
// columns (2.5 rows) instead of 208 (two rows), 23 tiles per frame instead of 29.
struct HaloGeom {
  int N;       // MMA columns = flat indices per tile (a multiple of 16)
  int Q;       // rows of a pixel block: N + 2 (W + 1) + 2
  int tiles_per_frame;
};

// max_cols: 256, the two accumulator halves of tensor memory (in fp32-accuracy mode one half carries a tile's running total and
// the other its partial sums, see the epilogue)
__host__ __device__ inline bool halo_geometry(int H, int W, HaloGeom* g, int max_cols = kTilePx) {
  const int W1 = W + 1, total = H * W1;
  int n_max = (kHaloMaxRows - 2 * W1 - 2) & ~15;          // the block of N + 2 (W + 1) + 2 rows has to fit
  if (n_max > max_cols) n_max = max_cols;
  if (n_max < 16 || total < 1) return false;
  const int need = (total + 15) & ~15;
  if (need > n_max && n_max < 128) return false;          // very wide maps: tiles that narrow lose to the per-tap kernel
  const int N = need < n_max ? need : n_max;
  g->N = N;
  g->Q = N + 2 * W1 + 2;
  g->tiles_per_frame = (total + N - 1) / N;
  return true;
}

// Descriptor of a block entered a whole number of 128-byte rows after its 1024-byte-aligned base.  Measured on B200
// (the 3x3 parity tests run with both settings): the swizzle is applied to the absolute shared-memory address, so the
// descriptor's "matrix base offset" field stays 0; setting it to (address >> 7) & 7 gives wrong results.
__device__ __forceinline__ uint64_t kmajor_sw128_desc_at(uint32_t saddr, int base_offset_mode) {
  uint64_t d = kmajor_sw128_desc(saddr);
  if (base_offset_mode) d |= (uint64_t)((saddr >> 7) & 7) << 49;
  return d;
}


// CG = 2: a CTA PAIR (thread-block cluster of two) computes 256 output channels x N pixels per tile with tcgen05.mma.cta_group::2.
// Each CTA keeps its own co tile's weight ring (the A operand, 128 rows) and HALF of the pixel block (the B operand: the rows of
// N/2 output columns plus the halo), so the tensor core reads 8 KB of shared memory per 128 clk and SM instead of 12, and a CTA
// copies 0.7 of the pixel rows it copied alone — shared-memory bandwidth was what held the single-CTA kernel at 0.78 of the bf16
// peak.  The leader (cluster rank 0) issues every MMA; the follower's producers report their half block to the
// leader's `peer_blk`, its otherwise idle MMA warp relays the completion of its weight stages to `peer_a` (a bulk copy can
// only signal a barrier of the CTA it writes to), tcgen05.commit arrives on the barriers of BOTH CTAs, and both CTAs' epilogue
// warps hand the accumulator halves back to the leader.
template <int CG>
__global__ void __launch_bounds__(kHaloThreads, 1) conv3x3_tc_halo_kernel(ConvArgs a, HaloGeom geo, int base_offset_mode /* 0: see kmajor_sw128_desc_at */) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* blocks = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);   // [2][kHaloBlock]
  uint8_t* wring = blocks + 2 * kHaloBlock;                                          // [kRing][kStageA]
  float* epi = reinterpret_cast<float*>(wring + kRing * kStageA);                    // [4][32][33]
  uint64_t* bars = reinterpret_cast<uint64_t*>(wring + kRing * kStageA + kHaloEpi);
  uint64_t* full_a = bars;                     // [kRing]
  uint64_t* empty_a = bars + kRing;            // [kRing]
  uint64_t* full_blk = bars + 2 * kRing;       // [2]
  uint64_t* empty_blk = bars + 2 * kRing + 2;  // [2]
  uint64_t* acc_full = bars + 2 * kRing + 4;   // [2]
  uint64_t* acc_empty = bars + 2 * kRing + 6;  // [2] one arrival per epilogue warp of the pair, at the leader
  uint64_t* peer_a = bars + 2 * kRing + 8;     // [kRing] CG = 2, leader: the follower's weight stage has landed (relayed)
  uint64_t* peer_blk = bars + 3 * kRing + 8;   // [2]     CG = 2, leader: the follower's half block has landed (its producer warps)
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = CG == 2 ? cluster_ctarank_conv() : 0u;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 12) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_base_s)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_base_s)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  if (tid == 0) {
    for (int s = 0; s < kRing; ++s) {
      mbarrier_init(&full_a[s], 1);
      mbarrier_init(&empty_a[s], 1);
      mbarrier_init(&peer_a[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbarrier_init(&full_blk[s], 8);
      mbarrier_init(&empty_blk[s], 1);
      mbarrier_init(&acc_full[s], 1);
      mbarrier_init(&acc_empty[s], 4 * CG);
      mbarrier_init(&peer_blk[s], 8);
    }
    mbarrier_init_fence();
  }
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_conv();   // the peer's barriers are initialised before anyone signals them
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  const int W1 = a.W + 1, HW = a.H * a.W;
  const int n_co_tiles = (a.Cout + kTileCo - 1) / kTileCo;
  const int ncc = conv_chunks(a);
  // a tile is CG co tiles x N pixels, walked by a cluster of CG CTAs in lock step (host: n_co_tiles % CG == 0)
  const int n_co_groups = n_co_tiles / CG;
  const int n_tiles = a.B * geo.tiles_per_frame * n_co_groups;
  const int first_tile = (int)blockIdx.x / CG, tile_step = (int)gridDim.x / CG;
  const int my_tiles = (n_tiles - first_tile + tile_step - 1) / tile_step;
  // tile -> (this CTA's co tile, frame, first flat index): the co groups of one pixel tile are neighbours in the grid (shared pixels in L2)
  auto tile_coords = [&](int tile, int& co_tile, int& b, int& n0) {
    co_tile = (tile % n_co_groups) * CG + (int)rank;
    const int pt = tile / n_co_groups;
    b = pt / geo.tiles_per_frame;
    n0 = (pt - b * geo.tiles_per_frame) * geo.N;
  };
  const int n_half = geo.N / CG;                       // output columns whose pixel rows this CTA holds
  const int q_rows = n_half + 2 * W1 + 2;              // rows of this CTA's pixel block

  if (warp < 8) {
    // ---- pixel blocks: one per (tile, 64-channel chunk) ----
    const int row0 = tid >> 3, chunk = tid & 7;
    const uint32_t dst_off = (uint32_t)(row0 * 128 + ((chunk ^ (row0 & 7)) << 4));
    const int n_blocks = my_tiles * ncc;
    int tile = first_tile, cc = 0;
    long long goff[kHaloRowsPerThread];   // element offset of this thread's piece of padded row q = row0 + 32 j, or -1: zeros
    for (int bi = 0; bi <= n_blocks; ++bi) {
      if (bi >= 1) {
        cp_async_wait_group<0>();      // block bi-1 has landed (this thread's part)
        fence_proxy_async_shared();
        __syncwarp();
        if (lane == 0) {
          if (CG == 1 || rank == 0) mbarrier_arrive(&full_blk[(bi - 1) & 1]);
          else mbarrier_arrive_cluster(&peer_blk[(bi - 1) & 1], 0);      // the leader issues the MMAs that read this half block
        }
      }
      if (bi < n_blocks) {
        if (cc == 0) {
          int co_tile, b, n0;
          tile_coords(tile, co_tile, b, n0);
          n0 += (int)rank * n_half;                    // this CTA's columns start here
#pragma unroll
          for (int j = 0; j < kHaloRowsPerThread; ++j) {
            // block row q holds the pixel of flat index n0 + q - (W + 1) - 1 (one image row and one pixel before the tile's
            // first output); the slot x = W of every row, and everything outside the frame, is zero
            const int q = row0 + 32 * j;
            const int gflat = n0 + q - W1 - 1;
            const int y = gflat >= 0 ? gflat / W1 : -1, x = gflat - y * W1;
            goff[j] = (q < q_rows && gflat >= 0 && y < a.H && x < a.W) ? (((long long)b * a.H + y) * a.W + x) * a.x_pitch + chunk * 8 : -1;
          }
        }
        const int buf = bi & 1;
        if (bi >= 2) mbarrier_wait(&empty_blk[buf], ((bi >> 1) - 1) & 1);
        const uint32_t bdst = smem_addr(blocks + buf * kHaloBlock) + dst_off;
        const int c_first = conv_chunk_channel(a, cc);
#pragma unroll
        for (int j = 0; j < kHaloRowsPerThread; ++j) {
          if (row0 + 32 * j < q_rows) {
            const bool ok = goff[j] >= 0;
            cp_async16_zfill(bdst + j * 4096, a.x + (ok ? goff[j] + c_first : 0), ok);
          }
        }
        cp_async_commit_group();
        if (++cc == ncc) {
          cc = 0;
          tile += tile_step;
        }
      }
    }
  } else if (warp == 13) {
    // ---- weight stages: (tile, chunk, tap) in the order the MMAs use them ----
    const int total = my_tiles * ncc * 9;
    int tile = first_tile, cc = 0, tap = 0;
    int co_tile, b, y0;
    tile_coords(tile, co_tile, b, y0);
    for (int g = 0; g < total; ++g) {
      const uint32_t slot = (uint32_t)g % kRing;
      if (g >= kRing) mbarrier_wait(&empty_a[slot], (((uint32_t)g / kRing) - 1) & 1);
      if (elect_one()) {
        mbarrier_expect_tx(&full_a[slot], kStageA);
        bulk_copy_global_to_shared(wring + slot * kStageA, a.wimg + ((size_t)co_tile * 9 * ncc + (size_t)tap * ncc + cc) * kStageA,
                                   kStageA, &full_a[slot]);
      }
      __syncwarp();
      if (++tap == 9) {
        tap = 0;
        if (++cc == ncc) {
          cc = 0;
          tile += tile_step;
          tile_coords(tile, co_tile, b, y0);
        }
      }
    }
  } else if (warp == 12 && CG == 2 && rank != 0) {
    // ---- follower of a pair: forward the completion of every weight stage to the leader ----
    const int total = my_tiles * ncc * 9;
    for (int g = 0; g < total; ++g) {
      const uint32_t slot = (uint32_t)g % kRing;
      mbarrier_wait(&full_a[slot], ((uint32_t)g / kRing) & 1);
      if (lane == 0) mbarrier_arrive_cluster_relaxed(&peer_a[slot], 0);
      __syncwarp();
    }
  } else if (warp == 12) {
    // ---- MMA issuer ----
    const uint32_t idesc = conv_idesc(a, geo.N) + (CG == 2 ? ((uint32_t)(kTileCo >> 4) << 24) : 0u);   // M = 128 CG
    const int seg = halo_segment_chunks(a, ncc);
    // Accumulator halves: tile t keeps its TOTAL in half t & 1 — the first K segment accumulates straight into it — and every
    // later segment (fp32-accuracy mode) goes to the other half, from where the epilogue adds it to the total (see there).
    // A half is reused when the epilogue has released it: uses0 / uses1 count how often each half has been handed to the MMAs.
    int g = 0, bi = 0, uses0 = 0, uses1 = 0;
    for (int tile_seq = 0; tile_seq < my_tiles; ++tile_seq) {
      uint32_t buf = 0, d = 0;
      for (int cc = 0; cc < ncc; ++cc, ++bi) {
        const bool seg_first = cc % seg == 0, seg_last = cc % seg == seg - 1 || cc == ncc - 1;
        if (seg_first) {
          buf = (uint32_t)((tile_seq & 1) ^ (cc == 0 ? 0 : 1));
          const int k = buf ? uses1++ : uses0++;
          if (k >= 1) mbarrier_wait(&acc_empty[buf], (k - 1) & 1);
          d = tmem + buf * kTilePx;
        }
        mbarrier_wait(&full_blk[bi & 1], (bi >> 1) & 1);
        if constexpr (CG == 2) mbarrier_wait(&peer_blk[bi & 1], (bi >> 1) & 1);
        const uint32_t blk = smem_addr(blocks + (bi & 1) * kHaloBlock);
        for (int tap = 0; tap < 9; ++tap, ++g) {
          const uint32_t slot = (uint32_t)g % kRing;
          mbarrier_wait(&full_a[slot], ((uint32_t)g / kRing) & 1);
          if constexpr (CG == 2) mbarrier_wait(&peer_a[slot], ((uint32_t)g / kRing) & 1);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t a_addr = smem_addr(wring + slot * kStageA);
            const uint32_t b_addr = blk + (uint32_t)((tap / 3) * W1 + tap % 3) * 128;   // (1+dy)(W+1) + (1+dx) rows in
#pragma unroll
            for (int s = 0; s < kKC / 16; ++s)
              umma_ss_cg<CG>(d, kmajor_sw128_desc(a_addr + s * 32), kmajor_sw128_desc_at(b_addr + s * 32, base_offset_mode), idesc,
                             !(seg_first && tap == 0 && s == 0));
            tc_commit_cg<CG>(&empty_a[slot]);
            if (tap == 8) tc_commit_cg<CG>(&empty_blk[bi & 1]);
            if (tap == 8 && seg_last) tc_commit_cg<CG>(&acc_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue: warp 8+q owns TMEM lanes [32q, 32q+32) = 32 output channels ----
    // A TMEM load gives a lane ONE channel and 32 columns; stored like that, a warp's store instruction touches 32
    // channel planes with 4 bytes each (32 partial sectors: the stores, not the MMAs, set the pace of the small blocks).
    // Each warp turns its 32 x 32 piece through a private shared-memory tile instead: then a lane is one PIXEL, and a
    // store instruction writes up to 128 contiguous bytes of one channel plane.
    const int quad = warp & 3;
    float* tp = epi + quad * 32 * 33;
    const int seg = halo_segment_chunks(a, ncc), n_seg = (ncc + seg - 1) / seg;
    int tile = first_tile, uses0 = 0, uses1 = 0;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
    auto release = [&](uint32_t half) {     // hand an accumulator half back to the MMA issuer (the leader's barrier)
      if (CG == 1 || rank == 0) mbarrier_arrive(&acc_empty[half]);
      else mbarrier_arrive_cluster(&acc_empty[half], 0);
    };
    const bool nhwc_vec = (a.out_ct & 7) == 0 && (a.out_coff & 7) == 0 && (reinterpret_cast<uintptr_t>(a.out_nhwc) & 15) == 0;
    for (int tile_seq = 0; tile_seq < my_tiles; ++tile_seq, tile += tile_step) {
      int co_tile, b, n0;
      tile_coords(tile, co_tile, b, n0);
      const int co0 = co_tile * kTileCo + quad * 32;
      const float bias = (a.bias && co0 + lane < a.Cout) ? __ldg(a.bias + co0 + lane) : 0.f;
      const float oscale = conv_out_scale(a, co0 + lane);
      const int n_ch = a.Cout - co0 < 32 ? a.Cout - co0 : 32;   // channels of this warp that exist (may be <= 0)
      float* oplane = a.out + ((size_t)b * a.Cout + co0) * HW;
      // K segments (fp32-accuracy mode): the tensor core TRUNCATES the fp32 accumulator once per instruction, a bias of
      // ~2.3e-8 of the sum per MMA in a chain (measured: 1.1e-5 at 432 MMAs, 2.9e-5 at 1,296).  So a chain is `seg` chunks
      // (108 MMAs) long, and the partial sums are added — round to nearest, on the CUDA cores — to the tile's total.
      // Tensor memory is two halves of 256 columns: the tile's first segment accumulates straight into half T = tile & 1 (the
      // total), every later one into the other half P, and this warp adds P to T (one tcgen05.ld of each, 32 adds, one
      // tcgen05.st per 32 columns; only its own lanes, so no synchronisation) and hands P back.  After the last segment the
      // total is read once more, scaled, biased and stored, and T goes back too — by then the next tile's first segment is
      // accumulating in what was P.  (First version: two partial halves AND a total, 3 x 160 columns: tiles of at most 160
      // pixels, 112 for a 100-pixel-wide map, and the A operand re-read from shared memory every 56 clk; a version before
      // that added the partial sums into the output tile in global memory: 24.9 ms for the fusion module against 5.8.)
      const uint32_t T = (uint32_t)(tile_seq & 1), P = T ^ 1u;
      {
        const int k = T ? uses1++ : uses0++;
        mbarrier_wait(&acc_full[T], k & 1);
        tc_fence_after_sync();
      }
      for (int sg = 1; sg < n_seg; ++sg) {
        const int k = P ? uses1++ : uses0++;
        mbarrier_wait(&acc_full[P], k & 1);
        tc_fence_after_sync();
#pragma unroll 1
        for (int col0 = 0; col0 < geo.N; col0 += 32) {
          uint32_t r[32], t[32];
          CONV_TC_LD32(r, lane_base + P * kTilePx + (uint32_t)col0);
          CONV_TC_LD32(t, lane_base + T * kTilePx + (uint32_t)col0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(t[j]));
          CONV_TC_ST32(lane_base + T * kTilePx + (uint32_t)col0, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");   // the total is re-read by this thread
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) release(P);
      }
#pragma unroll 1
      for (int col0 = 0; col0 < geo.N; col0 += 32) {
        uint32_t r[32];
        CONV_TC_LD32(r, lane_base + T * kTilePx + (uint32_t)col0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = fmaf(__uint_as_float(r[j]), oscale, bias);
          if (a.relu) v = fmaxf(v, 0.f);
          r[j] = __float_as_uint(v);
          tp[lane * 33 + j] = v;
        }
        __syncwarp();
        // this lane's pixel: column n = col0 + lane, flat index n0 + n
        const int n = col0 + lane;
        const int y = (n0 + n) / W1, x = n0 + n - y * W1;
        const bool px_ok = n < geo.N && x < a.W && y < a.H;
        if (a.out_nhwc && px_ok && n_ch > 0)
          store_pixel_nhwc_bf16(tp, lane, a.out_nhwc + (((size_t)b * a.H + y) * a.W + x) * a.out_ct + a.out_coff + co0, n_ch, nhwc_vec);
        float* dst = oplane + (px_ok ? y * a.W + x : 0);
        if (a.out) {
          for (int c = 0; c < n_ch; ++c) {
            const float v = tp[c * 33 + lane];
            if (px_ok) dst[(size_t)c * HW] = v;
          }
        }
        __syncwarp();
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) release(T);
    }
  }

  cp_async_wait_group<0>();
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_conv();   // no CTA leaves (or frees tensor memory) while its peer may still signal or compute
  if (warp == 12) {
    tc_fence_after_sync();
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem));
  }
}

// (Cout, Cin, taps) fp32 -> stages [co tile][tap][ci chunk] of [128 co][64 ci] bf16, 16-byte chunk c of row r at c ^ (r & 7)
__global__ void __launch_bounds__(256) conv_pack_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, uint8_t* __restrict__ img,
                                                        long long n_chunks) {
  const int ncc = Cin / kKC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_chunks; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 7);
    const int r = (int)((i >> 3) & 127);
    const long long stage = i >> 10;
    const int cc = (int)(stage % ncc);
    const int tap = (int)((stage / ncc) % taps);
    const int ct = (int)(stage / ((long long)ncc * taps));
    const int co = ct * kTileCo + r;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ci = cc * kKC + c * 8 + e;
      v[e] = __float2bfloat16_rn(co < Cout ? __ldg(w + ((size_t)co * Cin + ci) * taps + tap) : 0.f);
    }
    *reinterpret_cast<uint4*>(img + stage * kStageA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(v);
  }
}

// (B, C, H, W) fp32 -> channels [c_offset, c_offset + C) of (B, H, W, C_total) bf16.  A thread owns one pixel (or four
// consecutive ones) and 8 consecutive channels: eight loads, each 128 (512) contiguous bytes per warp (a lane is a pixel of one
// channel plane).  PX = 4 with 16-byte-aligned output: the block's 128 pixels x 64 channels are turned through shared memory
// (chunks XOR-swizzled, no bank conflicts either way) and a warp's store instruction writes four whole 128-byte lines; else
// every thread stores its own 16 bytes (or scalars).
template <int PX>   // pixels per thread: 4 (128-bit loads along the plane; needs HW % 4 == 0) or 1
__global__ void __launch_bounds__(256) nchw_to_nhwc_bf16_kernel(const float* __restrict__ in, int B, int C, int HW,
                                                                __nv_bfloat16* __restrict__ out, int C_total, int c_offset) {
  __shared__ __align__(16) uint4 stage[PX == 4 ? 128 * 8 : 1];   // [pixel][chunk of eight channels]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_cg = ceil_div(C, 64), n_pg = ceil_div(HW, 32 * PX);
  const long long n_items = (long long)B * n_cg * n_pg;   // block items: 64 channels x 32*PX pixels
  const bool vec_out = (C_total & 7) == 0 && (c_offset & 7) == 0 && ((uintptr_t)out & 15) == 0;
  const bool staged = PX == 4 && vec_out;
  for (long long t = blockIdx.x; t < n_items; t += gridDim.x) {
    const int pg = (int)(t % n_pg), cg = (int)((t / n_pg) % n_cg), b = (int)(t / ((long long)n_pg * n_cg));
    const int p = (pg * 32 + lane) * PX, c0 = cg * 64 + warp * 8;
    const bool live = p < HW && c0 < C;
    if (!staged && !live) continue;
    float v[8][PX];
    if (live) {
      const float* src = in + ((size_t)b * C + c0) * HW + p;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (PX == 4) {
          float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c0 + e < C) q = ld_stream_f4(reinterpret_cast<const float4*>(src + (size_t)e * HW));
          v[e][0] = q.x; v[e][PX > 1 ? 1 : 0] = q.y; v[e][PX > 2 ? 2 : 0] = q.z; v[e][PX > 3 ? 3 : 0] = q.w;
        } else {
          v[e][0] = (c0 + e < C) ? __ldg(src + (size_t)e * HW) : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < PX; ++i) {
        if (p + i >= HW) break;
        __align__(16) __nv_bfloat162 w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] = __floats2bfloat162_rn(v[2 * e][i], v[2 * e + 1][i]);
        if (staged) {
          stage[(lane * 4 + i) * 8 + (warp ^ (lane & 7))] = *reinterpret_cast<const uint4*>(w);
        } else {
          __nv_bfloat16* dst = out + ((size_t)b * HW + p + i) * C_total + c_offset + c0;
          if (vec_out && c0 + 8 <= C) {
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(w);
          } else {
            for (int e = 0; e < 8 && c0 + e < C; ++e) dst[e] = __float2bfloat16_rn(v[e][i]);
          }
        }
      }
    }
    if (staged) {
      __syncthreads();
      // 128 pixels x 8 chunks = 1024 pieces of 16 bytes: four per thread, the chunk index fastest
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = k * 256 + (int)threadIdx.x;
        const int r = idx >> 3, ch = idx & 7;
        const int pix = pg * 128 + r, c = cg * 64 + ch * 8;
        if (pix >= HW || c >= C) continue;
        const uint4 val = stage[r * 8 + (ch ^ ((r >> 2) & 7))];
        __nv_bfloat16* dst = out + ((size_t)b * HW + pix) * C_total + c_offset + c;
        if (c + 8 <= C) {
          *reinterpret_cast<uint4*>(dst) = val;
        } else {
          const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(&val);
          for (int e = 0; e < 8 && c + e < C; ++e) dst[e] = hv[e];
        }
      }
      __syncthreads();   // the stage is reused by the next item
    }
  }
}

// camera_features.mean(dim=1) (src/fusion.py:233-234) written straight as the channels-last bf16 input of camera_proj's
// first convolution: (B, n_cam, C, HW) fp32 -> channels [c_offset, c_offset + C) of (B, HW, C_total) bf16.  Same thread
// shape as the layout kernel (four pixels x eight channels, 128-bit loads along the plane), same arithmetic as
// camera_mean_vec4_kernel (sum in camera order, IEEE divide), so the result is the bf16 rounding of that kernel's output.
template <int NCAM>   // cameras known at compile time (6: the nuScenes rig): all loads of two channels in flight; 0: any number
__global__ void __launch_bounds__(256) camera_mean_nhwc_bf16_kernel(const float* __restrict__ in, int B, int n_cam, int C, int HW,
                                                                    __nv_bfloat16* __restrict__ out, int C_total, int c_offset) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_cg = ceil_div(C, 64), n_pg = ceil_div(HW, 128);
  const long long n_items = (long long)B * n_cg * n_pg;
  const float denom = (float)n_cam;
  const size_t cam_stride = (size_t)C * HW;
  for (long long t = blockIdx.x; t < n_items; t += gridDim.x) {
    const int pg = (int)(t % n_pg), cg = (int)((t / n_pg) % n_cg), b = (int)(t / ((long long)n_pg * n_cg));
    const int p = (pg * 32 + lane) * 4, c0 = cg * 64 + warp * 8;
    if (p >= HW || c0 >= C) continue;
    float v[8][4];
    const float* base = in + (((size_t)b * n_cam) * C + c0) * HW + p;
    if (NCAM > 0 && c0 + 8 <= C) {
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        float4 q[2][NCAM > 0 ? NCAM : 1];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int cam = 0; cam < NCAM; ++cam)
            q[h][cam] = ld_stream_f4(reinterpret_cast<const float4*>(base + (size_t)(e + h) * HW + cam * cam_stride));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 s4 = q[h][0];
#pragma unroll
          for (int cam = 1; cam < NCAM; ++cam) {
            s4.x = __fadd_rn(s4.x, q[h][cam].x); s4.y = __fadd_rn(s4.y, q[h][cam].y);
            s4.z = __fadd_rn(s4.z, q[h][cam].z); s4.w = __fadd_rn(s4.w, q[h][cam].w);
          }
          v[e + h][0] = __fdiv_rn(s4.x, denom); v[e + h][1] = __fdiv_rn(s4.y, denom);
          v[e + h][2] = __fdiv_rn(s4.z, denom); v[e + h][3] = __fdiv_rn(s4.w, denom);
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + e < C) {
          const float* src = base + (size_t)e * HW;
          s4 = ld_stream_f4(reinterpret_cast<const float4*>(src));
          for (int cam = 1; cam < n_cam; ++cam) {
            const float4 q = ld_stream_f4(reinterpret_cast<const float4*>(src + cam * cam_stride));
            s4.x = __fadd_rn(s4.x, q.x); s4.y = __fadd_rn(s4.y, q.y); s4.z = __fadd_rn(s4.z, q.z); s4.w = __fadd_rn(s4.w, q.w);
          }
        }
        v[e][0] = __fdiv_rn(s4.x, denom); v[e][1] = __fdiv_rn(s4.y, denom); v[e][2] = __fdiv_rn(s4.z, denom); v[e][3] = __fdiv_rn(s4.w, denom);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat16* dst = out + ((size_t)b * HW + p + i) * C_total + c_offset + c0;
      if (c0 + 8 <= C) {
        __align__(16) __nv_bfloat162 w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] = __floats2bfloat162_rn(v[2 * e][i], v[2 * e + 1][i]);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(w);
      } else {
        for (int e = 0; e < 8 && c0 + e < C; ++e) dst[e] = __float2bfloat16_rn(v[e][i]);
      }
    }
  }
}

// ---- fp32-accuracy mode: weight image, input layout, max|x| ------------------------------------------------------------------
// Image: stages [co tile][tap][3 x (Cin/64)] of [128 co][64 ci] fp16 — per 64-channel chunk the triple (w_hi, w_hi, w_lo), in
// the order the kernels consume them — then unscale[co] (padded to whole co tiles): the inverse of the power of two that
// brought max |w[co]| into [1, 2).
__global__ void __launch_bounds__(256) conv_split_scale_kernel(const float* __restrict__ w, int Cout, int per_co, float* __restrict__ unscale,
                                                               int n_slots) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_slots) return;
  float m = 0.f;
  if (warp < Cout)
    for (int i = lane; i < per_co; i += 32) m = fmaxf(m, fabsf(__ldg(w + (size_t)warp * per_co + i)));
#pragma unroll
  for (int d = 16; d; d >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, d));
  if (lane == 0) {
    int ex = 0;
    float inv = 1.f;
    if (m > 0.f && isfinite(m)) {
      frexpf(m, &ex);
      inv = ldexpf(1.f, ex - 1);
    }
    unscale[warp] = inv;
  }
}
__global__ void __launch_bounds__(256) conv_pack_split_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, uint8_t* __restrict__ img,
                                                              const float* __restrict__ unscale, long long n_chunks) {
  const int ncc = Cin / kKC, ncc3 = 3 * ncc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_chunks; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 7);
    const int r = (int)((i >> 3) & 127);
    const long long stage = i >> 10;
    const int j = (int)(stage % ncc3);
    const int tap = (int)((stage / ncc3) % taps);
    const int ct = (int)(stage / ((long long)ncc3 * taps));
    const int co = ct * kTileCo + r, term = j / ncc, cc = j % ncc;
    const float sc = co < Cout ? 1.f / unscale[co] : 0.f;
    __align__(16) __half v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ci = cc * kKC + c * 8 + e;
      const float x = co < Cout ? __ldg(w + ((size_t)co * Cin + ci) * taps + tap) * sc : 0.f;
      const __half hi = __float2half_rn(x);
      v[e] = term < 2 ? hi : __float2half_rn(x - __half2float(hi));
    }
    *reinterpret_cast<uint4*>(img + stage * kStageA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(v);
  }
}

// max |x| of a tensor as float bits (non-negative floats order like unsigned integers): atomicMax into *stat
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n, uint32_t* __restrict__ stat) {
  float m = 0.f;
  const long long n4 = (((uintptr_t)x & 15) == 0) ? n / 4 : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 q = ld_stream_f4(reinterpret_cast<const float4*>(x) + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fmaxf(fabsf(q.z), fabsf(q.w))));
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
  for (int d = 16; d; d >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(stat, __float_as_uint(m));
}

// (B, C, H, W) fp32 -> (B, H, W, [hi C_total | lo C_total]) fp16, channels [c_offset, c_offset + C) of both halves, scaled by
// activation_scale(*stat).  Thread shape of nchw_to_nhwc_bf16_kernel on the way in: four pixels x eight channels, 128-bit loads
// along the plane.  PX = 4: the CTA's 128 pixels x 64 channels are turned through shared memory (16-byte chunks XOR-swizzled: no
// bank conflicts either way) so that a warp's store instruction writes four whole 128-byte lines — 64 channels of a pixel and
// half — instead of 32 pieces of 16 bytes in 32 different lines (the 374 MB camera tensors took 380 us that way: 2 TB/s).
template <int PX>
__global__ void __launch_bounds__(256) nchw_to_nhwc_split_kernel(const float* __restrict__ in, int B, int C, int HW, __half* __restrict__ out,
                                                                 int C_total, int c_offset, const uint32_t* __restrict__ stat) {
  __shared__ __align__(16) uint4 stage[PX == 4 ? 2 * 128 * 8 : 1];   // [hi | lo][pixel][chunk of eight channels]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_cg = ceil_div(C, 64), n_pg = ceil_div(HW, 32 * PX);
  const long long n_items = (long long)B * n_cg * n_pg;
  const float S = activation_scale(__ldg(stat));
  for (long long t = blockIdx.x; t < n_items; t += gridDim.x) {
    const int pg = (int)(t % n_pg), cg = (int)((t / n_pg) % n_cg), b = (int)(t / ((long long)n_pg * n_cg));
    const int p = (pg * 32 + lane) * PX, c0 = cg * 64 + warp * 8;
    const bool live = p < HW && c0 < C;
    if (PX != 4 && !live) continue;
    float v[8][PX];
    if (live) {
      const float* src = in + ((size_t)b * C + c0) * HW + p;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (PX == 4) {
          float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c0 + e < C) q = ld_stream_f4(reinterpret_cast<const float4*>(src + (size_t)e * HW));
          v[e][0] = q.x; v[e][PX > 1 ? 1 : 0] = q.y; v[e][PX > 2 ? 2 : 0] = q.z; v[e][PX > 3 ? 3 : 0] = q.w;
        } else {
          v[e][0] = (c0 + e < C) ? __ldg(src + (size_t)e * HW) : 0.f;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < PX; ++i) {
      if (!live || p + i >= HW) break;
      __align__(16) __half hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float x = v[e][i] * S;
        hi[e] = __float2half_rn(x);
        lo[e] = __float2half_rn(x - __half2float(hi[e]));
      }
      if (PX == 4) {
        const int r = lane * 4 + i, ch = warp ^ (lane & 7);
        stage[r * 8 + ch] = *reinterpret_cast<const uint4*>(hi);
        stage[(128 + r) * 8 + ch] = *reinterpret_cast<const uint4*>(lo);
      } else {
        __half* dst = out + ((size_t)b * HW + p + i) * (2 * (size_t)C_total) + c_offset + c0;
        if (c0 + 8 <= C) {
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
          *reinterpret_cast<uint4*>(dst + C_total) = *reinterpret_cast<const uint4*>(lo);
        } else {
          for (int e = 0; e < 8 && c0 + e < C; ++e) {
            dst[e] = hi[e];
            dst[C_total + e] = lo[e];
          }
        }
      }
    }
    if (PX == 4) {
      __syncthreads();
      // 2 halves x 128 pixels x 8 chunks = 2048 pieces of 16 bytes: eight per thread, the chunk index fastest
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int idx = k * 256 + (int)threadIdx.x;
        const int half = idx >> 10, r = (idx >> 3) & 127, ch = idx & 7;
        const int pix = pg * 128 + r, c = cg * 64 + ch * 8;
        if (pix >= HW || c >= C) continue;
        const uint4 val = stage[(half * 128 + r) * 8 + (ch ^ ((r >> 2) & 7))];
        __half* dst = out + ((size_t)b * HW + pix) * (2 * (size_t)C_total) + (size_t)half * C_total + c_offset + c;
        if (c + 8 <= C) {
          *reinterpret_cast<uint4*>(dst) = val;
        } else {
          const __half* hv = reinterpret_cast<const __half*>(&val);
          for (int e = 0; e < 8 && c + e < C; ++e) dst[e] = hv[e];
        }
      }
      __syncthreads();   // the stage is reused by the next item
    }
  }
}

// A stack of k 3x3 (padding 1) convolutions applied to a spatially CONSTANT image (the radar branch: a (B,C) vector
// broadcast to (B,C,H,W), src/fusion.py:277-281) has only (2k+1)^2 distinct output pixels per channel: what a pixel sees
// depends on how many of the k rows / columns towards each border exist.  The stack therefore runs on a (2k+1) x (2k+1)
// image and this kernel spreads the result: out[y][x] = small[cls(y)][cls(x)], cls(i) = i for i < k, s-1-(n-1-i) for
// i >= n-k, k otherwise (s = 2k+1).  Writes NCHW fp32 and/or a channel slice of a channels-last bf16 tensor.
__device__ __forceinline__ int border_class(int i, int n, int s) {
  const int k = s >> 1;
  return i < k ? i : (i >= n - k ? s - (n - i) : k);
}
__global__ void __launch_bounds__(256) border_expand_kernel(const float* __restrict__ small, int B, int C, int s, int H, int W,
                                                            float* __restrict__ out_nchw, __nv_bfloat16* __restrict__ out_nhwc,
                                                            int C_total, int c_offset) {
  const long long HW = (long long)H * W;
  if (out_nhwc) {
    const int cg = (C + 7) / 8;
    const long long n = (long long)B * HW * cg;
    const bool vec_ok = (C_total & 7) == 0 && (c_offset & 7) == 0 && (reinterpret_cast<uintptr_t>(out_nhwc) & 15) == 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const int g = (int)(i % cg);
      const long long bp = i / cg;
      const int p = (int)(bp % HW), b = (int)(bp / HW);
      const int cy = border_class(p / W, H, s), cx = border_class(p % W, W, s);
      const float* src = small + (((size_t)b * C + g * 8) * s + cy) * s + cx;
      __nv_bfloat16* dst = out_nhwc + ((size_t)b * HW + p) * C_total + c_offset + g * 8;
      if (vec_ok && g * 8 + 8 <= C) {     // eight channels of a pixel: one 16-byte store
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16_rn(__ldg(src + (size_t)e * s * s));
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
      } else {
        for (int e = 0; e < 8 && g * 8 + e < C; ++e) dst[e] = __float2bfloat16_rn(__ldg(src + (size_t)e * s * s));
      }
    }
  }
  if (out_nchw) {
    const long long n = (long long)B * C * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const int p = (int)(i % HW);
      const long long bc = i / HW;
      out_nchw[i] = __ldg(small + (bc * s + border_class(p / W, H, s)) * s + border_class(p % W, W, s));
    }
  }
}

}  // namespace
}  // namespace b200bev

using namespace b200bev;

extern "C" B200BEV_API int b200bev_border_expand(const float* small, int B, int C, int s, int H, int W, float* out_nchw, void* out_nhwc,
                                     int C_total, int c_offset, void* stream) {
  if (!small || (!out_nchw && !out_nhwc) || B <= 0 || C <= 0 || s < 1 || !(s & 1) || H < s || W < s) return B200BEV_ERR_INVALID_ARGUMENT;
  if (out_nhwc && (c_offset < 0 || c_offset + C > C_total)) return B200BEV_ERR_INVALID_ARGUMENT;
  const long long work = (long long)B * C * H * W / (out_nhwc && !out_nchw ? 8 : 1);
  long long blocks = (work + 255) / 256;
  if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
  border_expand_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(small, B, C, s, H, W, out_nchw, (__nv_bfloat16*)out_nhwc, C_total,
                                                                       c_offset);
  return launch_status();
}

extern "C" B200BEV_API size_t b200bev_conv_pack_bytes(int Cout, int Cin, int taps) {
  if (Cout <= 0 || Cin <= 0 || Cin % kKC != 0 || (taps != 1 && taps != 9)) return 0;
  return (size_t)ceil_div(Cout, kTileCo) * taps * (Cin / kKC) * kStageA;
}

extern "C" B200BEV_API int b200bev_conv_pack_bf16(const float* weight, int Cout, int Cin, int taps, void* image, size_t image_bytes,
                                      void* stream) {
  const size_t need = b200bev_conv_pack_bytes(Cout, Cin, taps);
  if (!weight || !image) return B200BEV_ERR_INVALID_ARGUMENT;
  if (need == 0) return B200BEV_ERR_UNSUPPORTED;
  if (image_bytes < need || ((uintptr_t)image & 15)) return B200BEV_ERR_WORKSPACE;
  const long long n_chunks = (long long)(need / 16);
  long long blocks = (n_chunks + 255) / 256;
  if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
  conv_pack_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(weight, Cout, Cin, taps, (uint8_t*)image, n_chunks);
  return launch_status();
}

extern "C" B200BEV_API int b200bev_camera_mean_nhwc_bf16(const float* feats, int B, int n_cam, int C, int H, int W, void* out_nhwc,
                                             int C_total, int c_offset, void* stream) {
  if (!feats || !out_nhwc || B <= 0 || n_cam <= 0 || C <= 0 || H <= 0 || W <= 0 || c_offset < 0 || c_offset + C > C_total)
    return B200BEV_ERR_INVALID_ARGUMENT;
  // 128-bit loads along the plane and 128-bit stores of 8 channels
  if (((H * W) & 3) || ((uintptr_t)feats & 15) || (C_total & 7) || (c_offset & 7) || ((uintptr_t)out_nhwc & 15)) return B200BEV_ERR_UNSUPPORTED;
  const long long tiles = (long long)B * ceil_div(C, 64) * ceil_div(H * W, 128);
  long long blocks = tiles < (long long)sm_count() * 16 ? tiles : (long long)sm_count() * 16;
  if (n_cam == 6)
    camera_mean_nhwc_bf16_kernel<6><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(feats, B, n_cam, C, H * W, (__nv_bfloat16*)out_nhwc,
                                                                                C_total, c_offset);
  else
    camera_mean_nhwc_bf16_kernel<0><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(feats, B, n_cam, C, H * W, (__nv_bfloat16*)out_nhwc,
                                                                                C_total, c_offset);
  return launch_status();
}

extern "C" B200BEV_API int b200bev_nchw_to_nhwc_bf16(const float* in, int B, int C, int H, int W, void* out_nhwc, int C_total,
                                         int c_offset, void* stream) {
  if (!in || !out_nhwc || B <= 0 || C <= 0 || H <= 0 || W <= 0 || c_offset < 0 || c_offset + C > C_total)
    return B200BEV_ERR_INVALID_ARGUMENT;
  const bool quad = ((H * W) & 3) == 0 && ((uintptr_t)in & 15) == 0;   // four pixels per thread, 128-bit loads
  const long long tiles = (long long)B * ceil_div(C, 64) * ceil_div(H * W, quad ? 128 : 32);
  long long blocks = tiles < (long long)sm_count() * 16 ? tiles : (long long)sm_count() * 16;
  if (quad)
    nchw_to_nhwc_bf16_kernel<4><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, B, C, H * W, (__nv_bfloat16*)out_nhwc, C_total, c_offset);
  else
    nchw_to_nhwc_bf16_kernel<1><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, B, C, H * W, (__nv_bfloat16*)out_nhwc, C_total, c_offset);
  return launch_status();
}

namespace {
int launch_conv(const void* x_nhwc, int B, int H, int W, int Cin, const void* weight_image, const float* bias, int Cout, int taps,
                int relu, float* out_nchw, void* out_nhwc, int out_ct, int out_coff, void* stream, const uint32_t* x_stat = nullptr) {
  if (!x_nhwc || !weight_image || (!out_nchw && !out_nhwc) || B <= 0 || H <= 0 || W <= 0 || Cout <= 0) return B200BEV_ERR_INVALID_ARGUMENT;
  if (Cin <= 0 || Cin % kKC != 0 || (taps != 1 && taps != 9)) return B200BEV_ERR_UNSUPPORTED;
  if ((long long)B * H * W * Cin >= (1ll << 40)) return B200BEV_ERR_UNSUPPORTED;
  if (((uintptr_t)x_nhwc | (uintptr_t)weight_image) & 15) return B200BEV_ERR_INVALID_ARGUMENT;
  if (out_nhwc && (out_coff < 0 || out_coff + Cout > out_ct)) return B200BEV_ERR_INVALID_ARGUMENT;
  ConvArgs a{(const __nv_bfloat16*)x_nhwc, (const uint8_t*)weight_image, bias, out_nchw, B, H, W, Cin, Cout, taps, relu,
             (__nv_bfloat16*)out_nhwc, out_ct, out_coff, 0, Cin, nullptr, nullptr};
  if (x_stat) {   // fp32-accuracy mode: split fp16 operands, the unscale array sits behind the stages
    a.split = 1;
    a.x_pitch = 2 * Cin;
    a.unscale = reinterpret_cast<const float*>((const uint8_t*)weight_image + (size_t)ceil_div(Cout, kTileCo) * taps * 3 * (Cin / kKC) * kStageA);
    a.x_stat = x_stat;
  }
  const long long tiles = (((long long)B * H * W + kTilePx - 1) / kTilePx) * ceil_div(Cout, kTileCo);
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  const char* impl = debug_env("B200BEV_CONV_IMPL");
  HaloGeom geo;
  const bool want_halo = !(impl && impl[0] == 'p');   // "per-tap": the per-tap kernel for 3x3 too (A/B timing)
  if (taps == 9 && want_halo && halo_geometry(H, W, &geo, kTilePx)) {
    const int n_co = ceil_div(Cout, kTileCo);
    const long long htiles = (long long)B * geo.tiles_per_frame * n_co;
    // CTA pairs (cta_group::2) when the output channels come in pairs of co tiles and a block has enough k chunks to pay for
    // the pair's hand-overs (measured, 32 frames: 768->512 358 -> 338 us, 512->512 at 57 x 100 545 -> 514 us — both then AT the
    // measured bf16 burst peak —, 512->256 141 -> 137, 256->256 80 -> 80, 128->256 51 -> 55)
    // (full 256-column tiles only: the narrower tiles of small maps stay on the single-CTA form, the one they were tested on)
    const bool pair_ok = n_co % 2 == 0 && geo.N == kTilePx && Cin >= 256 && !(impl && impl[0] == '1');
    if (pair_ok) {
      B200BEV_CUDA_TRY(cudaFuncSetAttribute(conv3x3_tc_halo_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmem));
      cudaLaunchConfig_t cfg{};
      const long long pair_tiles = htiles / 2;
      cfg.gridDim = dim3((unsigned)(2 * (pair_tiles < sm_count() / 2 ? pair_tiles : sm_count() / 2)));
      cfg.blockDim = dim3(kHaloThreads);
      cfg.dynamicSmemBytes = kHaloSmem;
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      B200BEV_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_tc_halo_kernel<2>, a, geo, 0));
      return launch_status();
    }
    B200BEV_CUDA_TRY(cudaFuncSetAttribute(conv3x3_tc_halo_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmem));
    const int hgrid = (int)(htiles < sm_count() ? htiles : sm_count());
    conv3x3_tc_halo_kernel<1><<<hgrid, kHaloThreads, kHaloSmem, (cudaStream_t)stream>>>(a, geo, 0);
    return launch_status();
  }
  // 1x1: the TMA-fed kernel (one tensor load per stage instead of 2,048 cp.async)
  if (taps == 1 && (long long)B * H * W < (1ll << 31) && !(impl && impl[0] == 'c')) {   // "c": the cp.async kernel (A/B timing)
    CUtensorMap xmap;
    if (make_pixel_map(&xmap, x_nhwc, (long long)B * H * W, a.x_pitch, a.split != 0)) {
      B200BEV_CUDA_TRY(cudaFuncSetAttribute(conv1x1_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmem));
      conv1x1_tma_kernel<<<grid, kTmaThreads, kConvSmem, (cudaStream_t)stream>>>(a, xmap);
      return launch_status();
    }
  }
  // per-tap kernel (1x1, and 3x3 on maps too wide for the shared pixel block), optionally as CTA pairs.
  // Measured: correct (parity green) and SLOWER — camera_proj.3 (1x1, 512 -> 256, 32 x 57 x 100) 94 -> 152 us: this kernel is
  // paced by the turn-around of its four ring slots, not by shared-memory bandwidth, and a pair adds a hop to every stage.  The
  // pair instantiation stays in the source behind the debug switch B200BEV_CONV_IMPL=2.
  if (ceil_div(Cout, kTileCo) % 2 == 0 && Cin >= 256 && impl && impl[0] == '2') {
    B200BEV_CUDA_TRY(cudaFuncSetAttribute(conv_tc_ws_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmem));
    const long long pair_tiles = tiles / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * (pair_tiles < sm_count() / 2 ? pair_tiles : sm_count() / 2)));
    cfg.blockDim = dim3(kWsThreads);
    cfg.dynamicSmemBytes = kConvSmem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200BEV_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_tc_ws_kernel<2>, a));
    return launch_status();
  }
  B200BEV_CUDA_TRY(cudaFuncSetAttribute(conv_tc_ws_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmem));
  conv_tc_ws_kernel<1><<<grid, kWsThreads, kConvSmem, (cudaStream_t)stream>>>(a);
  return launch_status();
}
}  // namespace

extern "C" B200BEV_API int b200bev_conv_bn_relu_bf16(const void* x_nhwc, int B, int H, int W, int Cin, const void* weight_image,
                                         const float* bias, int Cout, int taps, int relu, float* out_nchw, void* stream) {
  if (!out_nchw) return B200BEV_ERR_INVALID_ARGUMENT;
  return launch_conv(x_nhwc, B, H, W, Cin, weight_image, bias, Cout, taps, relu, out_nchw, nullptr, 0, 0, stream);
}

extern "C" B200BEV_API int b200bev_conv_bn_relu_bf16_nhwc(const void* x_nhwc, int B, int H, int W, int Cin, const void* weight_image,
                                              const float* bias, int Cout, int taps, int relu, void* out_nhwc, int out_c_total,
                                              int out_c_offset, float* out_nchw, void* stream) {
  if (!out_nhwc) return B200BEV_ERR_INVALID_ARGUMENT;
  return launch_conv(x_nhwc, B, H, W, Cin, weight_image, bias, Cout, taps, relu, out_nchw, out_nhwc, out_c_total, out_c_offset, stream);
}

// ---- fp32-accuracy convolution blocks (three fp16 products per fp32 product) -------------------------------------------------
extern "C" B200BEV_API size_t b200bev_conv_pack_split_bytes(int Cout, int Cin, int taps) {
  if (Cout <= 0 || Cin <= 0 || Cin % kKC != 0 || (taps != 1 && taps != 9)) return 0;
  const size_t tiles = (size_t)ceil_div(Cout, kTileCo);
  return tiles * taps * 3 * (Cin / kKC) * kStageA + tiles * kTileCo * sizeof(float);
}

extern "C" B200BEV_API int b200bev_conv_pack_split(const float* weight, int Cout, int Cin, int taps, void* image, size_t image_bytes,
                                       void* stream) {
  const size_t need = b200bev_conv_pack_split_bytes(Cout, Cin, taps);
  if (!weight || !image) return B200BEV_ERR_INVALID_ARGUMENT;
  if (need == 0) return B200BEV_ERR_UNSUPPORTED;
  if (image_bytes < need || ((uintptr_t)image & 15)) return B200BEV_ERR_WORKSPACE;
  const int n_slots = ceil_div(Cout, kTileCo) * kTileCo;
  const size_t stage_bytes = need - (size_t)n_slots * sizeof(float);
  float* unscale = reinterpret_cast<float*>((uint8_t*)image + stage_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  conv_split_scale_kernel<<<(n_slots * 32 + 255) / 256, 256, 0, st>>>(weight, Cout, Cin * taps, unscale, n_slots);
  B200BEV_CUDA_TRY(cudaGetLastError());
  const long long n_chunks = (long long)(stage_bytes / 16);
  long long blocks = (n_chunks + 255) / 256;
  if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
  conv_pack_split_kernel<<<(int)blocks, 256, 0, st>>>(weight, Cout, Cin, taps, (uint8_t*)image, unscale, n_chunks);
  return launch_status();
}

extern "C" B200BEV_API int b200bev_absmax(const float* x, int64_t n, void* stat, void* stream) {
  if (!x || !stat || n <= 0) return B200BEV_ERR_INVALID_ARGUMENT;
  long long blocks = (n / 4 + 255) / 256 + 1;
  if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
  absmax_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, (long long)n, (uint32_t*)stat);
  return launch_status();
}

extern "C" B200BEV_API int b200bev_nchw_to_nhwc_split(const float* in, int B, int C, int H, int W, void* out_nhwc, int C_total,
                                          int c_offset, const void* stat, void* stream) {
  if (!in || !out_nhwc || !stat || B <= 0 || C <= 0 || H <= 0 || W <= 0 || c_offset < 0 || c_offset + C > C_total)
    return B200BEV_ERR_INVALID_ARGUMENT;
  if ((C_total & 7) || (c_offset & 7) || ((uintptr_t)out_nhwc & 15)) return B200BEV_ERR_UNSUPPORTED;
  const bool quad = ((H * W) & 3) == 0 && ((uintptr_t)in & 15) == 0;
  const long long tiles = (long long)B * ceil_div(C, 64) * ceil_div(H * W, quad ? 128 : 32);
  long long blocks = tiles < (long long)sm_count() * 16 ? tiles : (long long)sm_count() * 16;
  if (quad)
    nchw_to_nhwc_split_kernel<4><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, B, C, H * W, (__half*)out_nhwc, C_total, c_offset,
                                                                              (const uint32_t*)stat);
  else
    nchw_to_nhwc_split_kernel<1><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, B, C, H * W, (__half*)out_nhwc, C_total, c_offset,
                                                                              (const uint32_t*)stat);
  return launch_status();
}

extern "C" B200BEV_API int b200bev_conv_bn_relu_split(const void* x_split, const void* x_stat, int B, int H, int W, int Cin,
                                          const void* weight_image, const float* bias, int Cout, int taps, int relu, float* out_nchw,
                                          void* stream) {
  if (!out_nchw || !x_stat) return B200BEV_ERR_INVALID_ARGUMENT;
  return launch_conv(x_split, B, H, W, Cin, weight_image, bias, Cout, taps, relu, out_nchw, nullptr, 0, 0, stream, (const uint32_t*)x_stat);
}
