// S1b — tensor-core path of the fused PointNet shared-MLP + max: tcgen05.mma with the activations
// resident in TENSOR MEMORY for the whole layer chain.
//
// Reference arithmetic: PointNetLiDAREncoder.forward, src/encoders.py:289-298 (eval mode, BatchNorm
// folded), for the layer widths C-64-128-256-512-1024 of configs/base.yaml:176-185.  bf16 operands,
// fp32 accumulation; parity bound 1e-2 of max|ref| (north_star).
//
// One persistent CTA per SM, tile = 128 points = the 128 TMEM lanes.  D[point][channel] = A.B^T with
//   A = activations, [128 points][K] bf16, living in TMEM (two bf16 per 32-bit column)
//   B = weights,     [128 channels][64 k] bf16 stages in shared memory (K-major, 128-byte swizzle)
//   D = fp32 accumulators in TMEM
// so an activation never touches shared memory or HBM: the epilogue warps read D with tcgen05.ld, add
// the bias, apply ReLU, round to bf16 and write the next layer's A operand back with tcgen05.st.
// Shared memory is left entirely to a 12-stage ring of 16 KB weight stages that one producer thread
// streams with cp.async.bulk (the 1.39 MB weight image is pre-tiled and pre-swizzled once by
// b200bev_pointnet_pack_bf16, so a stage is one contiguous bulk copy; it stays L2-resident).
//
//   warp 0      producer: cp.async.bulk global -> smem ring, mbarrier expect_tx / complete_tx
//   warp 1      MMA issuer: one thread, 4 x tcgen05.mma (128x128x16) per stage, tcgen05.commit
//               releases the stage and publishes the accumulator
//   warps 2..17 epilogue (4 warps per TMEM lane quadrant, each owning a quarter of the columns):
//               layer 1 (K = 4) on CUDA cores, layers 2-4 TMEM -> bias/ReLU/bf16 -> TMEM,
//               layer 5: max over the 128 points of the tile.  Points are TMEM lanes, so this is a
//               cross-lane reduction: a 31-shuffle transposing butterfly per 32 channels leaves lane l
//               with channel l's maximum; it is folded into 32 running-max registers per thread and
//               flushed once per frame (bias + ReLU after the max: both are monotone per channel).
//
// TMEM columns (512): [0,256) act4 (act1 at [0,32) and act2 at [64,128) reuse it earlier in the
// tile), [256,384) act3, which becomes accumulator 0 for layer 5, [384,512) accumulator 1.
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

#ifndef B200BEV_CELL_MMA_UNROLL
#define B200BEV_CELL_MMA_UNROLL 0   // 1: the MMA issuer's chunk loop unrolled in cell mode too (experiments)
#endif
#ifndef B200BEV_TC_TRIP
#define B200BEV_TC_TRIP 0   // 0: per-mode default; 1 or 2 forces the ring pairs per MMA trip (experiments)
#endif

namespace b200bev {
namespace {

constexpr int kTileM = 128;
constexpr int kStageBytes = 16384;  // 128 rows x 64 bf16
constexpr int kStagesGlobal = 12;   // ring depth when shared memory holds nothing but weights
constexpr int kStagesCell = 8;      // cell mode gives 72 KB to the per-warp transposing tiles below
constexpr int kStagesCellPair = 8;  // a CTA pair keeps half of every stage: the same bytes carry twice the ring pairs
constexpr int kTStride = 36;        // floats per channel row of a warp's private 32 x 32 tile: 16-B aligned rows, and both the
                                    // lane = point stores and the lane = channel 128-bit loads are free of bank conflicts
constexpr int kWarpTile = 32 * kTStride;
constexpr int kEpiThreads = 512;     // 16 epilogue warps: 4 per TMEM lane quadrant
constexpr int kTcThreads = 64 + kEpiThreads;
constexpr int kPairBytes = 2 * kStageBytes;
// Weight stream of one tile, in PAIRS of 16 KB stages (one 32 KB bulk copy, one full/empty barrier each):
// layer 2 has a single stage, so its pair carries a padding stage that nothing reads.
constexpr int kPairsPerTile = 1 + 2 + 8 + 32;
constexpr int kStagesPerTile = 2 * kPairsPerTile;   // 86: 85 weight stages + 1 padding stage
constexpr int kBiasFloats = 128 + 256 + 512 + 1024;
constexpr int kMaxCin = 16;

constexpr uint32_t kColAct1 = 0, kColAct2 = 64, kColAct4 = 0, kColAct3 = 256, kColAcc0 = 256, kColAcc1 = 384;
// third accumulator: the upper half of the act4 region, free until layer 4 writes its chunks 2 and 3
constexpr uint32_t kColAccX = 128;
constexpr int kEpiWarps = kEpiThreads / 32;

// Accumulator of chunk c of network layer 2 + `layer` (0..3); the MMA issuer and the epilogue must agree.
//   layer 2: acc1.  layer 3: acc1, accX.  layer 4: acc1, accX, acc1, acc1 (accX is overwritten by the
//   activations of chunks 2 and 3).  layer 5: acc0 / acc1 alternating (act3, which acc0 aliases, is dead).
__device__ __forceinline__ int acc_buffer(int layer, int c) {
  if (layer == 3) return c & 1;
  if (layer == 0) return 1;
  if (layer == 1) return c == 0 ? 1 : 2;
  return c == 1 ? 2 : 1;
}
__device__ __forceinline__ uint32_t acc_column(int buf) { return buf == 0 ? kColAcc0 : buf == 1 ? kColAcc1 : kColAccX; }

struct TcArgs {
  const float* pts;
  int B, N, C;
  const uint8_t* tc;  // [85 stages][W1^T (C x 64) f32][b1 64][bias L2..L5]
  float* out_global;  // (B,1024) or nullptr
  // cell mode
  const int32_t* perm;     // (B,N) points ordered by cell
  const int32_t* offsets;  // (B,n_cells+1)
  int n_cells;
  float* out_canvas;       // (B,n_cells,1024)
  int tiles_per_frame;
  long long total_tiles;
  int cluster;             // 1: one CTA per tile stream (cta_group::1); 2: CTA pairs (cta_group::2), each CTA holding half of every weight stage
  unsigned long long* trace;  // debug only (B200BEV_TC_TRACE): clock stamps of CTA 0, else nullptr
  int debug;                  // debug only (B200BEV_TC_DEBUG): bit 0 = skip the weight copies after the first tile,
                              // bit 1 = skip the layer-5 epilogue, bit 2 = issue MMAs without waiting for weights
};

// Debug timeline: (event id << 48 | clock) appended by one thread per role of CTA 0.
constexpr int kTraceMma = 0, kTraceEpi = 8192, kTraceLen = 16384;
template <bool TRACE>
__device__ __forceinline__ void trace_ev(const TcArgs& a, int& idx, int base, unsigned id) {
  if (TRACE && a.trace != nullptr && blockIdx.x == 0 && idx < 8192) {
    a.trace[base + idx] = ((unsigned long long)id << 48) | ((unsigned long long)clock64() & 0xffffffffffffull);
    ++idx;
  }
}

constexpr int kImageStages = kStagesPerTile;
// blob: [single-CTA stage image][fp32 tail: W1^T, b1, biases][pair image: rank 0 stream, rank 1 stream]
// The pair image holds, per CTA of a cta_group::2 pair, the 64 of every stage's 128 channel rows that CTA feeds to the
// MMA (rank r: rows 64r .. 64r+63), stage after stage, so a ring pair is one contiguous 16 KB copy per CTA.
__host__ __device__ inline size_t tc_pair_image_offset(int C) {
  const size_t tail_end = (size_t)kImageStages * kStageBytes + ((size_t)C * 64 + 64 + kBiasFloats) * sizeof(float);
  return (tail_end + 1023) & ~(size_t)1023;
}
__host__ __device__ inline size_t tc_blob_bytes(int C) { return tc_pair_image_offset(C) + (size_t)kImageStages * kStageBytes; }

// ---- PTX wrappers -------------------------------------------------------------------------------
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive2(uint64_t* bar) {   // two arrivals at once
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], 2;" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// tcgen05.commit of a CTA pair's MMAs, arriving on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// Same, without release semantics: for the relay, which forwards the completion of a bulk copy (the copy's bytes are
// in shared memory before its mbarrier completes; the relay itself has written nothing that needs publishing)
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive2_remote_relaxed(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra], 2;\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// One lane of a fully converged warp (elect.sync): the branch stays warp-uniform for the compiler.
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T, 128 x 128 x 16 per CTA, bf16 -> f32.  CG = 2: one instruction of the leader CTA
// drives both CTAs of the pair (M = 256), each with its own A and D and half of B (tests/cuda/umma2_probe.cu).
template <int CG>
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
        "r"(a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
        "r"(a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);  // start address / 16
  d |= (uint64_t)1 << 16;                   // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                   // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}

#define TC_LD32(r, taddr)                                                                                                  \
  asm volatile(                                                                                                            \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"      \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                           \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),        \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),             \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),            \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                          \
      : "r"(taddr)                                                                                                         \
      : "memory")

#define TC_ST16(taddr, r, o)                                                                                               \
  asm volatile(                                                                                                            \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), \
      "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]),             \
      "r"(r[o + 7]), "r"(r[o + 8]), "r"(r[o + 9]), "r"(r[o + 10]), "r"(r[o + 11]), "r"(r[o + 12]), "r"(r[o + 13]),         \
      "r"(r[o + 14]), "r"(r[o + 15])                                                                                       \
      : "memory")

#define TC_ST8(taddr, r)                                                                                      \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), \
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])                    \
               : "memory")

// relu, round to bf16 and pack in one instruction (F2FP.RELU); element with the even k index in the low half
__device__ __forceinline__ uint32_t relu_pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// two fp32 additions / fused multiply-adds per instruction (FADD2 / FFMA2 of sm_100): the same IEEE results as the scalar forms
__device__ __forceinline__ float2 add_f32x2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fma_f32x2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  // element with the even k index in the low half (verified on hardware by tests/cuda/umma_probe.cu)
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// One butterfly level of the transposing max-reduction: halves the live values, doubles the lanes
// each value covers.
template <int HALF>
__device__ __forceinline__ void butterfly_level(float* v, int lane) {
  const bool up = (lane & HALF) != 0;
#pragma unroll
  for (int j = 0; j < HALF; ++j) {
    const float send = up ? v[j] : v[j + HALF];
    const float keep = up ? v[j + HALF] : v[j];
    v[j] = fmaxf(keep, __shfl_xor_sync(FULL_MASK, send, HALF));
  }
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

template <bool CELL, bool TRACE, int CG>
__global__ void __launch_bounds__(kTcThreads, 1) pointnet_mlp_tc_kernel(TcArgs a) {
  constexpr int kStages = CELL ? (CG == 2 ? kStagesCellPair : kStagesCell) : kStagesGlobal;
  constexpr int kSlotBytes = kPairBytes / CG;                    // bytes of one ring pair in THIS CTA's shared memory
  constexpr int kRingPairs = kStages * kStageBytes / kSlotBytes;  // a pair of CTAs keeps twice as many pairs in flight
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // SWIZZLE_128B: 1024-B aligned tiles
  float* tile_s = reinterpret_cast<float*>(ring + (size_t)kStages * kStageBytes);   // CELL: one [32 channels][kTStride] tile per epilogue warp
  int* cid_s = reinterpret_cast<int*>(tile_s + (CELL ? kEpiWarps * kWarpTile : 0));   // CELL: cell id of each tile slot
  uint32_t* endmask_s = reinterpret_cast<uint32_t*>(cid_s + (CELL ? 128 : 0));      // CELL: run-end flags, one word per warp
  int* scan_s = reinterpret_cast<int*>(endmask_s + (CELL ? 4 : 0));                 // CELL: [0,4) warp maxima, [4,6) "more cells" flags
  float* bias_s = reinterpret_cast<float*>(scan_s + (CELL ? 8 : 0));                // L2 | L3 | L4 | L5
  float* w1_s = bias_s + kBiasFloats;                                               // W1^T (C x 64), then b1 (64)
  uint64_t* bars = reinterpret_cast<uint64_t*>(w1_s + kMaxCin * 64 + 64);
  uint64_t* full = bars;                    // [kRingPairs used] this CTA's weights of a ring pair landed
  uint64_t* empty = bars + kStages;         // [kRingPairs used] MMAs that read the pair retired
  uint64_t* peer_full = empty + kStages;    // [kRingPairs used] CG = 2, leader only: the follower's half landed too
  uint64_t* acc_full = peer_full + kStages; // [3]        accumulator complete
  uint64_t* acc_empty = acc_full + 3;       // [3]        accumulator drained by the epilogue (one arrival per warp)
  uint64_t* act_ready = acc_empty + 3;      // [4]        K-pair kp (128 channels) of the next A operand is in TMEM (one arrival per warp)
  uint64_t* accx = act_ready + 4;           // [1]        this CTA's warps have drained accX (layer 4, chunk 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accx + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const float* tail = reinterpret_cast<const float*>(a.tc + (size_t)kImageStages * kStageBytes);
  for (int i = tid; i < a.C * 64 + 64; i += kTcThreads) w1_s[i] = __ldg(tail + i);
  for (int i = tid; i < kBiasFloats; i += kTcThreads) bias_s[i] = __ldg(tail + a.C * 64 + 64 + i);
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&peer_full[i], 1);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps * CG);   // the epilogue warps of BOTH CTAs report to the leader
    }
    for (int i = 0; i < 4; ++i) mbar_init(&act_ready[i], kEpiWarps * CG);
    mbar_init(accx, kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Work is dealt out in tile SLOTS of CG tiles: slot t of a cluster is tile t*CG + rank for its CTA `rank`.  Every
  // cluster runs the same number of slots; tiles past the end of the work list are dummies (MMAs run, nothing is
  // stored).  With CG = 2 the two CTAs of a pair walk their tiles in lock step: one MMA instruction stream, issued
  // by the leader, serves both.
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;
  const long long n_slots = (a.total_tiles + CG - 1) / CG;
  const long long n_clusters = gridDim.x / CG;
  const long long per_cta = (n_slots + n_clusters - 1) / n_clusters;
  const long long t_begin = per_cta * (blockIdx.x / CG);
  const long long t_end = t_begin + per_cta;

  if (warp == 0) {
    // ================================ producer ================================
    // The stream moves PAIRS of stages: one 32 KB bulk copy and one full/empty barrier per pair of ring slots
    // (one thread can start a bulk copy only every ~380 clk whatever its size, tests/cuda/bulk_rate.cu, and every
    // barrier the MMA warp polls costs it issue time).
    // The whole warp runs the loop and one elected lane issues, as in the MMA warp: issued from a single
    // divergent lane, every UBLKCP sat in an ELECT + 4x R2UR.BROADCAST waterfall and cost ~380 clk.
    {
      const long long n_pairs = (t_end - t_begin) * kPairsPerTile;
      const bool free_weights = (a.debug & 1) != 0;   // experiment: how fast is the kernel when weights cost nothing?
      // CG = 1: the whole 32 KB pair; CG = 2: this CTA's 16 KB of it, from its own stream of the pair image
      const uint8_t* image = CG == 1 ? a.tc : a.tc + tc_pair_image_offset(a.C) + (size_t)cta_rank * kPairsPerTile * kSlotBytes;
      uint32_t pair = 0, phase = 0;
      int img = 0;   // pair within the tile's stream
      for (long long j = 0; j < n_pairs; ++j) {
        mbar_wait(&empty[pair], phase ^ 1);
        uint8_t* dst = ring + (size_t)pair * kSlotBytes;
        const uint8_t* src = image + (size_t)img * kSlotBytes;
        if (elect_one()) {
          if (free_weights && j >= kPairsPerTile) {
            mbar_arrive(&full[pair]);
          } else {
            mbar_expect_tx(&full[pair], kSlotBytes);
            bulk_copy_g2s(dst, src, kSlotBytes, &full[pair]);
          }
        }
        __syncwarp();
        if (++img == kPairsPerTile) img = 0;
        if (++pair == kRingPairs) { pair = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && CG == 2 && cta_rank != 0) {
    // ================================ follower of a pair: relay ================================
    // The leader issues the MMAs for both CTAs and must know that THIS CTA's half of a ring pair has landed: a
    // bulk copy can only signal a barrier of the CTA it writes to, so this warp forwards every completion.
    const long long n_pairs = (t_end - t_begin) * kPairsPerTile;
    uint32_t pair = 0, phase = 0;
    for (long long j = 0; j < n_pairs; ++j) {
      mbar_wait(&full[pair], phase);
      if (lane == 0) mbar_arrive_remote_relaxed(&peer_full[pair], 0);
      __syncwarp();
      if (++pair == kRingPairs) { pair = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The WHOLE warp runs this loop and one elected lane issues: with warp-uniform control flow the
    // operands of tcgen05.mma / tcgen05.commit sit in uniform registers.  Run from a single divergent
    // lane instead, every UTCHMMA was wrapped in an ELECT + 4x R2UR.BROADCAST "waterfall" loop and
    // took ~140 clk to issue — more than the 64 clk it executes (timeline in profiles/r01_tc_timeline.md).
    {
      // instruction descriptor: D f32 (bit 4), A bf16 (bit 7), B bf16 (bit 10), both K-major, N = 128, M = 128 per CTA
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | (((128u * CG) >> 4) << 24);
      const uint32_t tmem_u = __shfl_sync(FULL_MASK, tmem, 0);                        // provably uniform
      const uint32_t ring_u = __shfl_sync(FULL_MASK, smem_u32(ring), 0);
      // ring pairs per trip: one (8 MMAs).  Two per trip halve the polls, but measured slower once the layer shapes
      // became compile-time constants (global mode 1.02 vs 1.05 ms) and stall a 4-pair ring (cell mode)
      constexpr int kTrip = (B200BEV_TC_TRIP) ? (B200BEV_TC_TRIP) : 1;
      uint32_t pair = 0, phase = 0;   // ring pair slot and the parity of its round
      uint32_t act_phase = 0;         // bit kp: parity of the next phase of act_ready[kp]
      uint32_t acc_parity = 0;  // bit b: parity of the next use of accumulator b
      int tr = 0;
      const bool wait_weights = !(a.debug & 4);      // debug bit 2: issue without waiting for the weights
      // 8 MMAs over one ring pair (K = 128): D[d_addr] (+)= A[a_addr .. a_addr+64 columns) . B[pair]^T
      auto issue_pair = [&](uint32_t d_addr, uint32_t a_addr, uint32_t pr, bool first, bool half) {
        const uint64_t bdesc0 = make_b_desc(ring_u + pr * kSlotBytes);
#pragma unroll
        for (int s = 0; s < 4; ++s) umma_ts<CG>(d_addr, a_addr + s * 8, bdesc0 + (uint64_t)(s * 2), idesc, !(first && s == 0));
        if (!half) {
          const uint64_t bdesc1 = make_b_desc(ring_u + pr * kSlotBytes + kSlotBytes / 2);
#pragma unroll
          for (int s = 0; s < 4; ++s) umma_ts<CG>(d_addr, a_addr + 32 + s * 8, bdesc1 + (uint64_t)(s * 2), idesc, 1u);
        }
        if constexpr (CG == 1) tc_commit(&empty[pr]);
        else tc_commit_pair(&empty[pr]);
      };
      auto commit_acc = [&](int buf) {
        if constexpr (CG == 1) tc_commit(&acc_full[buf]);
        else tc_commit_pair(&acc_full[buf]);
      };
      // One network layer, its shape a compile-time constant: chunks and trips are fully unrolled, so the accumulator
      // choice, the A-operand columns and the loop tests cost the issuing thread nothing at run time — every
      // instruction on this path delays the tensor pipe (tests/cuda/umma_rate.cu: a handful of extra integer
      // instructions per 4 MMAs take the pipe from 64 to 80 clk per MMA).
      auto issue_layer = [&](auto layer_tag) {
        constexpr int layer = decltype(layer_tag)::value;           // 0..3 = network layers 2..5
        constexpr int kpairs = layer == 0 ? 1 : (1 << layer) >> 1;   // K / 128 (layer 2: half a pair)
        constexpr int nchunks = 1 << layer;                          // N / 128
        constexpr uint32_t a_col = layer == 0 ? kColAct1 : layer == 1 ? kColAct2 : layer == 2 ? kColAct3 : kColAct4;
        if (lane == 0) trace_ev<TRACE>(a, tr, kTraceMma, 0x100 + layer);            // first MMA of the layer
        auto issue_chunk = [&](int c) {
          const int buf = acc_buffer(layer, c);
          mbar_wait(&acc_empty[buf], ((acc_parity >> buf) & 1) ^ 1);
          acc_parity ^= 1u << buf;
          tc_fence_after();
          if (lane == 0) trace_ev<TRACE>(a, tr, kTraceMma, 0x120 + layer * 8 + c);  // accumulator free, issuing
          const uint32_t d_addr = tmem_u + acc_column(buf);
#pragma unroll
          for (int kp = 0; kp < kpairs; kp += kTrip) {
            // up to two ring pairs (16 MMAs = 1024 clk of tensor work) per trip: the issue side of a trip (barrier
            // polls, fence, descriptor arithmetic in the uniform datapath, releases) must stay below the time the
            // MMAs it issues take, or the tensor pipe starves
            const bool two = kTrip == 2 && kp + 1 < kpairs;
            uint32_t pr1 = pair + 1, ph1 = phase;
            if (pr1 == kRingPairs) { pr1 = 0; ph1 ^= 1; }
            // The A operand arrives chunk by chunk: K-pair kp of this layer is the 128 channels that chunk kp of
            // the layer before produced, so the first chunk of a layer starts while the pipe still works on the
            // previous layer's last chunks.  One barrier per K-pair index: on each of them production and
            // consumption alternate strictly (a single barrier would let the epilogue get two phases ahead of
            // this warp, which a parity wait cannot tell from zero phases ahead).
            if (c == 0) {
              mbar_wait(&act_ready[kp], (act_phase >> kp) & 1);
              act_phase ^= 1u << kp;
              if (two) {
                mbar_wait(&act_ready[kp + 1], (act_phase >> (kp + 1)) & 1);
                act_phase ^= 1u << (kp + 1);
              }
            }
            if (wait_weights) {
              mbar_wait(&full[pair], phase);
              if constexpr (CG == 2) mbar_wait(&peer_full[pair], phase);
              if (two) {
                mbar_wait(&full[pr1], ph1);
                if constexpr (CG == 2) mbar_wait(&peer_full[pr1], ph1);
              }
            }
            tc_fence_after();
            const uint32_t a_addr = tmem_u + a_col + kp * 64;
            if (elect_one()) {
              issue_pair(d_addr, a_addr, pair, kp == 0, layer == 0);
              if (two) issue_pair(d_addr, a_addr + 64, pr1, false, false);
              if (kp + kTrip >= kpairs) commit_acc(buf);
            }
            __syncwarp();
            if (two) { pair = pr1; phase = ph1; }
            if (++pair == kRingPairs) { pair = 0; phase ^= 1; }
          }
          if (lane == 0) trace_ev<TRACE>(a, tr, kTraceMma, 0x160 + layer * 8 + c);  // chunk issued
        };
        // Global mode unrolls the chunks too (accumulator choice and the c == 0 tests fold away: 1.10 -> 1.05 ms).
        // Cell mode keeps them rolled: its epilogue is the bottleneck and lives on instruction fetch — with
        // the chunks unrolled here the issuer got faster and the epilogue 40 % slower.
        if constexpr (CELL && !(B200BEV_CELL_MMA_UNROLL)) {
#pragma unroll 1
          for (int c = 0; c < nchunks; ++c) issue_chunk(c);
        } else {
#pragma unroll
          for (int c = 0; c < nchunks; ++c) issue_chunk(c);
        }
      };
      for (long long t = t_begin; t < t_end; ++t) {
        issue_layer(std::integral_constant<int, 0>{});
        issue_layer(std::integral_constant<int, 1>{});
        issue_layer(std::integral_constant<int, 2>{});
        issue_layer(std::integral_constant<int, 3>{});
      }
    }
  } else {
    // ================================ epilogue (16 warps, 512 threads) ================================
    // Four warps share each TMEM lane quadrant and split the accumulator columns four ways ("part"):
    // with one epilogue warp per scheduler the dependent ALU/shuffle chains ran at ~5 clk per
    // instruction and the tensor pipe idled two thirds of the time (profiles/r01_tc_*); four warps
    // per scheduler hide that latency.
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may touch (hardware: warp id % 4)
    const int part = (warp - 2) >> 2;          // which quarter of the columns / of the tile's points this warp owns
    const int row = quad * 32 + lane;          // point within the tile = TMEM lane
    const uint32_t tm = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t full_phase = 0;                   // bit b: parity of the next completion of accumulator b
    // global mode: running maximum of the raw layer-5 accumulators of the current frame, one value per 128-channel chunk c:
    //   rmax[c] <-> channel c*128 + part*32 + lane, over the 32 points of this warp's TMEM lane quadrant (what the
    //   butterfly leaves in this lane)
    float rmax[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rmax[i] = -INFINITY;
    int cur_frame = -1;
    int tr = 0;
    const bool vec4_points = a.C == 4 && (reinterpret_cast<uintptr_t>(a.pts) & 15) == 0;
    float4 xpre = make_float4(0.f, 0.f, 0.f, 0.f);   // next tile's point, prefetched under layer 5
    bool have_pre = false;
    const int c_out = 1024;
    const float* b5 = bias_s + 128 + 256 + 512;
    // cell mode, part 0 only: cell of the tile's first slot, carried from tile to tile, and the in-grid count
    int c_carry = 0, n_in = 0;
    bool carry_valid = false;

    // one arrival per warp: every lane has fenced its tensor-memory accesses, __syncwarp orders them
    // before lane 0's arrive (512 per-thread arrivals on one barrier word cost more than the work they guard)
    // (CG = 2: acc_empty and act_ready live in the leader, whose MMA warp waits on them; the follower arrives remotely)
    auto warp_arrive = [&](uint64_t* bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1 || cta_rank == 0) mbar_arrive(bar);
        // relaxed: what the leader's MMA warp needs is that this warp's tensor-memory accesses have COMPLETED
        // (tcgen05.wait::ld / wait::st above) — tensor memory is not part of the generic memory model a release
        // would order, and a release.cluster arrive stalls the warp for ~1,000 clk
        else mbar_arrive_remote_relaxed(bar, 0);
      }
    };
    // the same, counting for two warps (cell mode, layer 5: half of the warps drain a chunk, the barrier expects all of them)
    auto warp_arrive2 = [&](uint64_t* bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1 || cta_rank == 0) mbar_arrive2(bar);
        else mbar_arrive2_remote_relaxed(bar, 0);
      }
    };
    uint32_t accx_parity = 0;
    auto acc_wait = [&](int buf) {
      mbar_wait(&acc_full[buf], (full_phase >> buf) & 1);
      full_phase ^= 1u << buf;
      tc_fence_after();
    };

    auto flush = [&](int frame) {
      if (CELL || a.out_global == nullptr) return;   // cell mode sends its maxima chunk by chunk
      int* o = reinterpret_cast<int*>(a.out_global + (size_t)frame * c_out);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int col = i * 128 + part * 32 + lane;
        const float v = fmaxf(rmax[i] + b5[col], 0.0f);   // bias + ReLU commute with the max
        atomicMax(o + col, __float_as_int(v));
        rmax[i] = -INFINITY;
      }
    };

    // the point of tile slot `slot_` of frame `f_` (zeros for a slot past the end of the frame)
    auto load_point = [&](int f_, int slot_, float* x) {
      const bool ok = slot_ < a.N;
      int p_ = slot_;
      if (CELL && ok) p_ = __ldg(a.perm + (size_t)f_ * a.N + slot_);
      const float* src = a.pts + ((size_t)f_ * a.N + (ok ? p_ : 0)) * a.C;
      if (vec4_points) {
        const float4 v = ok ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
#pragma unroll
        for (int k = 4; k < kMaxCin; ++k) x[k] = 0.0f;
      } else {
#pragma unroll
        for (int k = 0; k < kMaxCin; ++k) x[k] = (k < a.C && ok) ? __ldg(src + k) : 0.0f;
      }
    };
    // layer 1 on CUDA cores (K = C_in is 4): part p's 16 of the 64 channels of relu(W1 x + b1), as bf16 pairs.
    // Weights and bias of a part are contiguous: 128-bit broadcast loads.
    auto layer1 = [&](const float* x, uint32_t* packed, int p) {
      float acc[16];
      {
        const float4* bp = reinterpret_cast<const float4*>(w1_s + a.C * 64 + p * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 b4 = bp[q];
          acc[4 * q] = b4.x; acc[4 * q + 1] = b4.y; acc[4 * q + 2] = b4.z; acc[4 * q + 3] = b4.w;
        }
      }
      auto add_k = [&](int k, float xk) {
        const float4* wp = reinterpret_cast<const float4*>(w1_s + k * 64 + p * 16);
        const float2 x2 = make_float2(xk, xk);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w4 = wp[q];
          const float2 a0 = fma_f32x2(make_float2(w4.x, w4.y), x2, make_float2(acc[4 * q], acc[4 * q + 1]));
          const float2 a1 = fma_f32x2(make_float2(w4.z, w4.w), x2, make_float2(acc[4 * q + 2], acc[4 * q + 3]));
          acc[4 * q] = a0.x; acc[4 * q + 1] = a0.y; acc[4 * q + 2] = a1.x; acc[4 * q + 3] = a1.y;
        }
      };
      if (a.C == 4) {
        add_k(0, x[0]); add_k(1, x[1]); add_k(2, x[2]); add_k(3, x[3]);
      } else {
#pragma unroll
        for (int k = 0; k < kMaxCin; ++k)
          if (k < a.C) add_k(k, x[k]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) packed[j] = relu_pack_bf16x2(acc[2 * j], acc[2 * j + 1]);
    };
    // act1 -> TMEM [0,32) and tell the MMA warp.  Legal only while no MMA reads act4 (which aliases act1):
    // before the first tile, or after the last layer-5 accumulator of the previous tile has completed.
    // cell mode, from the second tile on: the warps that drain chunk 7 (parts 2 and 3) see the end of the tile's MMAs first,
    // so they compute and store layer 1 for parts 0 and 1 as well (same TMEM lane quadrant, same points) and arrive for two;
    // parts 0 and 1 walk chunk 6 meanwhile and are off the path to the next tile's first MMA
    auto publish_act1_pair = [&](const uint32_t* packed) {
      TC_ST8(tm + kColAct1 + part * 8, packed);
      TC_ST8(tm + kColAct1 + (part - 2) * 8, (packed + 8));
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      warp_arrive2(&act_ready[0]);
    };
    auto publish_act1 = [&](const uint32_t* packed) {
      TC_ST8(tm + kColAct1 + part * 8, packed);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      warp_arrive(&act_ready[0]);
    };

    // (frame, tile-in-frame) of the running tile slot, advanced incrementally: one 64-bit division per
    // kernel instead of four per tile in each of the 16 warps
    const long long tile0 = t_begin * CG + cta_rank;   // this CTA's first tile
    int tf = (int)(tile0 / a.tiles_per_frame), tt = (int)(tile0 % a.tiles_per_frame);
    {
      // layer 1 of the first tile; every later tile's layer 1 is computed under the previous tile's layer 5
      const bool dummy0 = tile0 >= a.total_tiles;
      float x[kMaxCin];
      load_point(dummy0 ? 0 : tf, dummy0 ? a.N : tt * kTileM + row, x);
      uint32_t packed[8];
      layer1(x, packed, part);
      publish_act1(packed);
    }
    for (long long t = t_begin; t < t_end; ++t) {
#ifndef B200BEV_TC_TRACER_TID
#define B200BEV_TC_TRACER_TID 64   // the epilogue thread whose clock stamps the debug timeline records (64: warp 2, part 0; 320: warp 10, part 2)
#endif
      const bool tracer = (tid == B200BEV_TC_TRACER_TID);
      if (tracer) trace_ev<TRACE>(a, tr, kTraceEpi, 0x200);          // tile start
      const long long tile = t * CG + cta_rank;
      const bool dummy = tile >= a.total_tiles;
      const int f = dummy ? cur_frame : tf;
      const int s0 = dummy ? a.N : tt * kTileM;
      // this CTA's next tile, for the prefetch below
      int tf_next = tf, tt_next = tt + CG;
      while (tt_next >= a.tiles_per_frame) { tt_next -= a.tiles_per_frame; ++tf_next; }
      const bool new_frame = f != cur_frame;
      if (new_frame) {
        if (cur_frame >= 0) flush(cur_frame);
        cur_frame = f;
      }
      const int slot = s0 + row;                // position in input order (global) or in cell order (CELL)
      const bool valid = slot < a.N;            // a dummy tile has no valid slot
      const bool partial = s0 + kTileM > a.N;
      if (CELL) {
        // Cell id of every tile slot (part 0, 128 threads <-> 128 slots).  The slots are in cell order, so
        // cid(slot) = largest c with offsets[c] <= slot: every cell that starts inside the tile drops its
        // index at its first slot (atomicMax: of several cells starting at one slot, all but the last are
        // empty), a prefix maximum fills the runs, and the cell of the last slot is carried to the next
        // tile — one coalesced read of `offsets` per tile instead of a 12-step binary search per slot.
        if (part == 0) {
          auto bar128 = [] { asm volatile("bar.sync 2, 128;" ::: "memory"); };
          const int32_t* foff = a.offsets + (size_t)(f < 0 ? 0 : f) * (a.n_cells + 1);
          if (!dummy && (new_frame || !carry_valid)) {
            n_in = __ldg(foff + a.n_cells);       // in-grid points come first in perm
            int lo = 0;
            if (s0 > 0 && s0 < n_in) {            // a CTA that starts in the middle of a frame: search once
              int hi = a.n_cells;
              while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(foff + mid) <= s0) lo = mid; else hi = mid;
              }
            }
            c_carry = lo;
            carry_valid = true;
          }
          const bool in_grid = !dummy && s0 < n_in;   // uniform over the 128 threads
          int v = -1;
          if (in_grid) {
            cid_s[row] = row == 0 ? c_carry : -1;
            bar128();
            int c_scan = c_carry, it = 0;
            while (true) {
              const int c = c_scan + row;
              const int o = c < a.n_cells ? __ldg(foff + c) : 0x7fffffff;
              // a cell that starts before the tile (the carried one, or — for a CTA of a pair, which skips its peer's
              // tile — any cell that started in between) competes for slot 0
              if (o < s0 + kTileM && o < n_in) atomicMax(&cid_s[o > s0 ? o - s0 : 0], c);
              if (row == kTileM - 1) scan_s[4 + (it & 1)] = (o < s0 + kTileM) && (c + 1 < a.n_cells);
              bar128();
              if (!scan_s[4 + (it & 1)]) break;
              c_scan += kTileM;
              ++it;
            }
            v = cid_s[row];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) v = max(v, __shfl_up_sync(FULL_MASK, v, d));   // -1 is the identity
            if (lane == 31) scan_s[quad] = v;
            bar128();
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < quad) v = max(v, scan_s[q]);
            c_carry = max(max(scan_s[0], scan_s[1]), max(scan_s[2], scan_s[3]));   // cell of the tile's last slot
          }
          const int cid = (in_grid && valid && slot < n_in) ? v : -1;
          cid_s[row] = cid;
          bar128();
          // run ends: one flag word per 32 slots.  The last slot of a word never carries a flag for a run
          // that goes on in the next word; the walker of that word hands such a run over with an atomic.
          const int next = row < kTileM - 1 ? cid_s[row + 1] : -2;
          const unsigned ends = __ballot_sync(FULL_MASK, cid >= 0 && next != cid);
          if (lane == 0) endmask_s[quad] = ends;
        }
        // cid_s / endmask_s become visible to the walkers at the barrier in front of the layer-5 loop
      }

      // ---- layers 2..4: accumulator -> bias, ReLU, bf16 -> next A operand in TMEM ----
#pragma unroll 1
      for (int layer = 0; layer < 3; ++layer) {
        const int nchunks = 1 << layer;
        const uint32_t out_col = layer == 0 ? kColAct2 : layer == 1 ? kColAct3 : kColAct4;
        const float* bl = bias_s + (layer == 0 ? 0 : layer == 1 ? 128 : 384);
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          const int buf = acc_buffer(layer, c);
          acc_wait(buf);
          if (tracer) trace_ev<TRACE>(a, tr, kTraceEpi, 0x210 + layer * 8 + c);   // accumulator seen
          {
            const int q = part;   // this warp's 32 of the chunk's 128 columns
            uint32_t r[32];
            TC_LD32(r, tm + acc_column(buf) + q * 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            warp_arrive(&acc_empty[buf]);   // the values are in registers: the next chunk may overwrite the accumulator
            if (layer == 2 && c == 1 && lane == 0) mbar_arrive(accx);   // ... and, locally: this warp is done with accX
            uint32_t packed[16];
            const float4* bq = reinterpret_cast<const float4*>(bl + c * 128 + q * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {     // bias by FADD2, ReLU + bf16 + pack by F2FP.RELU: two instructions per two values
              const float4 b4 = bq[j];
              const float2 s0 = add_f32x2(make_float2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])), make_float2(b4.x, b4.y));
              const float2 s1 = add_f32x2(make_float2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), make_float2(b4.z, b4.w));
              packed[2 * j] = relu_pack_bf16x2(s0.x, s0.y);
              packed[2 * j + 1] = relu_pack_bf16x2(s1.x, s1.y);
            }
            // chunks 2 and 3 of layer 4 land on accX: every warp of this CTA must have drained it (layer 4, chunk 1) first
            if (layer == 2 && c == 2) {
              mbar_wait(accx, accx_parity);
              accx_parity ^= 1;
            }
            TC_ST16(tm + out_col + c * 64 + q * 16, packed, 0);
          }
          // chunk c is K-pair c of the next layer: tell the MMA warp these 128 channels are in place
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          warp_arrive(&act_ready[c]);
          if (tracer) trace_ev<TRACE>(a, tr, kTraceEpi, 0x230 + layer * 8 + c);   // next operand stored
        }
      }

      // ---- prefetch the next tile's point: its DRAM latency hides under layer 5 ----
      const bool more = t + 1 < t_end;
      const bool dn = tile + CG >= a.total_tiles;
      const int fn = dn ? 0 : tf_next;
      const int slotn = dn ? a.N : tt_next * kTileM + row;
      have_pre = false;
      if (vec4_points && more && (!CELL || part >= 2)) {
        xpre = make_float4(0.f, 0.f, 0.f, 0.f);
        if (slotn < a.N) {
          const int pn = CELL ? __ldg(a.perm + (size_t)fn * a.N + slotn) : slotn;
          xpre = __ldg(reinterpret_cast<const float4*>(a.pts + ((size_t)fn * a.N + pn) * 4));
        }
        have_pre = true;
      }

      // ---- layer 5: max over the tile's points, channel chunk by channel chunk ----
      // layer 1 of the NEXT tile (computed while the tensor pipe works on this tile's last chunk)
      auto next_layer1 = [&](uint32_t* packed1) {
        float x[kMaxCin];
        if (have_pre) {
          x[0] = xpre.x; x[1] = xpre.y; x[2] = xpre.z; x[3] = xpre.w;
#pragma unroll
          for (int k = 4; k < kMaxCin; ++k) x[k] = 0.0f;
        } else {
          load_point(fn, slotn, x);
        }
        layer1(x, packed1, part);
        if constexpr (CELL) layer1(x, packed1 + 8, part - 2);   // cell mode: also the share of the warp that does not drain chunk 7
      };
      if constexpr (!CELL) {
        // The loop stays ROLLED; the eight running maxima live in registers all the same: rmax[0] is always the
        // current chunk's, and the array is rotated by one after every chunk — back in place after the eighth.
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          const int buf = c & 1;
          uint32_t packed1[8];
          if (c == 7 && more) next_layer1(packed1);
          acc_wait(buf);
          if (tracer) trace_ev<TRACE>(a, tr, kTraceEpi, 0x250 + c);    // layer-5 accumulator seen
          // every MMA of this tile has completed (the pipe retires in order): act4 is dead, act1 may be replaced
          if (c == 7 && more) publish_act1(packed1);
          const uint32_t acc_col = buf ? kColAcc1 : kColAcc0;
          if (a.debug & 2) {   // debug bit 1: no layer-5 epilogue at all (results are garbage; timing experiment)
            warp_arrive(&acc_empty[buf]);
            continue;
          }
          {
            const int g = part;   // this warp's 32 of the chunk's 128 channels
            uint32_t r[32];
            TC_LD32(r, tm + acc_col + g * 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // this warp's share of the chunk is in registers: hand the accumulator back to the MMA warp
            warp_arrive(&acc_empty[buf]);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (partial && !valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = -INFINITY;
            }
            butterfly_level<16>(v, lane);
            butterfly_level<8>(v, lane);
            butterfly_level<4>(v, lane);
            butterfly_level<2>(v, lane);
            butterfly_level<1>(v, lane);
            rmax[0] = fmaxf(rmax[0], v[0]);  // lane l holds channel c*128 + g*32 + l
            if (tracer) trace_ev<TRACE>(a, tr, kTraceEpi, 0x260 + c);  // layer-5 chunk reduced
          }
          {
            const float r0 = rmax[0];
#pragma unroll
            for (int i = 0; i < 7; ++i) rmax[i] = rmax[i + 1];
            rmax[7] = r0;
          }
        }
      } else {
        // Cell mode.  A warp drains every SECOND chunk (the chunks whose accumulator buffer is part >> 1) and takes 64 of its
        // 128 channels, as two 32 x 32 blocks A and B that go through the SAME run walk: the run boundaries belong to the
        // warp's 32 slots, not to the channels, so the control flow of the walk (one warp-uniform test per four slots, one
        // branch per run end, the cell id, the canvas address) is paid once per 64 channels, and the two maximum chains
        // are independent instructions that hide each other's latency.  (One block per warp and chunk, the first form:
        // ~230 dependent instructions per chunk and warp at four warps per scheduler took 2,500 clk against 1,750 clk of
        // MMAs — the run walk alone was 0.32 ms of the 1.40 ms kernel.)  A warp never looks at the accumulator barriers of
        // the other buffer during layer 5 (four phases each per tile: its parity bookkeeping stays right), the draining warps
        // arrive for two on acc_empty, and the one thing that hangs on chunk 7 — act1 of the next tile may be stored once every
        // MMA of this tile has completed — is done by the warps that drain chunk 7, for all four parts (publish_act1_pair).
        // The per-frame global maximum goes to out_global chunk by chunk (two 128-byte atomics per warp and chunk, L2-resident)
        // instead of living in eight rotating registers: at 96 registers per thread those were spilled.
        epi_bar_sync();     // this tile's cell ids and run-end flags (written by part 0 at the top) are in place
        const int own = part >> 1;                 // accumulator buffer / chunk parity this warp drains
        const int half = part & 1;                 // which 64 channels of the chunk
        float* tp = tile_s + (warp - 2) * kWarpTile;
        const float* trow = tp + lane * kTStride;
        const int* cids = cid_s + quad * 32;
        // the same word in every lane: broadcast through a shuffle so that the compiler knows the branches on its
        // bits are warp-uniform (a possibly-divergent branch per slot costs a BSSY/BSYNC pair and its resolve latency)
        const uint32_t ends = __shfl_sync(FULL_MASK, endmask_s[quad], 0);
        const int first_end = __ffs(ends) - 1;     // the run that ends here may have begun in the stretch before
        float* canvas_f = a.out_canvas + (size_t)(f < 0 ? 0 : f) * a.n_cells * c_out;
        uint32_t packed1[16];
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const bool last = cc == 3;
          const int c = 2 * cc + own;
          if (own == 1 && last && more) next_layer1(packed1);
          acc_wait(own);
          if (tracer) trace_ev<TRACE>(a, tr, kTraceEpi, 0x250 + c);    // layer-5 accumulator seen
          // chunk 7 complete: every MMA of this tile has completed (the pipe retires in order), act4 is dead, act1 may be replaced
          if (own == 1 && last && more) publish_act1_pair(packed1);
          const uint32_t acc_col = (own ? kColAcc1 : kColAcc0) + half * 64;
          if (a.debug & 2) {   // debug bit 1: no layer-5 epilogue at all (results are garbage; timing experiment)
            warp_arrive2(&acc_empty[own]);
          } else {
            // Transpose the two 32 points x 32 channels blocks through the warp's PRIVATE shared-memory tile, one after the
            // other: tp[channel][point].  Nobody else touches the tile, so the hand-over is a __syncwarp, not a block barrier.
            float4 qa[8], qb[8];
            {
              uint32_t ra[32], rb[32];
              TC_LD32(ra, tm + acc_col);
              TC_LD32(rb, tm + acc_col + 32);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              // both blocks are in registers: hand the accumulator back to the MMA warp
              warp_arrive2(&acc_empty[own]);
              if (partial && !valid) {
#pragma unroll
                for (int j = 0; j < 32; ++j) ra[j] = rb[j] = 0xff800000u;   // -inf
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) tp[j * kTStride + lane] = __uint_as_float(ra[j]);
              __syncwarp();
#pragma unroll
              for (int b4 = 0; b4 < 8; ++b4) qa[b4] = *reinterpret_cast<const float4*>(trow + b4 * 4);
              __syncwarp();             // block A is read out
#pragma unroll
              for (int j = 0; j < 32; ++j) tp[j * kTStride + lane] = __uint_as_float(rb[j]);
              __syncwarp();
              // block B: the first sixteen slots now, the rest half way through the walk (96 registers per thread do not
              // hold 2 x 32 values next to the walk's state)
#pragma unroll
              for (int b4 = 0; b4 < 4; ++b4) qb[b4] = *reinterpret_cast<const float4*>(trow + b4 * 4);
            }
            if (tracer) trace_ev<TRACE>(a, tr, kTraceEpi, 0x271);   // both blocks transposed
            // This lane now owns channels ch and ch + 32 over the warp's 32 tile slots [quad*32, quad*32+32), which are in
            // cell order: the maximum of each run of equal cell ids goes to the canvas.  A run that touches the first or
            // the last slot of the stretch may continue in a neighbouring stretch or tile -> atomic max; any other run is
            // the cell's only writer -> plain store.  The 32 lanes see the same run boundaries and write two 128-byte
            // pieces of a canvas row.  Zeros are stored like any other value (the canvas starts at zero, and a store that
            // depends on the value would be a divergent branch per run end).
            const int ch = c * 128 + half * 64 + lane;
            const float bias_a = b5[ch], bias_b = b5[ch + 32];
            float* canvas = canvas_f + ch;
            float ma = -INFINITY, mb = -INFINITY, ga = -INFINITY, gb = -INFINITY;
            auto walk_group = [&](int b4) {
              const float ea[4] = {qa[b4].x, qa[b4].y, qa[b4].z, qa[b4].w};
              const float eb[4] = {qb[b4].x, qb[b4].y, qb[b4].z, qb[b4].w};
              if (((ends >> (b4 * 4)) & 0xfu) == 0u) {   // no run ends in these four slots: most groups
                ma = fmaxf(fmaxf(ma, fmaxf(ea[0], ea[1])), fmaxf(ea[2], ea[3]));
                mb = fmaxf(fmaxf(mb, fmaxf(eb[0], eb[1])), fmaxf(eb[2], eb[3]));
                return;
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int pt = b4 * 4 + k;
                ma = fmaxf(ma, ea[k]);
                mb = fmaxf(mb, eb[k]);
                if (ends & (1u << pt)) {
                  const float va = fmaxf(ma + bias_a, 0.0f), vb = fmaxf(mb + bias_b, 0.0f);
                  float* dst = canvas + (size_t)cids[pt] * c_out;
#ifdef B200BEV_DEBUG_ENV
                  if (a.debug & 16) {          // debug bit 4: no canvas stores at all (garbage results; timing experiment)
                    ga = fmaxf(ga, va);
                    gb = fmaxf(gb, vb);
                  } else if (a.debug & 8) {    // debug bit 3: plain stores in place of the atomics
                    dst[0] = va;
                    dst[32] = vb;
                  } else
#endif
                  if (pt == first_end || pt == 31) {
                    atomicMax(reinterpret_cast<int*>(dst), __float_as_int(va));
                    atomicMax(reinterpret_cast<int*>(dst + 32), __float_as_int(vb));
                  } else {
                    dst[0] = va;
                    dst[32] = vb;
                  }
                  ga = fmaxf(ga, ma);
                  gb = fmaxf(gb, mb);
                  ma = -INFINITY;
                  mb = -INFINITY;
                }
              }
            };
#pragma unroll
            for (int b4 = 0; b4 < 4; ++b4) walk_group(b4);
#pragma unroll
            for (int b4 = 4; b4 < 8; ++b4) qb[b4] = *reinterpret_cast<const float4*>(trow + b4 * 4);
            __syncwarp();             // the tile is read out: the next chunk may overwrite it
#pragma unroll
            for (int b4 = 4; b4 < 8; ++b4) walk_group(b4);
            // a run still open at the last slot goes on in the next stretch: hand its partial maximum over
            if (!(ends >> 31) && cids[31] >= 0) {
              float* dst = canvas + (size_t)cids[31] * c_out;
#ifdef B200BEV_DEBUG_ENV
              if (a.debug & 16) {
              } else if (a.debug & 8) {
                dst[0] = fmaxf(ma + bias_a, 0.0f);
                dst[32] = fmaxf(mb + bias_b, 0.0f);
              } else
#endif
              {
                atomicMax(reinterpret_cast<int*>(dst), __float_as_int(fmaxf(ma + bias_a, 0.0f)));
                atomicMax(reinterpret_cast<int*>(dst + 32), __float_as_int(fmaxf(mb + bias_b, 0.0f)));
              }
            }
            // whatever is left (open run, out-of-grid tail) still counts globally; bias + ReLU commute with the max
            if (a.out_global != nullptr && f >= 0) {
              int* o = reinterpret_cast<int*>(a.out_global + (size_t)f * c_out + ch);
              atomicMax(o, __float_as_int(fmaxf(fmaxf(ga, ma) + bias_a, 0.0f)));
              atomicMax(o + 32, __float_as_int(fmaxf(fmaxf(gb, mb) + bias_b, 0.0f)));
            }
            if (tracer) trace_ev<TRACE>(a, tr, kTraceEpi, 0x273);     // runs walked
          }
        }
        // cid_s / endmask_s belong to this tile until everyone has walked its last chunk
        epi_bar_sync();
      }
      tf = tf_next;
      tt = tt_next;
    }
    if (cur_frame >= 0) flush(cur_frame);
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // no CTA leaves (or frees tensor memory) while its peer may still signal or compute
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// Re-tiles the fp32 blob (W^T (K x N) + bias per layer) into the stage image: stage -> (layer, 128-channel
// block, 64-k block); row r of a stage holds 64 bf16 of W[n0 + r][k0 ..], its eight 16-byte chunks XOR-
// swizzled with (r & 7) exactly as SWIZZLE_128B reads them.
__global__ void __launch_bounds__(256) pack_bf16_kernel(const float* __restrict__ params, int C, uint8_t* __restrict__ tc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (stage, row, chunk)
  const int total = kImageStages * 128 * 8;
  // fp32 blob offsets of layers 1..5
  long long w_off[5], b_off[5];
  const int dims[6] = {C, 64, 128, 256, 512, 1024};
  long long off = 0;
  for (int l = 0; l < 5; ++l) {
    w_off[l] = off;
    off += (long long)dims[l] * dims[l + 1];
    b_off[l] = off;
    off += dims[l + 1];
  }
  if (idx < total) {
    const int s_img = idx / (128 * 8), r = (idx / 8) % 128, ch = idx % 8;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);   // s_img == 1: the padding stage of the layer-2 pair
    if (s_img != 1) {
      const int s = s_img == 0 ? 0 : s_img - 1;
      int layer, i;
      if (s < 1) { layer = 1; i = s; }
      else if (s < 5) { layer = 2; i = s - 1; }
      else if (s < 21) { layer = 3; i = s - 5; }
      else { layer = 4; i = s - 21; }
      const int kchunks = dims[layer] / 64;
      const int n0 = (i / kchunks) * 128, k0 = (i % kchunks) * 64;
      const int Nout = dims[layer + 1];
      const float* wt = params + w_off[layer];
      __align__(16) __nv_bfloat16 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16_rn(wt[(size_t)(k0 + ch * 8 + j) * Nout + n0 + r]);
      val = *reinterpret_cast<uint4*>(v);
    }
    // single-CTA image: 128 rows per stage
    *reinterpret_cast<uint4*>(tc + (size_t)s_img * kStageBytes + r * 128 + ((ch ^ (r & 7)) * 16)) = val;
    // pair image (cta_group::2): rank h of a pair streams rows 64h .. 64h+63 of every stage, 8 KB per stage
    const int h = r / 64, rr = r % 64;
    *reinterpret_cast<uint4*>(tc + tc_pair_image_offset(C) + ((size_t)h * kImageStages + s_img) * (kStageBytes / 2) + rr * 128 +
                              ((ch ^ (rr & 7)) * 16)) = val;
  }
  // tail: W1^T, b1, then the biases of layers 2..5, fp32
  float* tailp = reinterpret_cast<float*>(tc + (size_t)kImageStages * kStageBytes);
  const int n_w1 = C * 64 + 64;
  for (int i = idx; i < n_w1 + kBiasFloats; i += gridDim.x * blockDim.x) {
    float v;
    if (i < n_w1) {
      v = params[i];  // W1^T then b1 are the first C*64 + 64 floats of the blob
    } else {
      int j = i - n_w1, l = 1;
      while (j >= dims[l + 1]) { j -= dims[l + 1]; ++l; }
      v = params[b_off[l] + j];
    }
    tailp[i] = v;
  }
}

bool tc_dims_supported(const int32_t* dims, int n_layers) {
  return dims && n_layers == 5 && dims[0] >= 1 && dims[0] <= kMaxCin && dims[1] == 64 && dims[2] == 128 && dims[3] == 256 &&
         dims[4] == 512 && dims[5] == 1024;
}

// 1: every CTA runs its own MMA stream (cta_group::1).  2: CTA pairs (cta_group::2) — the leader's one instruction
// stream drives both SMs, each CTA keeps half of every weight stage, so the shared-memory traffic of the B operand
// and of the weight stream, and the MMA instructions issued per tile, are halved.  B200BEV_TC_CLUSTER=1|2 overrides.
int tc_cluster_size() {
  const char* e = debug_env("B200BEV_TC_CLUSTER");
  if (e) {
    const int v = atoi(e);
    if (v == 1 || v == 2) return v;
  }
  return 2;   // measured on B200 (32 x 35,000 points): pairs 0.98 ms global / 1.43 ms cell, single CTAs 1.03 / 1.53
}

size_t tc_smem_bytes(bool cell, int cluster) {
  const int stages = cell ? (cluster == 2 ? kStagesCellPair : kStagesCell) : kStagesGlobal;
  return 1024 + (size_t)stages * kStageBytes + (cell ? (kEpiWarps * kWarpTile + 128 + 4) * sizeof(float) : 0) +
         (cell ? 8 * sizeof(int) : 0) + (kBiasFloats + kMaxCin * 64 + 64) * sizeof(float) +
         (3 * stages + 11) * sizeof(uint64_t) + 16;
}

}  // namespace

int pointnet_encode_tc(const float* points, int B, int N, int C, const float* params, const int32_t* dims, int n_layers,
                       const int32_t* perm, const int32_t* offsets, int n_cells, const void* tc_params, float* out_global,
                       float* out_canvas, cudaStream_t st) {
  (void)params;
  if (!tc_dims_supported(dims, n_layers) || dims[0] != C) return B200BEV_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(tc_params) & 15) != 0) return B200BEV_ERR_INVALID_ARGUMENT;
  const bool cell = out_canvas != nullptr;   // perm without a canvas: the order does not matter for a global max
  TcArgs a{};
  a.pts = points; a.B = B; a.N = N; a.C = C;
  a.tc = reinterpret_cast<const uint8_t*>(tc_params);
  a.out_global = out_global;
  a.perm = perm; a.offsets = offsets; a.n_cells = n_cells; a.out_canvas = out_canvas;
  a.tiles_per_frame = ceil_div(N, kTileM);
  a.total_tiles = (long long)B * a.tiles_per_frame;
  if (out_global) B200BEV_CUDA_TRY(cudaMemsetAsync(out_global, 0, (size_t)B * 1024 * sizeof(float), st));
  if (out_canvas) B200BEV_CUDA_TRY(cudaMemsetAsync(out_canvas, 0, (size_t)B * n_cells * 1024 * sizeof(float), st));
  // CTA pairs (cta_group::2) when asked for and there are at least two tiles
  int cluster = tc_cluster_size();
  if (a.total_tiles < 2) cluster = 1;
  a.cluster = cluster;
  const size_t smem = tc_smem_bytes(cell, cluster);
  // debug timeline: B200BEV_TC_TRACE=<file> dumps CTA 0's clock stamps after a (synchronous) launch
  if (const char* dbg = debug_env("B200BEV_TC_DEBUG")) a.debug = atoi(dbg);
  const char* trace_path = debug_env("B200BEV_TC_TRACE");
  unsigned long long* d_trace = nullptr;
  if (trace_path) {
    B200BEV_CUDA_TRY(cudaMalloc(&d_trace, kTraceLen * sizeof(unsigned long long)));
    B200BEV_CUDA_TRY(cudaMemsetAsync(d_trace, 0, kTraceLen * sizeof(unsigned long long), st));
    a.trace = d_trace;
  }
  long long grid = (sm_count() / cluster) * cluster;
  const long long need = ceil_div64(a.total_tiles, cluster) * cluster;
  if (grid > need) grid = need;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(kTcThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // the clock-stamp instrumentation is compiled out of the production instantiations
  void (*kern)(TcArgs);
  if (cluster == 2)
    kern = cell ? (d_trace ? pointnet_mlp_tc_kernel<true, true, 2> : pointnet_mlp_tc_kernel<true, false, 2>)
                : (d_trace ? pointnet_mlp_tc_kernel<false, true, 2> : pointnet_mlp_tc_kernel<false, false, 2>);
  else
    kern = cell ? (d_trace ? pointnet_mlp_tc_kernel<true, true, 1> : pointnet_mlp_tc_kernel<true, false, 1>)
                : (d_trace ? pointnet_mlp_tc_kernel<false, true, 1> : pointnet_mlp_tc_kernel<false, false, 1>);
  B200BEV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B200BEV_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
  if (d_trace) {
    B200BEV_CUDA_TRY(cudaStreamSynchronize(st));
    unsigned long long* h = (unsigned long long*)malloc(kTraceLen * sizeof(unsigned long long));
    B200BEV_CUDA_TRY(cudaMemcpy(h, d_trace, kTraceLen * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (FILE* fp = fopen(trace_path, "w")) {
      for (int i = 0; i < kTraceLen; ++i)
        if (h[i]) fprintf(fp, "%d %llx %llu\n", i < kTraceEpi ? 0 : 1, h[i] >> 48, h[i] & 0xffffffffffffull);
      fclose(fp);
    }
    free(h);
    cudaFree(d_trace);
  }
  return launch_status();
}

}  // namespace b200bev

using namespace b200bev;

extern "C" B200BEV_API size_t b200bev_pointnet_pack_bf16_bytes(const int32_t* dims, int n_layers) {
  if (!tc_dims_supported(dims, n_layers)) return 0;
  return tc_blob_bytes(dims[0]);
}

extern "C" B200BEV_API int b200bev_pointnet_pack_bf16(const float* params, const int32_t* dims, int n_layers, void* tc_params,
                                                      size_t tc_bytes, void* stream) {
  if (!params || !tc_params) return B200BEV_ERR_INVALID_ARGUMENT;
  if (!tc_dims_supported(dims, n_layers)) return B200BEV_ERR_UNSUPPORTED;
  if (tc_bytes < tc_blob_bytes(dims[0]) || (reinterpret_cast<uintptr_t>(tc_params) & 15) != 0) return B200BEV_ERR_WORKSPACE;
  const int total = kImageStages * 128 * 8;
  pack_bf16_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(params, dims[0], reinterpret_cast<uint8_t*>(tc_params));
  return launch_status();
}
