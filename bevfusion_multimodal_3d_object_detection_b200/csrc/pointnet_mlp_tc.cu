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
//   warps 2..5  epilogue: layer 1 (K = 4) on CUDA cores, layers 2-4 TMEM -> bias/ReLU/bf16 -> TMEM,
//               layer 5: max over the 128 points of the tile.  Points are TMEM lanes, so this is a
//               cross-lane reduction: a 31-shuffle transposing butterfly per 32 channels leaves lane l
//               with channel l's maximum; it is folded into 32 running-max registers per thread and
//               flushed once per frame (bias + ReLU after the max: both are monotone per channel).
//
// TMEM columns (512): [0,256) act4 (act1 at [0,32) and act2 at [64,128) reuse it earlier in the
// tile), [256,384) act3, which becomes accumulator 0 for layer 5, [384,512) accumulator 1.
#include <cuda_bf16.h>

#include "common.cuh"

namespace b200bev {
namespace {

constexpr int kTileM = 128;
constexpr int kStageBytes = 16384;  // 128 rows x 64 bf16
constexpr int kStages = 12;
constexpr int kTcThreads = 192;
constexpr int kStagesPerTile = 1 + 4 + 16 + 64;
constexpr int kBiasFloats = 128 + 256 + 512 + 1024;
constexpr int kMaxCin = 16;

constexpr uint32_t kColAct1 = 0, kColAct2 = 64, kColAct4 = 0, kColAct3 = 256, kColAcc0 = 256, kColAcc1 = 384;

struct TcArgs {
  const float* pts;
  int B, N, C;
  const uint8_t* tc;  // [85 stages][W1^T (C x 64) f32][b1 64][bias L2..L5]
  float* out_global;
  int tiles_per_frame;
  long long total_tiles;
};

__host__ __device__ inline size_t tc_blob_bytes(int C) {
  return (size_t)kStagesPerTile * kStageBytes + ((size_t)C * 64 + 64 + kBiasFloats) * sizeof(float);
}

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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T, 128 x 128 x 16, bf16 -> f32
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
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

__global__ void __launch_bounds__(kTcThreads, 1) pointnet_mlp_tc_kernel(TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // SWIZZLE_128B: 1024-B aligned tiles
  float* bias_s = reinterpret_cast<float*>(ring + (size_t)kStages * kStageBytes);  // L2 | L3 | L4 | L5
  float* w1_s = bias_s + kBiasFloats;                                               // W1^T (C x 64), then b1 (64)
  uint64_t* bars = reinterpret_cast<uint64_t*>(w1_s + kMaxCin * 64 + 64);
  uint64_t* full = bars;                    // [kStages]  weights landed
  uint64_t* empty = bars + kStages;         // [kStages]  MMAs that read the stage retired
  uint64_t* acc_full = empty + kStages;     // [2]        accumulator complete
  uint64_t* acc_empty = acc_full + 2;       // [2]        accumulator drained by the epilogue
  uint64_t* act_ready = acc_empty + 2;      // [1]        next layer's A operand is in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_ready + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const float* tail = reinterpret_cast<const float*>(a.tc + (size_t)kStagesPerTile * kStageBytes);
  for (int i = tid; i < a.C * 64 + 64; i += kTcThreads) w1_s[i] = __ldg(tail + i);
  for (int i = tid; i < kBiasFloats; i += kTcThreads) bias_s[i] = __ldg(tail + a.C * 64 + 64 + i);
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], kTileM);
    mbar_init(&acc_empty[1], kTileM);
    mbar_init(act_ready, kTileM);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const long long per_cta = (a.total_tiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = per_cta * blockIdx.x;
  const long long t_end = t_begin + per_cta < a.total_tiles ? t_begin + per_cta : a.total_tiles;

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        for (int s = 0; s < kStagesPerTile; ++s) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], kStageBytes);
          bulk_copy_g2s(ring + (size_t)stage * kStageBytes, a.tc + (size_t)s * kStageBytes, kStageBytes, &full[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      // instruction descriptor: D f32 (bit 4), A bf16 (bit 7), B bf16 (bit 10), both K-major, N=128, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t stage = 0, phase = 0, act_phase = 0;
      uint32_t acc_parity = 0;  // bit b: parity of the next use of accumulator b
      for (long long t = t_begin; t < t_end; ++t) {
#pragma unroll 1
        for (int layer = 0; layer < 4; ++layer) {  // network layers 2..5
          const int kchunks = 1 << layer;          // K / 64
          const int nchunks = 1 << layer;          // N / 128
          const uint32_t a_col = layer == 0 ? kColAct1 : layer == 1 ? kColAct2 : layer == 2 ? kColAct3 : kColAct4;
          mbar_wait(act_ready, act_phase);
          act_phase ^= 1;
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < nchunks; ++c) {
            const int buf = (layer == 3) ? (c & 1) : 1;
            mbar_wait(&acc_empty[buf], ((acc_parity >> buf) & 1) ^ 1);
            acc_parity ^= 1u << buf;
            tc_fence_after();
            const uint32_t d_addr = tmem + (buf ? kColAcc1 : kColAcc0);
#pragma unroll 1
            for (int kc = 0; kc < kchunks; ++kc) {
              mbar_wait(&full[stage], phase);
              tc_fence_after();
              const uint64_t bdesc = make_b_desc(smem_u32(ring + (size_t)stage * kStageBytes));
#pragma unroll
              for (int s = 0; s < 4; ++s)
                umma_ts(d_addr, tmem + a_col + kc * 32 + s * 8, bdesc + (uint64_t)(s * 2), idesc, (kc | s) != 0);
              tc_commit(&empty[stage]);
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            tc_commit(&acc_full[buf]);
          }
        }
      }
    }
  } else {
    // ================================ epilogue (128 threads) ================================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may touch
    const int row = quad * 32 + lane;          // point within the tile = TMEM lane
    const uint32_t tm = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t full_phase[2] = {0, 0};
    float rmax[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) rmax[i] = -INFINITY;
    int cur_frame = -1;
    const int c_out = 1024;
    const float* b5 = bias_s + 128 + 256 + 512;

    auto flush = [&](int frame) {
      int* o = reinterpret_cast<int*>(a.out_global + (size_t)frame * c_out);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int col = i * 32 + lane;
        const float v = fmaxf(rmax[i] + b5[col], 0.0f);   // bias + ReLU commute with the max
        atomicMax(o + col, __float_as_int(v));
        rmax[i] = -INFINITY;
      }
    };

    for (long long t = t_begin; t < t_end; ++t) {
      const int f = (int)(t / a.tiles_per_frame);
      const int s0 = (int)(t % a.tiles_per_frame) * kTileM;
      if (f != cur_frame) {
        if (cur_frame >= 0) flush(cur_frame);
        cur_frame = f;
      }
      const int p = s0 + row;
      const bool valid = p < a.N;
      const bool partial = s0 + kTileM > a.N;

      // ---- layer 1 on CUDA cores (K = C_in is 4): act1 = relu(W1 x + b1) -> bf16 -> TMEM [0,32) ----
      {
        float x[kMaxCin];
        const float* src = a.pts + ((size_t)f * a.N + (valid ? p : 0)) * a.C;
        if (a.C == 4 && (reinterpret_cast<uintptr_t>(a.pts) & 15) == 0) {
          const float4 v = valid ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
          x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
#pragma unroll
          for (int k = 4; k < kMaxCin; ++k) x[k] = 0.0f;
        } else {
#pragma unroll
          for (int k = 0; k < kMaxCin; ++k) x[k] = (k < a.C && valid) ? __ldg(src + k) : 0.0f;
        }
        uint32_t packed[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v0 = w1_s[a.C * 64 + 2 * j], v1 = w1_s[a.C * 64 + 2 * j + 1];
#pragma unroll
          for (int k = 0; k < kMaxCin; ++k) {
            if (k < a.C) {
              v0 = fmaf(w1_s[k * 64 + 2 * j], x[k], v0);
              v1 = fmaf(w1_s[k * 64 + 2 * j + 1], x[k], v1);
            }
          }
          packed[j] = pack_bf16x2(fmaxf(v0, 0.0f), fmaxf(v1, 0.0f));
        }
        TC_ST16(tm + kColAct1, packed, 0);
        TC_ST16(tm + kColAct1 + 16, packed, 16);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(act_ready);
      }

      // ---- layers 2..4: accumulator -> bias, ReLU, bf16 -> next A operand in TMEM ----
#pragma unroll 1
      for (int layer = 0; layer < 3; ++layer) {
        const int nchunks = 1 << layer;
        const uint32_t out_col = layer == 0 ? kColAct2 : layer == 1 ? kColAct3 : kColAct4;
        const float* bl = bias_s + (layer == 0 ? 0 : layer == 1 ? 128 : 384);
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&acc_full[1], full_phase[1]);
          full_phase[1] ^= 1;
          tc_fence_after();
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {
            uint32_t r[32];
            TC_LD32(r, tm + kColAcc1 + q * 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint32_t packed[16];
            const float* bq = bl + c * 128 + q * 32;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float v0 = fmaxf(__uint_as_float(r[2 * j]) + bq[2 * j], 0.0f);
              const float v1 = fmaxf(__uint_as_float(r[2 * j + 1]) + bq[2 * j + 1], 0.0f);
              packed[j] = pack_bf16x2(v0, v1);
            }
            TC_ST16(tm + out_col + c * 64 + q * 16, packed, 0);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          mbar_arrive(&acc_empty[1]);
          if (c == nchunks - 1) mbar_arrive(act_ready);
        }
      }

      // ---- layer 5: max over the tile's points, channel chunk by channel chunk ----
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int buf = c & 1;
        mbar_wait(&acc_full[buf], full_phase[buf]);
        full_phase[buf] ^= 1;
        tc_fence_after();
        const uint32_t acc_col = buf ? kColAcc1 : kColAcc0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t r[32];
          TC_LD32(r, tm + acc_col + g * 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (g == 3) {  // the whole chunk is in registers: hand the accumulator back to the MMA warp
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
          }
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
          rmax[c * 4 + g] = fmaxf(rmax[c * 4 + g], v[0]);  // lane l holds channel c*128 + g*32 + l
        }
      }
    }
    if (cur_frame >= 0) flush(cur_frame);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// Re-tiles the fp32 blob (W^T (K x N) + bias per layer) into the stage image: stage -> (layer, 128-channel
// block, 64-k block); row r of a stage holds 64 bf16 of W[n0 + r][k0 ..], its eight 16-byte chunks XOR-
// swizzled with (r & 7) exactly as SWIZZLE_128B reads them.
__global__ void __launch_bounds__(256) pack_bf16_kernel(const float* __restrict__ params, int C, uint8_t* __restrict__ tc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (stage, row, chunk)
  const int total = kStagesPerTile * 128 * 8;
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
    const int s = idx / (128 * 8), r = (idx / 8) % 128, ch = idx % 8;
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
    *reinterpret_cast<uint4*>(tc + (size_t)s * kStageBytes + r * 128 + ((ch ^ (r & 7)) * 16)) = *reinterpret_cast<uint4*>(v);
  }
  // tail: W1^T, b1, then the biases of layers 2..5, fp32
  float* tailp = reinterpret_cast<float*>(tc + (size_t)kStagesPerTile * kStageBytes);
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

size_t tc_smem_bytes() {
  return 1024 + (size_t)kStages * kStageBytes + (kBiasFloats + kMaxCin * 64 + 64) * sizeof(float) +
         (2 * kStages + 5) * sizeof(uint64_t) + 16;
}

}  // namespace

int pointnet_encode_tc(const float* points, int B, int N, int C, const float* params, const int32_t* dims, int n_layers,
                       const int32_t* perm, const int32_t* offsets, int n_cells, const void* tc_params, float* out_global,
                       float* out_canvas, cudaStream_t st) {
  (void)params; (void)offsets; (void)n_cells;
  if (!tc_dims_supported(dims, n_layers) || dims[0] != C) return B200BEV_ERR_UNSUPPORTED;
  if (perm || out_canvas || !out_global) return B200BEV_ERR_UNSUPPORTED;  // per-cell canvas: fp32 path for now
  if ((reinterpret_cast<uintptr_t>(tc_params) & 15) != 0) return B200BEV_ERR_INVALID_ARGUMENT;
  TcArgs a{};
  a.pts = points; a.B = B; a.N = N; a.C = C;
  a.tc = reinterpret_cast<const uint8_t*>(tc_params);
  a.out_global = out_global;
  a.tiles_per_frame = ceil_div(N, kTileM);
  a.total_tiles = (long long)B * a.tiles_per_frame;
  const size_t smem = tc_smem_bytes();
  B200BEV_CUDA_TRY(cudaFuncSetAttribute(pointnet_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B200BEV_CUDA_TRY(cudaMemsetAsync(out_global, 0, (size_t)B * 1024 * sizeof(float), st));
  long long grid = sm_count();
  if (grid > a.total_tiles) grid = a.total_tiles;
  pointnet_mlp_tc_kernel<<<(int)grid, kTcThreads, smem, st>>>(a);
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
  const int total = kStagesPerTile * 128 * 8;
  pack_bf16_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(params, dims[0], reinterpret_cast<uint8_t*>(tc_params));
  return launch_status();
}
