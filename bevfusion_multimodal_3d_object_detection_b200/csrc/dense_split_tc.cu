// N2 — a dense layer on a small batch at FP32 ACCURACY on the tensor cores: the 164 MB `lidar_init.2` of
// FlexibleBEVFusion (src/fusion.py:144-148, applied :258) at batch 9..64, where the FFMA streaming kernel of dense_stream.cu
// is paced by the FMA pipe (98 us at batch 32) and not by the one thing the layer has to do: read its weight once (26 us).
//
//   out[b][o] = act(sum_k x[b][k] W[o][k] + bias[o]),      parity 1e-5 of max|ref| (nn.Linear in fp32)
//
// Every fp32 product is three fp16 tensor-core products with fp32 accumulation in tensor memory (the scheme of
// pointnet_mlp_split.cu): W = (w_hi + w_lo) / 2^k per output row (pack time: max_k |w| scaled into [1,2)), x = (x_hi + x_lo) / S
// with S the power of two that brings max|x| into [2^13, 2^14).  The weight image has the SIZE of the fp32 weight (two fp16
// halves), so the tensor-core form costs no extra HBM bytes.
//
// D[o][b]: M = 128 output rows = the TMEM lanes (a warp's store of one batch column is 32 consecutive floats of `out`),
// N = batch.  The batch operand holds x_hi and x_lo as 2 Bp rows of ONE shared-memory tile, so per 16-wide k step
//     MMA 1  w_hi (128 x 16) . [x_hi ; x_lo] (2 Bp x 16)  ->  columns [0, Bp) += w_hi.x_hi,  [Bp, 2 Bp) += w_hi.x_lo
//     MMA 2  w_lo (128 x 16) . x_hi (Bp x 16)             ->  columns [0, Bp) += w_lo.x_hi
// and the epilogue adds the two column groups on the CUDA cores (accumulation chains of K/8 and K/16 instructions: the fp32
// accumulator is truncated once per instruction).  The whole batch operand (K x 2 Bp fp16) is converted once per
// CTA and stays in shared memory; the weight arrives as 32 KB stages {w_hi, w_lo} of one 64-wide k block through a ring of 1-D
// bulk copies, in the order a CTA uses them — a tile's stages are one contiguous 32 KB x K/64 run of the image and the grid
// walks the image front to back.  Warp-specialised: 1 copy warp, 1 MMA warp, 4 epilogue warps; accumulator double-buffered.
#include <cuda_fp16.h>

#include "async_copy.cuh"
#include "common.cuh"

namespace b200bev {
namespace {

constexpr int kRows = 128;                 // output rows per tile = MMA M
constexpr int kKB = 64;                    // k per stage: 128-byte fp16 rows
constexpr int kWHalf = kRows * kKB * 2;    // 16 KB
constexpr int kWStage = 2 * kWHalf;        // 32 KB
constexpr int kThreads = 6 * 32;           // warp 0: weight copies, warp 1: MMA issue, warps 2..5: epilogue
constexpr int kConvThreads = kThreads - 32;   // warps 1..5 convert the batch operand while warp 0 already streams weights
constexpr int kMaxRing = 6;
constexpr int kMaxBatch = 64;              // rows of x per launch (2 x 64 = 128 accumulator columns per buffer)
constexpr int kTmemCols = 256;
constexpr size_t kSmemMax = 227 * 1024;

inline bool shape_ok(int O, int K) { return O > 0 && K > 0 && O % kRows == 0 && K % kKB == 0; }
inline size_t image_bytes(int O, int K) { return (size_t)(O / kRows) * (K / kKB) * kWStage + 2 * (size_t)O * sizeof(float); }
__host__ __device__ inline int padded_batch(int B) { return B <= 16 ? 16 : (B <= 32 ? 32 : 64); }
inline size_t x_bytes(int Bp, int K) { return (size_t)(K / kKB) * 2 * Bp * 128; }
inline int ring_depth(int Bp, int K) {
  const size_t fixed = x_bytes(Bp, K) + 1024 /*alignment*/ + 256 /*barriers*/;
  if (fixed + 2 * kWStage > kSmemMax) return 0;
  const size_t n = (kSmemMax - fixed) / kWStage;
  return (int)(n > kMaxRing ? kMaxRing : n);
}

// One warp per output row: 2^k that brings the row's max |w| into [1,2) (1 for an all-zero or non-finite row), the row scaled,
// split into fp16 hi / lo and stored the way SWIZZLE_128B reads a K-major tile; inverse scale and bias go to the tail.
__global__ void __launch_bounds__(256) dense_split_pack_kernel(const float* __restrict__ w, const float* __restrict__ bias, int O,
                                                               int K, uint8_t* __restrict__ img) {
  const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= O) return;
  const float* src = w + (size_t)row * K;
  float m = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + k));
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, d));
  float inv = 1.f;
  if (m > 0.f && isfinite(m)) {
    int ex = 0;
    frexpf(m, &ex);                        // m = f * 2^ex, f in [0.5, 1): 2^(1-ex) brings it into [1, 2)
    inv = ldexpf(1.f, ex - 1);
  }
  const float sc = 1.f / inv;              // exact: a power of two
  const int nkb = K / kKB, ot = row / kRows, r = row % kRows;
  for (int c = lane; c < K / 8; c += 32) {
    const int kb = c >> 3, ch = c & 7;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + c * 8)), b = __ldg(reinterpret_cast<const float4*>(src + c * 8 + 4));
    const float v[8] = {a.x * sc, a.y * sc, a.z * sc, a.w * sc, b.x * sc, b.y * sc, b.z * sc, b.w * sc};
    __align__(16) __half hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hi[j] = __float2half_rn(v[j]);
      lo[j] = __float2half_rn(v[j] - __half2float(hi[j]));
    }
    uint8_t* st = img + ((size_t)ot * nkb + kb) * kWStage + r * 128 + ((ch ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(st) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(st + kWHalf) = *reinterpret_cast<const uint4*>(lo);
  }
  if (lane == 0) {
    float* tail = reinterpret_cast<float*>(img + (size_t)(O / kRows) * nkb * kWStage);
    tail[row] = inv;
    tail[O + row] = bias ? bias[row] : 0.f;
  }
}

struct DArgs {
  const float* x;        // (B, K) f32
  const uint8_t* img;    // weight image
  float* out;            // (B, O) f32
  int B, Bp, K, O, relu, ring;
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ uint64_t sw128_kmajor(uint32_t saddr) {   // K-major, 128-byte swizzle, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void conv_group_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory"); }
#define DENSE_LD16(r, taddr)                                                                                                \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"       \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                    \
               : "r"(taddr)                                                                                                  \
               : "memory")

__global__ void __launch_bounds__(kThreads, 1) dense_split_tc_kernel(DArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int nkb = a.K / kKB, Bp = a.Bp;
  const uint32_t x_tile = (uint32_t)(2 * Bp * 128);            // bytes of the batch operand per k block: x_hi rows, then x_lo rows
  uint8_t* xt = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  uint8_t* wring = xt + (size_t)nkb * x_tile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wring + (size_t)a.ring * kWStage);
  uint64_t* full_w = bars;                         // [ring] bulk-copy bytes
  uint64_t* empty_w = bars + kMaxRing;             // [ring] tcgen05.commit
  uint64_t* acc_full = bars + 2 * kMaxRing;        // [2]
  uint64_t* acc_empty = acc_full + 2;              // [2] one arrival per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* red = reinterpret_cast<float*>(tmem_slot + 2);        // [6] block reduction of max|x|

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < a.ring; ++s) {
      mbarrier_init(&full_w[s], 1);
      mbarrier_init(&empty_w[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbarrier_init(&acc_full[s], 1);
      mbarrier_init(&acc_empty[s], 4);
    }
    mbarrier_init_fence();
  }
  __syncthreads();                                  // barriers initialised
  const int n_tiles = a.O / kRows;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = my_tiles * nkb;
  float S = 1.f;
  uint32_t tmem = 0;

  if (warp == 0) {
    // ---- weight stages in the order the MMAs use them: one 32 KB bulk copy each.  The first `ring` stages are on their way
    // while the other five warps still convert the batch operand; every byte of the image is read exactly once, so its lines
    // are marked evict-first in L2 (164 MB would otherwise push the step's activations out). ----
    const uint64_t policy = l2_evict_first_policy();
    int tile = blockIdx.x, kb = 0;
    for (int g = 0; g < total; ++g) {
      const uint32_t slot = (uint32_t)g % (uint32_t)a.ring;
      if (g >= a.ring) mbarrier_wait(&empty_w[slot], (((uint32_t)g / (uint32_t)a.ring) - 1) & 1);
      if (elect_one()) {
        mbarrier_expect_tx(&full_w[slot], kWStage);
        bulk_copy_global_to_shared_hint(wring + slot * kWStage, a.img + ((size_t)tile * nkb + kb) * kWStage, kWStage, &full_w[slot],
                                        policy);
      }
      __syncwarp();
      if (++kb == nkb) {
        kb = 0;
        tile += gridDim.x;
      }
    }
  } else {
    // ---- the batch operand: max|x| -> S, then every 8-float chunk scaled, split and stored swizzled (warps 1..5) ----
    const int t5 = tid - 32;
    const int n4 = a.B * a.K / 4;
    float m = 0.f;
    for (int i = t5; i < n4; i += kConvThreads) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.x) + i);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, d));
    if (lane == 0) red[warp] = m;
    tc_fence_before();
    conv_group_sync();
    tc_fence_after();
    tmem = *tmem_slot;
    m = fmaxf(fmaxf(fmaxf(red[1], red[2]), fmaxf(red[3], red[4])), red[5]);
    S = activation_scale(__float_as_uint(m));
    const int chunks_per_row = a.K / 8;
    for (int i = t5; i < Bp * chunks_per_row; i += kConvThreads) {
      const int b = i / chunks_per_row, c = i - b * chunks_per_row, kb = c >> 3, ch = c & 7;
      uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
      if (b < a.B) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)b * a.K + c * 8));
        const float4 q = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)b * a.K + c * 8 + 4));
        const float v[8] = {p.x * S, p.y * S, p.z * S, p.w * S, q.x * S, q.y * S, q.z * S, q.w * S};
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __half2 hh = __floats2half2_rn(v[2 * e], v[2 * e + 1]);            // the even k in the low half
          const float2 back = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(v[2 * e] - back.x, v[2 * e + 1] - back.y);
          h[e] = *reinterpret_cast<const uint32_t*>(&hh);
          l[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        hi = make_uint4(h[0], h[1], h[2], h[3]);
        lo = make_uint4(l[0], l[1], l[2], l[3]);
      }
      uint8_t* dst = xt + (size_t)kb * x_tile + b * 128 + ((ch ^ (b & 7)) << 4);   // Bp % 8 == 0: row Bp + b swizzles like row b
      *reinterpret_cast<uint4*>(dst) = hi;
      *reinterpret_cast<uint4*>(dst + Bp * 128) = lo;
    }
    fence_proxy_async_shared();          // generic-proxy stores ordered before the tensor core's asynchronous-proxy reads
    conv_group_sync();
  }

  if (warp == 1) {
    // ---- MMA issuer.  Instruction descriptor: D f32 (bit 4), A and B fp16 (format 0), both K-major, N, M = 128 ----
    const uint32_t idesc_hi = (1u << 4) | ((uint32_t)((2 * Bp) >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
    const uint32_t idesc_lo = (1u << 4) | ((uint32_t)(Bp >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
    int kb = 0, tile_seq = 0;
    for (int c = 0; c < total; ++c) {
      const uint32_t ws = (uint32_t)c % (uint32_t)a.ring, buf = tile_seq & 1;
      if (kb == 0 && tile_seq >= 2) mbarrier_wait(&acc_empty[buf], ((tile_seq >> 1) - 1) & 1);
      mbarrier_wait(&full_w[ws], ((uint32_t)c / (uint32_t)a.ring) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t w_hi = smem_addr(wring + ws * kWStage), w_lo = w_hi + kWHalf;
        const uint32_t xs = smem_addr(xt) + (uint32_t)kb * x_tile;
        const uint32_t d = tmem + buf * (uint32_t)(2 * Bp);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          mma_f16_ss(d, sw128_kmajor(w_hi + s * 32), sw128_kmajor(xs + s * 32), idesc_hi, !(kb == 0 && s == 0));
          mma_f16_ss(d, sw128_kmajor(w_lo + s * 32), sw128_kmajor(xs + s * 32), idesc_lo, 1u);
        }
        tc_commit(&empty_w[ws]);
        if (kb == nkb - 1) tc_commit(&acc_full[buf]);
      }
      __syncwarp();
      if (++kb == nkb) {
        kb = 0;
        ++tile_seq;
      }
    }
  } else if (warp >= 2) {
    // ---- epilogue: warp w owns TMEM lanes [32 (w % 4), +32) = 32 output rows; a lane is one output row ----
    const int quad = warp & 3;
    const float inv_S = 1.f / S;                                   // exact: a power of two
    const float* tail = reinterpret_cast<const float*>(a.img + (size_t)n_tiles * nkb * kWStage);
    int tile = blockIdx.x;
    for (int tile_seq = 0; tile_seq < my_tiles; ++tile_seq, tile += gridDim.x) {
      const uint32_t buf = tile_seq & 1;
      const int o = tile * kRows + quad * 32 + lane;
      const float unscale = __ldg(tail + o) * inv_S, bias = __ldg(tail + a.O + o);
      mbarrier_wait(&acc_full[buf], (tile_seq >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem + ((uint32_t)(quad * 32) << 16) + buf * (uint32_t)(2 * Bp);
#pragma unroll 1
      for (int c0 = 0; c0 < Bp; c0 += 16) {
        if (c0 >= a.B) break;                                      // warp-uniform: padding columns are never stored
        uint32_t p[16], q[16];
        DENSE_LD16(p, tbase + (uint32_t)c0);
        DENSE_LD16(q, tbase + (uint32_t)(Bp + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* dst = a.out + (size_t)c0 * a.O + o;                 // a store instruction: 32 output rows of one batch row, 128 B
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (c0 + j < a.B) {
            float v = fmaf(__uint_as_float(p[j]) + __uint_as_float(q[j]), unscale, bias);
            if (a.relu) v = fmaxf(v, 0.f);
            dst[(size_t)j * a.O] = v;
          }
        }
      }
      tc_fence_before();     // the tensor-memory loads above are complete (wait::ld) before the buffer is handed back
      __syncwarp();
      if (lane == 0) mbarrier_arrive(&acc_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
  }
}

}  // namespace

int dense_layer(const float* x, int B, int K, const float* w, const float* bias, int O, int relu, float* out, cudaStream_t st);

int dense_layer_split(const float* x, int B, int K, const void* image, int O, int relu, float* out, cudaStream_t st) {
  if (!shape_ok(O, K)) return B200BEV_ERR_UNSUPPORTED;
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(image)) & 15) != 0) return B200BEV_ERR_INVALID_ARGUMENT;
  for (int b0 = 0; b0 < B; b0 += kMaxBatch) {
    const int nb = B - b0 < kMaxBatch ? B - b0 : kMaxBatch;
    DArgs a{x + (size_t)b0 * K, reinterpret_cast<const uint8_t*>(image), out + (size_t)b0 * O, nb, padded_batch(nb), K, O, relu, 0};
    a.ring = ring_depth(a.Bp, K);
    if (a.ring < 2) return B200BEV_ERR_UNSUPPORTED;
    const size_t smem = x_bytes(a.Bp, K) + (size_t)a.ring * kWStage + 1024 + 256;
    B200BEV_CUDA_TRY(cudaFuncSetAttribute(dense_split_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int gx = sm_count();
    if (gx > O / kRows) gx = O / kRows;
    dense_split_tc_kernel<<<gx, kThreads, smem, st>>>(a);
    const int rc = launch_status();
    if (rc) return rc;
  }
  return B200BEV_OK;
}

}  // namespace b200bev

using namespace b200bev;

extern "C" B200BEV_API size_t b200bev_dense_pack_split_bytes(int O, int K) { return shape_ok(O, K) ? image_bytes(O, K) : 0; }

extern "C" B200BEV_API int b200bev_dense_pack_split(const float* weight, const float* bias, int O, int K, void* image,
                                                    size_t image_size, void* stream) {
  if (!weight || !image) return B200BEV_ERR_INVALID_ARGUMENT;
  if (!shape_ok(O, K)) return B200BEV_ERR_UNSUPPORTED;
  if (image_size < image_bytes(O, K)) return B200BEV_ERR_WORKSPACE;
  if (((reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(image)) & 15) != 0) return B200BEV_ERR_INVALID_ARGUMENT;
  dense_split_pack_kernel<<<(O + 7) / 8, 256, 0, (cudaStream_t)stream>>>(weight, bias, O, K, reinterpret_cast<uint8_t*>(image));
  return launch_status();
}

extern "C" B200BEV_API int b200bev_dense_layer_split(const float* x, int B, int K, const void* image, int O, int relu, float* out,
                                                     void* stream) {
  if (!x || !image || !out || B <= 0 || K <= 0 || O <= 0) return B200BEV_ERR_INVALID_ARGUMENT;
  return dense_layer_split(x, B, K, image, O, relu, out, (cudaStream_t)stream);
}

extern "C" B200BEV_API int b200bev_lidar_init_split(const float* lidar_features, int B, int K, const float* w1, const float* b1,
                                                    int hidden, const void* image2, int O, float* hidden_ws, float* out,
                                                    void* stream) {
  if (!lidar_features || !w1 || !image2 || !hidden_ws || !out || B <= 0 || K <= 0 || hidden <= 0 || O <= 0)
    return B200BEV_ERR_INVALID_ARGUMENT;
  if (!shape_ok(O, hidden)) return B200BEV_ERR_UNSUPPORTED;
  const int rc = dense_layer(lidar_features, B, K, w1, b1, hidden, 1, hidden_ws, (cudaStream_t)stream);
  if (rc) return rc;
  return dense_layer_split(hidden_ws, B, hidden, image2, O, 0, out, (cudaStream_t)stream);
}
