// N2 — the dense layers around the BEV canvas on a small batch: out = act(x W^T + b).
//
// Replaces, in eval mode: FlexibleBEVFusion.lidar_init = Linear(1024,512) + ReLU + Linear(512, 128*25*25)
// (src/fusion.py:144-148, applied at :258) and radar_proj = Linear(256,256) + ReLU (src/fusion.py:170-173, :274).
// The second lidar_init layer is the largest weight read of a forward pass: (80000, 512) fp32 = 164 MB for 2.6 GFLOP
// at batch 32 — 16 FLOP per weight byte, so the layer is bound by streaming the weights from HBM, not by the math
// (SURVEY 8a A5, 8f N2).  fp32 FFMA throughout (parity 1e-5; one TF32 pass misses it).
//
// linear_stream_kernel<NB>: one CTA per SM, each owning a contiguous range of output rows (O/grid, so every CTA walks
// the same number of 64-row steps).  The batch tile (TB = 4*NB rows of x, K floats each) stays in shared memory for the
// whole kernel.  The 8 warps are 2 row groups x 4 K-quarters: a warp multiplies 32 weight rows x its quarter of K (every fourth 32-float chunk) against
// the whole batch tile, streaming its weights with its OWN cp.async ring (16-byte LDGSTS, 3-5 stages of 32 rows x 32 k,
// rows padded to 144 B so the eight row addresses of a 128-bit shared load fall in eight different bank groups) — no
// block-wide barrier in the streaming loop, and the ring keeps running across step boundaries.  A thread holds a
// 4 (rows) x NB (batch) register tile and reads both operands as float4 along k: 4 + NB shared loads per 16*NB FMAs.
// At the end of a step the four K-quarter partials meet in shared memory, bias and ReLU are applied, and each output
// value is written once (64 consecutive floats of one batch row per warp-pair: 256-byte segments).
//
// linear_rows_kernel: the any-shape form (one warp per output row, weights read once, x from L2) — used for the small
// layers (lidar_init.0: 2 MB of weights; radar_proj: 256 KB) and whenever the streaming kernel's shape rules fail.
#include <cstdlib>

#include "common.cuh"

namespace b200bev {
namespace {

constexpr int kThreads = 256;
constexpr int kKC = 32;           // k floats per ring row
constexpr int kRowF = kKC + 4;    // padded ring row (floats)
constexpr int kStepRows = 64;     // output rows per CTA step: 2 groups of 32
constexpr int kKSplit = 4;

struct LinArgs {
  const float* x;     // (B, K)
  const float* w;     // (O, K)
  const float* bias;  // (O) or null
  float* out;         // (B, O)
  int B, K, O, relu;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool real) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = real ? 16 : 0;  // 0: nothing is read, the 16 bytes are zero-filled
  // L2::256B: the 128-byte piece of a row this chunk needs arrives with its neighbour, which is the piece the next chunk
  // of the same rows needs — DRAM sees 256-byte bursts instead of isolated 128-byte ones
  asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// NB: batch rows per thread (tile of 4*NB rows); kStages: ring depth (kStages-1 chunks in flight per warp)
template <int NB, int kStages>
size_t stream_smem_bytes(int K) {
  const int TB = 4 * NB;
  return sizeof(float) * ((size_t)TB * (K + 4) + 8 * kStages * 32 * kRowF + 8 * TB * 33);
}

template <int NB, int kStages>
__global__ void __launch_bounds__(kThreads, 1) linear_stream_kernel(LinArgs a) {
  constexpr int TB = 4 * NB;
  extern __shared__ __align__(16) float smem[];
  const int K = a.K, xs_stride = K + 4;
  float* xs = smem;                                  // [TB][K+4]
  float* ring = xs + (size_t)TB * xs_stride;         // [8 warps][kStages][32][kRowF]
  float* red = ring + 8 * kStages * 32 * kRowF;      // [8 warps][TB][33]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int og = warp & 1, kq = warp >> 1;
  const int go = lane >> 2, gb = lane & 3;
  const int b0 = blockIdx.y * TB;

  // this CTA's rows
  const long long O = a.O;
  const int row_lo = (int)(O * blockIdx.x / gridDim.x), row_hi = (int)(O * (blockIdx.x + 1) / gridDim.x);
  const int n_steps = ceil_div(row_hi - row_lo, kStepRows);
  const int kq_len = K / kKSplit, nch = kq_len / kKC;
  const int total = n_steps * nch;  // chunks this warp streams

  float* my_ring = ring + (size_t)warp * kStages * 32 * kRowF;
  auto issue = [&](int q) {
    if (q < total) {
      const int step = q / nch, ch = q - step * nch;
      const int r0 = row_lo + step * kStepRows + og * 32;
      float* dst = my_ring + (size_t)(q % kStages) * 32 * kRowF;
      const int k0 = (ch * kKSplit + kq) * kKC;   // the four K-split warps of a row group read four ADJACENT 128-byte pieces
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int p = lane + 32 * t, r = p >> 3, c = p & 7;
        const int row = r0 + r;
        const bool real = row < row_hi;
        cp_async16(dst + r * kRowF + c * 4, a.w + (size_t)(real ? row : row_lo) * K + k0 + c * 4, real);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) issue(s);

  // batch tile -> shared memory (zero rows past B)
  for (int i = tid; i < TB * (K / 4); i += kThreads) {
    const int b = i / (K / 4), c = i - b * (K / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b0 + b < a.B) v = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)(b0 + b) * K) + c);
    *reinterpret_cast<float4*>(xs + (size_t)b * xs_stride + c * 4) = v;
  }
  __syncthreads();

  int q = 0;
  for (int step = 0; step < n_steps; ++step) {
    float acc[4][NB];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j) acc[i][j] = 0.f;

    for (int ch = 0; ch < nch; ++ch, ++q) {
      cp_async_wait<kStages - 2>();
      __syncwarp();
      issue(q + kStages - 1);  // refills the slot consumed in the previous iteration
      const float* wt = my_ring + (size_t)(q % kStages) * 32 * kRowF + go * kRowF;
      const float* xt = xs + (size_t)gb * xs_stride + (ch * kKSplit + kq) * kKC;
      // operands of the next four k are fetched while the FMAs of these four run: with two warps per scheduler the
      // shared-memory latency is not hidden by other warps
      float4 wv[2][4], xv[2][NB];
      auto fetch = [&](int set, int kk) {
#pragma unroll
        for (int i = 0; i < 4; ++i) wv[set][i] = *reinterpret_cast<const float4*>(wt + i * 8 * kRowF + kk);
#pragma unroll
        for (int j = 0; j < NB; ++j) xv[set][j] = *reinterpret_cast<const float4*>(xt + (size_t)j * 4 * xs_stride + kk);
      };
      fetch(0, 0);
#pragma unroll
      for (int kk = 0; kk < kKC; kk += 4) {
        const int cur = (kk >> 2) & 1;
        if (kk + 4 < kKC) fetch(cur ^ 1, kk + 4);
        // k component outermost: 4*NB independent FMAs between two that touch the same accumulator
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j)
              acc[i][j] = fmaf(reinterpret_cast<const float*>(&wv[cur][i])[c], reinterpret_cast<const float*>(&xv[cur][j])[c],
                               acc[i][j]);
      }
    }

    // the four K-quarters of a row group meet in shared memory
    float* my_red = red + (size_t)warp * TB * 33;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j) my_red[(j * 4 + gb) * 33 + i * 8 + go] = acc[i][j];
    __syncthreads();
    {
      const int o_l = tid & 63, g = o_l >> 5, o_in = o_l & 31;
      const int row = row_lo + step * kStepRows + o_l;
      if (row < row_hi) {
        const float bias = a.bias ? __ldg(a.bias + row) : 0.f;
        for (int b = tid >> 6; b < TB; b += 4) {
          if (b0 + b >= a.B) break;
          float s = 0.f;
#pragma unroll
          for (int kk = 0; kk < kKSplit; ++kk) s += red[(size_t)(kk * 2 + g) * TB * 33 + b * 33 + o_in];
          s += bias;
          if (a.relu) s = fmaxf(s, 0.f);
          a.out[(size_t)(b0 + b) * a.O + row] = s;
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
}

// ---- row-streaming form (default for batch <= 16) -------------------------------------------------------------------------
// What the streaming kernel above taught (profiles/r01_ncu_dense.txt, tests/cuda/stream_read.cu): with 8 warps per SM a
// B200 reads HBM at 3.4-3.6 TB/s whatever is in flight per warp (deeper rings, bulk copies, contiguous 32 KB runs all gave
// the same 51 us for this 164 MB weight); 16 warps reach 4.9-5.2 TB/s, 32 warps 5.6.  So this form runs 16 warps per SM
// and makes room for them by keeping the batch tile in REGISTERS: a lane keeps its K-slice of every batch row
// (KPL * BT <= 64 floats), the only shared-memory traffic is one 128-bit load per 4*BT FMAs, and a row is shared by
// KS = K / (32 * KPL) warps whose partial sums meet in shared memory once per chunk round.  The weights arrive as whole
// row slices through per-warp cp.async rings (4 slots of 2 KB); chunk round q of CTA x covers rows
// [(q * grid + x) * STEP, +STEP), so the grid reads one contiguous window at a time, like a grid-stride copy.  The K-sum
// over the 32 lanes is a transposing butterfly (BT-1 shuffles leave lane b with batch row b's sum).
constexpr int kRsWarps = 16;
constexpr int kRsThreads = kRsWarps * 32;
constexpr int kRsSlots = 4;
constexpr int kRsChunkBytes = 2048;

template <int HALF>
__device__ __forceinline__ void add_butterfly_level(float* v, int lane) {
  const bool up = (lane & HALF) != 0;
#pragma unroll
  for (int j = 0; j < HALF; ++j) {
    const float send = up ? v[j] : v[j + HALF];
    const float keep = up ? v[j + HALF] : v[j];
    v[j] = keep + __shfl_xor_sync(FULL_MASK, send, HALF);
  }
}
// v[0..BT) per lane -> v[0] = sum over all 32 lanes of v[lane & (BT-1)]
template <int BT>
__device__ __forceinline__ float warp_sum_transposed(float* v, int lane) {
  if constexpr (BT >= 16) add_butterfly_level<8>(v, lane);
  if constexpr (BT >= 8) add_butterfly_level<4>(v, lane);
  add_butterfly_level<2>(v, lane);
  add_butterfly_level<1>(v, lane);
  float s = v[0];
#pragma unroll
  for (int m = BT; m < 32; m <<= 1) s += __shfl_xor_sync(FULL_MASK, s, m);
  return s;
}

template <int BT, int KPL, int KS>   // batch tile; K floats per lane; warps per row: K = 32 * KPL * KS
__global__ void __launch_bounds__(kRsThreads, 1) linear_rowstream_kernel(LinArgs a) {
  constexpr int RL = kRsWarps / KS;                 // rows the CTA works on side by side
  constexpr int SLICE = 32 * KPL;                   // floats of a row one warp handles
  constexpr int RPC = kRsChunkBytes / (SLICE * 4);  // rows per chunk
  constexpr int STEP = RL * RPC;                    // rows the CTA consumes per chunk round
  extern __shared__ __align__(16) float smem[];
  float* ring = smem;                                              // [warps][kRsSlots][RPC][SLICE]
  float* red = ring + kRsWarps * kRsSlots * (kRsChunkBytes / 4);   // [2][RL][RPC][KS][BT] partial sums

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ks = warp % KS, rl = warp / KS;
  const int K = a.K, b0 = blockIdx.y * BT;
  const int row_hi = a.O;
  const int total_steps = ceil_div(a.O, STEP);
  const int n_chunks = total_steps > (int)blockIdx.x ? (total_steps - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int kw0 = ks * SLICE;
  auto step_row = [&](int q) { return (q * (int)gridDim.x + (int)blockIdx.x) * STEP; };

  float* my_ring = ring + (size_t)warp * kRsSlots * (kRsChunkBytes / 4);
  auto issue = [&](int q) {
    if (q < n_chunks) {
      const int r0 = step_row(q) + rl * RPC;
      float* dst = my_ring + (size_t)(q % kRsSlots) * (kRsChunkBytes / 4);
#pragma unroll
      for (int t = 0; t < kRsChunkBytes / 512; ++t) {
        const int p = lane + 32 * t;                  // 16-byte piece of the chunk
        const int r = p / (SLICE / 4), c = p - r * (SLICE / 4);
        const int row = r0 + r;
        const bool real = row < row_hi;
        cp_async16(dst + p * 4, a.w + (size_t)(real ? row : 0) * K + kw0 + c * 4, real);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s2 = 0; s2 < kRsSlots - 1; ++s2) issue(s2);

  // this lane's slice of the batch tile: float4 f of the slice is k = kw0 + (f * 32 + lane) * 4
  float xr[BT][KPL];
#pragma unroll
  for (int b = 0; b < BT; ++b)
#pragma unroll
    for (int f = 0; f < KPL / 4; ++f) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + b < a.B) v = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)(b0 + b) * K + kw0) + f * 32 + lane);
      xr[b][4 * f] = v.x; xr[b][4 * f + 1] = v.y; xr[b][4 * f + 2] = v.z; xr[b][4 * f + 3] = v.w;
    }

  for (int q = 0; q < n_chunks; ++q) {
    cp_async_wait<kRsSlots - 2>();
    __syncwarp();
    issue(q + kRsSlots - 1);
    const float* wt = my_ring + (size_t)(q % kRsSlots) * (kRsChunkBytes / 4);
    float* my_red = red + (size_t)(((q & 1) * RL + rl) * RPC) * KS * BT;
#pragma unroll
    for (int r = 0; r < RPC; ++r) {
      float acc[BT];
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[b] = 0.f;
#pragma unroll
      for (int f = 0; f < KPL / 4; ++f) {
        const float4 w4 = *reinterpret_cast<const float4*>(wt + r * SLICE + (f * 32 + lane) * 4);
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          acc[b] = fmaf(w4.x, xr[b][4 * f], acc[b]);
          acc[b] = fmaf(w4.y, xr[b][4 * f + 1], acc[b]);
          acc[b] = fmaf(w4.z, xr[b][4 * f + 2], acc[b]);
          acc[b] = fmaf(w4.w, xr[b][4 * f + 3], acc[b]);
        }
      }
      const float s = warp_sum_transposed<BT>(acc, lane);   // lane l holds batch row l & (BT-1)
      if (lane < BT) my_red[(r * KS + ks) * BT + lane] = s;
    }
    __syncthreads();   // the partial sums of every row of this chunk round are in shared memory (red is double-buffered)
    for (int i = tid; i < STEP * BT; i += kRsThreads) {
      const int rr = i % STEP, b = i / STEP;      // consecutive threads: consecutive output rows of one batch row
      const int row = step_row(q) + rr;
      if (row < row_hi && b0 + b < a.B) {
        const float* pr = red + (size_t)((q & 1) * RL * RPC + rr) * KS * BT + b;
        float v = 0.f;
#pragma unroll
        for (int k2 = 0; k2 < KS; ++k2) v += pr[k2 * BT];
        v += a.bias ? __ldg(a.bias + row) : 0.f;
        if (a.relu) v = fmaxf(v, 0.f);
        a.out[(size_t)(b0 + b) * a.O + row] = v;
      }
    }
  }
  cp_async_wait<0>();
}

template <int BT, int KPL, int KS>
int launch_rowstream(const LinArgs& a, cudaStream_t st) {
  constexpr int RL = kRsWarps / KS, RPC = kRsChunkBytes / (32 * KPL * 4);
  const size_t smem = (size_t)kRsWarps * kRsSlots * kRsChunkBytes + sizeof(float) * 2 * RL * RPC * KS * BT;
  auto kern = linear_rowstream_kernel<BT, KPL, KS>;
  B200BEV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int gx = sm_count();
  const int max_gx = ceil_div(a.O, RL * RPC);
  if (gx > max_gx) gx = max_gx;
  kern<<<dim3(gx, ceil_div(a.B, BT)), kRsThreads, smem, st>>>(a);
  return launch_status();
}

// any shape: warp per output row, 8 batch rows per pass
__global__ void __launch_bounds__(256) linear_rows_kernel(LinArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int o = warp; o < a.O; o += n_warps) {
    const float* wr = a.w + (size_t)o * a.K;
    const float bias = a.bias ? __ldg(a.bias + o) : 0.f;
    for (int bb = 0; bb < a.B; bb += 8) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int k = lane; k < a.K; k += 32) {
        const float w = __ldg(wr + k);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (bb + j < a.B) acc[j] = fmaf(w, __ldg(a.x + (size_t)(bb + j) * a.K + k), acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = acc[j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(FULL_MASK, s, d);
        if (lane == 0 && bb + j < a.B) {
          s += bias;
          if (a.relu) s = fmaxf(s, 0.f);
          a.out[(size_t)(bb + j) * a.O + o] = s;
        }
      }
    }
  }
}

template <int NB, int S>
int launch_stream(const LinArgs& a, cudaStream_t st) {
  const size_t smem = stream_smem_bytes<NB, S>(a.K);
  auto kern = linear_stream_kernel<NB, S>;
  B200BEV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int gx = sm_count();
  const int max_gx = ceil_div(a.O, kStepRows);
  if (gx > max_gx) gx = max_gx;
  const int TB = 4 * NB;
  kern<<<dim3(gx, ceil_div(a.B, TB)), kThreads, smem, st>>>(a);
  return launch_status();
}

constexpr size_t kSmemCap = 220 * 1024;

}  // namespace

int dense_layer(const float* x, int B, int K, const float* w, const float* bias, int O, int relu, float* out,
                cudaStream_t st) {
  LinArgs a{x, w, bias, out, B, K, O, relu};
  const bool aligned = (((uintptr_t)x | (uintptr_t)w) & 15) == 0;
  // row-streaming form (16 warps per SM, batch tile in registers): batch <= 8, K = 128 .. 512.  (A 16-row tile runs too —
  // linear_rowstream_kernel<16, 4, 4> — but its shuffles and FMAs take 74 us on the 164 MB layer where the kernel below
  // takes 65.)
  if (aligned && O >= 256 && B <= 8 && !debug_env("B200BEV_DENSE_OLD")) {
    if (K == 512) return launch_rowstream<8, 8, 2>(a, st);
    if (K == 256) return launch_rowstream<8, 8, 1>(a, st);
    if (K == 128) return launch_rowstream<8, 4, 1>(a, st);
  }
  // streaming form: K splits into four quarters of whole 32-float chunks, there is more than one 64-row step of work,
  // and a batch tile fits next to the rings (the widest tile that fits, the deepest ring next to it)
  if (aligned && K % (kKSplit * kKC) == 0 && O >= 2 * kStepRows) {
    if (B > 16 && stream_smem_bytes<8, 3>(K) <= kSmemCap) return launch_stream<8, 3>(a, st);
    if (B > 8) {
      if (stream_smem_bytes<4, 4>(K) <= kSmemCap) return launch_stream<4, 4>(a, st);
      if (stream_smem_bytes<4, 3>(K) <= kSmemCap) return launch_stream<4, 3>(a, st);
    }
    if (stream_smem_bytes<2, 5>(K) <= kSmemCap) return launch_stream<2, 5>(a, st);
    if (stream_smem_bytes<2, 3>(K) <= kSmemCap) return launch_stream<2, 3>(a, st);
  }
  long long blocks = ((long long)O + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  linear_rows_kernel<<<(int)blocks, 256, 0, st>>>(a);
  return launch_status();
}

}  // namespace b200bev

using namespace b200bev;

extern "C" B200BEV_API int b200bev_dense_layer(const float* x, int B, int K, const float* weight, const float* bias, int O,
                                   int relu, float* out, void* stream) {
  if (!x || !weight || !out || B <= 0 || K <= 0 || O <= 0) return B200BEV_ERR_INVALID_ARGUMENT;
  return dense_layer(x, B, K, weight, bias, O, relu, out, (cudaStream_t)stream);
}

extern "C" B200BEV_API int b200bev_lidar_init(const float* lidar_features, int B, int K, const float* w1, const float* b1,
                                  int hidden, const float* w2, const float* b2, int O, float* hidden_ws, float* out,
                                  void* stream) {
  if (!lidar_features || !w1 || !w2 || !hidden_ws || !out || B <= 0 || K <= 0 || hidden <= 0 || O <= 0)
    return B200BEV_ERR_INVALID_ARGUMENT;
  const int rc = dense_layer(lidar_features, B, K, w1, b1, hidden, 1, hidden_ws, (cudaStream_t)stream);
  if (rc) return rc;
  return dense_layer(hidden_ws, B, hidden, w2, b2, O, 0, out, (cudaStream_t)stream);
}
