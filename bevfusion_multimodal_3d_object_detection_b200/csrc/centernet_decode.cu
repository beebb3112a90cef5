// S3 — CenterNet heat-map peak extraction (3x3 max-pool NMS), two-stage top-K and box decode.
//
// Behaviour follows decode_centernet_predictions / _nms / _topk of the reference
// (src/centernet_target.py:326-452, duplicated at src/fusion_detection.py:695-820); the structure
// does not: the reference runs max_pool2d, ==, *, two torch.topk sorts, three gathers and a python
// loop over the batch with one host sync per sample.  Here one launch does all of it.
//
//   grid (C, B), 256 threads.  CTA (c,b):
//     1. warp-shuffle 3x3 max: each warp owns 30-column x 8-row patches, loads the patch plus a
//        one-cell halo straight into registers (every heat-map value is read once from HBM),
//        takes the horizontal max with two shuffles and the vertical max in registers, and writes
//        the order-preserving integer key of heat*keep to shared memory;
//     2. radix-select (4 passes of 8 bits over the keys in shared memory) finds the K-th largest
//        key, an ordered compaction keeps exactly K entries (ties: lowest index first), a bitonic
//        sort orders them (score desc, index asc) -> the class's candidate list in the workspace;
//     3. the last CTA of a sample to finish (atomic ticket) merges the C*K candidates the same
//        way, gathers the nine regression channels at the K winners and writes the boxes.
//
// Tie order: torch.topk leaves it unspecified; this kernel defines it as ascending
// flat index, which is what the oracle uses too.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace b200bev {

namespace {

constexpr int kDecodeThreads = 256;
constexpr int kMaxK = 2048;

struct DecodeArgs {
  const float* heat;
  const float* off;
  const float* size;
  const float* rot;
  const float* vel;
  int B, C, H, W, K, P;
  float voxel, x0, y0, z, thresh;
  float* boxes;
  float* scores;
  int64_t* labels;
  float* vels;
  int64_t* ys;
  int64_t* xs;
  int64_t* ind;
  int32_t* count;
  unsigned long long* cand;  // workspace: (B, C, K) packed (key << 32 | ~index)
  int* done;                 // workspace: (B) arrival tickets, zero on entry and on exit
  unsigned long long* dbg;   // debug only (B200BEV_DECODE_TRACE): clock stamps
};

__device__ __forceinline__ unsigned long long pack_entry(uint32_t key, uint32_t index) {
  return ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - index);
}

// Step 1: keys[y*W+x] = key(heat * keep) with keep = (3x3 max == heat), -inf padding
// (F.max_pool2d(heat, 3, stride=1, padding=1), src/centernet_target.py:419-421).
//
// SIGMOID: the plane holds the heat-map head's raw output and torch.sigmoid (src/fusion.py:870-871) is applied as the
// values are loaded — 1/(1+exp(-x)) with the accurate expf and an IEEE divide, the expression ATen's CUDA sigmoid
// evaluates, so the scores are the bits torch.sigmoid gives on the same device.  The peak test runs on the sigmoid
// values, as in the reference: two logits that round to one float after the sigmoid form a plateau there too.
__device__ __forceinline__ float sigmoid_f32(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

template <bool DO_NMS, bool SIGMOID>
__device__ void load_plane_keys(const float* __restrict__ plane, int H, int W, uint32_t* keys) {
  const int tid = threadIdx.x;
  if (!DO_NMS) {
    for (int i = tid; i < H * W; i += blockDim.x) {
      const float v = __ldg(plane + i);
      keys[i] = float_to_key(SIGMOID ? sigmoid_f32(v) : v);
    }
    return;
  }
  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int n_strips = ceil_div(W, 30), n_rblocks = ceil_div(H, 8);
  for (int item = warp; item < n_strips * n_rblocks; item += nwarps) {
    const int strip = item % n_strips, rb = item / n_strips;
    const int x = strip * 30 - 1 + lane;  // lanes 0 and 31 carry the halo columns
    const int y0 = rb * 8;
    const bool x_ok = (x >= 0) && (x < W);
    float v[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const int y = y0 - 1 + r;
      v[r] = (x_ok && y >= 0 && y < H) ? __ldg(plane + y * W + x) : -INFINITY;
    }
    if (SIGMOID) {
      // after all ten loads are in flight; sigmoid(-inf) = 1/(1+inf) = 0 would be a legal score, so the padding is
      // put back explicitly
#pragma unroll
      for (int r = 0; r < 10; ++r) v[r] = (v[r] == -INFINITY) ? -INFINITY : sigmoid_f32(v[r]);
    }
    float hm[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const float l = __shfl_up_sync(FULL_MASK, v[r], 1);
      const float rr = __shfl_down_sync(FULL_MASK, v[r], 1);
      hm[r] = fmaxf(v[r], fmaxf(l, rr));
    }
    if (lane >= 1 && lane <= 30 && x < W) {
#pragma unroll
      for (int r = 1; r <= 8; ++r) {
        const int y = y0 - 1 + r;
        if (y < H) {
          const float m = fmaxf(hm[r - 1], fmaxf(hm[r], hm[r + 1]));
          const float keep = (m == v[r]) ? 1.0f : 0.0f;
          keys[y * W + x] = float_to_key(v[r] * keep);
        }
      }
    }
  }
}

// Steps 2: top-K of keys[0..n) -> sel[0..K) sorted by (key desc, index asc); sel has P >= K slots
// (P a power of two).  hist: 256 counters, wsum: 32 ints, scal: 4 ints.  All threads must call.
__device__ void block_topk(const uint32_t* keys, int n, int K, unsigned long long* sel, int P,
                           uint32_t* hist, int* wsum, int* scal, unsigned long long* dbg = nullptr, int* dn = nullptr) {
  auto stamp = [&]() { if (dbg && threadIdx.x == 0) dbg[(*dn)++] = clock64(); };
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = nthr >> 5;
  uint32_t prefix = 0, mask = 0;
  int need = K;
  const int n_round = ceil_div(n, 32) * 32;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += nthr) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n_round; i += nthr) {
      uint32_t digit = 0xffffu + lane;  // distinct sentinel per lane: never aggregated
      if (i < n) {
        const uint32_t k = keys[i];
        if ((k & mask) == prefix) digit = (k >> shift) & 255u;
      }
      if (pass == 0) {
        // warp-aggregate: post-NMS most keys are the key of 0.0 and would serialise on one counter.  (Only in the
        // first pass: later passes see a few matching keys among lanes that all differ, and MATCH.ANY costs ~600 clk
        // on 32 distinct values — it was 3/4 of this function's time.)
        const unsigned peers = __match_any_sync(FULL_MASK, digit);
        if (digit < 256u && (peers & lanemask_lt()) == 0) atomicAdd(&hist[digit], (uint32_t)__popc(peers));
      } else if (digit < 256u) {
        atomicAdd(&hist[digit], 1u);
      }
    }
    __syncthreads();
    if (warp == 0) {
      // lane l owns digits 255-8l .. 248-8l, scanned from the top
      uint32_t cnt[8];
      int s = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cnt[j] = hist[255 - 8 * lane - j];
        s += (int)cnt[j];
      }
      const int incl = warp_incl_scan(s, lane);
      const int excl = incl - s;
      if (excl < need && need <= incl) {
        int cum = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (cum < need && cum + (int)cnt[j] >= need) {
            scal[0] = 255 - 8 * lane - j;
            scal[1] = need - cum;
          }
          cum += (int)cnt[j];
        }
      }
    }
    __syncthreads();
    prefix |= ((uint32_t)scal[0]) << shift;
    mask |= 0xffu << shift;
    need = scal[1];
    __syncthreads();
    stamp();
  }
  const uint32_t T = prefix;  // the K-th largest key
  const int need_eq = need;   // how many entries equal to T belong to the top K
  const int n_gt = K - need_eq;
  if (tid == 0) scal[2] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += nthr) {
    const uint32_t k = keys[i];
    if (k > T) {
      const int s = atomicAdd(&scal[2], 1);
      sel[s] = pack_entry(k, (uint32_t)i);
    }
  }
  stamp();
  // entries equal to T, lowest index first
  if (need_eq == 1) {
    // the tie-free case: the one entry equal to T with the lowest index — a shared-memory minimum, one barrier,
    // instead of walking the keys 256 at a time with two barriers per step
    if (tid == 0) scal[5] = 0x7fffffff;
    __syncthreads();
    int best = 0x7fffffff;
    for (int i = tid; i < n; i += nthr)
      if (keys[i] == T && i < best) best = i;
    if (best != 0x7fffffff) atomicMin(&scal[5], best);
    __syncthreads();
    if (tid == 0) sel[n_gt] = pack_entry(T, (uint32_t)scal[5]);
  }
  int running = 0;
  for (int base = 0; need_eq > 1 && base < n && running < need_eq; base += nthr) {
    const int i = base + tid;
    const bool f = (i < n) && (keys[i] == T);
    const unsigned bal = __ballot_sync(FULL_MASK, f);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < nwarps; ++w) {
      const int v = wsum[w];
      if (w < warp) before += v;
      total += v;
    }
    const int pos = running + before + __popc(bal & lanemask_lt());
    if (f && pos < need_eq) sel[n_gt + pos] = pack_entry(T, (uint32_t)i);
    running += total;
    __syncthreads();
  }
  for (int r = K + tid; r < P; r += nthr) sel[r] = 0ull;
  __syncthreads();
  stamp();
  if (K <= nthr) {
    // The common sizes: every entry counts the entries above it — K broadcast reads per thread and ONE barrier — and
    // lands at its rank (the packed entries are all different, so the ranks are a permutation).  The bitonic
    // network below needs 28 block-wide barriers for 128 entries (8,400 clk measured, this: ~1,500).
    const unsigned long long mine = tid < K ? sel[tid] : 0ull;
    int rank = 0;
    if (tid < K)
      for (int j = 0; j < K; ++j) rank += sel[j] > mine ? 1 : 0;
    __syncthreads();
    if (tid < K) sel[rank] = mine;
    __syncthreads();
    stamp();
    return;
  }
  // bitonic sort, descending
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < P; t += nthr) {
        const int u = t ^ j;
        if (u > t) {
          const unsigned long long a = sel[t], b = sel[u];
          const bool desc = ((t & k) == 0);
          if (desc ? (a < b) : (a > b)) {
            sel[t] = b;
            sel[u] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  stamp();
}

template <bool DO_NMS, bool DO_DECODE, bool SIGMOID>
__global__ void __launch_bounds__(kDecodeThreads) centernet_topk_kernel(DecodeArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* sel = reinterpret_cast<unsigned long long*>(smem_raw);
  uint32_t* hist = reinterpret_cast<uint32_t*>(sel + a.P);
  int* wsum = reinterpret_cast<int*>(hist + 256);
  int* scal = wsum + 32;
  uint32_t* keys = reinterpret_cast<uint32_t*>(scal + 8);

  const int tid = threadIdx.x;
  const int c = blockIdx.x, b = blockIdx.y;
  const int HW = a.H * a.W, K = a.K;

  int dn = 0;
  unsigned long long* dbg = (a.dbg && b == 0 && c == a.C - 1) ? a.dbg : nullptr;
  if (dbg && tid == 0) dbg[dn++] = clock64();
  load_plane_keys<DO_NMS, SIGMOID>(a.heat + ((size_t)b * a.C + c) * HW, a.H, a.W, keys);
  __syncthreads();
  if (dbg && tid == 0) dbg[dn++] = clock64();
  block_topk(keys, HW, K, sel, a.P, hist, wsum, scal, dbg, &dn);

  unsigned long long* cand_b = a.cand + (size_t)b * a.C * K;
  for (int r = tid; r < K; r += blockDim.x) cand_b[(size_t)c * K + r] = sel[r];
  __threadfence();
  __syncthreads();
  if (tid == 0) scal[3] = (atomicAdd(&a.done[b], 1) == a.C - 1);
  __syncthreads();
  if (!scal[3]) return;
  __threadfence();

  // Step 3: cross-class merge. Candidate j = class*K + rank keeps the order of
  // topk_scores.view(B, -1) (src/centernet_target.py:440-441).
  const int n2 = a.C * K;
  for (int j = tid; j < n2; j += blockDim.x) keys[j] = (uint32_t)(__ldcg(cand_b + j) >> 32);
  if (tid == 0) scal[4] = 0;
  __syncthreads();
  if (a.dbg && b == 0 && tid == 0) { dbg = a.dbg + 16; dn = 0; dbg[dn++] = clock64(); } else dbg = nullptr;
  block_topk(keys, n2, K, sel, a.P, hist, wsum, scal, dbg, &dn);

  int n_pass = 0;
  for (int r = tid; r < K; r += blockDim.x) {
    const unsigned long long e = sel[r];
    const uint32_t j = 0xffffffffu - (uint32_t)e;
    const uint32_t idx = 0xffffffffu - (uint32_t)__ldcg(cand_b + j);
    const float score = key_to_float((uint32_t)(e >> 32));
    const int y = (int)idx / a.W, x = (int)idx % a.W;
    const size_t o = (size_t)b * K + r;
    a.scores[o] = score;
    // topk_classes = topk_indices // (H*W) with topk_indices < H*W: always 0
    a.labels[o] = 0;
    if (a.ind) a.ind[o] = (int64_t)j;
    if (a.ys) a.ys[o] = y;
    if (a.xs) a.xs[o] = x;
    if (DO_DECODE) {
      const size_t px = (size_t)y * a.W + x;
      const float* off = a.off + (size_t)b * 2 * HW;
      const float* sz = a.size + (size_t)b * 3 * HW;
      const float* rt = a.rot + (size_t)b * 2 * HW;
      const float* vl = a.vel + (size_t)b * 2 * HW;
      // separate mul and add, as torch evaluates centers*voxel + origin (src/centernet_target.py:383-393)
      const float cx = __fadd_rn((float)x, __ldg(off + px));
      const float cy = __fadd_rn((float)y, __ldg(off + HW + px));
      float* bx = a.boxes + o * 7;
      bx[0] = __fadd_rn(__fmul_rn(cx, a.voxel), a.x0);
      bx[1] = __fadd_rn(__fmul_rn(cy, a.voxel), a.y0);
      bx[2] = a.z;
      bx[3] = __ldg(sz + px);
      bx[4] = __ldg(sz + HW + px);
      bx[5] = __ldg(sz + 2 * HW + px);
      bx[6] = atan2f(__ldg(rt + px), __ldg(rt + HW + px));
      a.vels[o * 2 + 0] = __ldg(vl + px);
      a.vels[o * 2 + 1] = __ldg(vl + HW + px);
      if (score > a.thresh) ++n_pass;
    }
  }
  if (DO_DECODE) {
    if (n_pass) atomicAdd(&scal[4], n_pass);
    __syncthreads();
    if (tid == 0) a.count[b] = scal[4];
  }
  if (tid == 0) a.done[b] = 0;
  if (dbg && tid == 0) dbg[dn++] = clock64();
}

// Stand-alone _nms (src/centernet_target.py:416-421): same warp patches, result written as floats.
__global__ void __launch_bounds__(256) centernet_nms_kernel(const float* __restrict__ heat, float* __restrict__ out,
                                                            int planes, int H, int W) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int total_warps = (gridDim.x * blockDim.x) >> 5;
  const int n_strips = ceil_div(W, 30), n_rblocks = ceil_div(H, 8);
  const long long items = (long long)planes * n_strips * n_rblocks;
  for (long long item = warp; item < items; item += total_warps) {
    const int strip = (int)(item % n_strips);
    const int rb = (int)((item / n_strips) % n_rblocks);
    const long long p = item / ((long long)n_strips * n_rblocks);
    const float* plane = heat + p * H * W;
    float* oplane = out + p * H * W;
    const int x = strip * 30 - 1 + lane;
    const int y0 = rb * 8;
    const bool x_ok = (x >= 0) && (x < W);
    float v[10], hm[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const int y = y0 - 1 + r;
      v[r] = (x_ok && y >= 0 && y < H) ? __ldg(plane + y * W + x) : -INFINITY;
    }
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const float l = __shfl_up_sync(FULL_MASK, v[r], 1);
      const float rr = __shfl_down_sync(FULL_MASK, v[r], 1);
      hm[r] = fmaxf(v[r], fmaxf(l, rr));
    }
    if (lane >= 1 && lane <= 30 && x < W) {
#pragma unroll
      for (int r = 1; r <= 8; ++r) {
        const int y = y0 - 1 + r;
        if (y < H) {
          const float m = fmaxf(hm[r - 1], fmaxf(hm[r], hm[r + 1]));
          oplane[y * W + x] = v[r] * ((m == v[r]) ? 1.0f : 0.0f);
        }
      }
    }
  }
}

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

size_t decode_smem_bytes(int C, int HW, int K, int P) {
  const size_t nkeys = (size_t)(HW > C * K ? HW : C * K);
  return (size_t)P * 8 + 256 * 4 + 32 * 4 + 8 * 4 + nkeys * 4;
}

template <bool DO_NMS, bool DO_DECODE, bool SIGMOID = false>
int launch_topk(const DecodeArgs& a, cudaStream_t st) {
  const size_t smem = decode_smem_bytes(a.C, a.H * a.W, a.K, a.P);
  auto kern = centernet_topk_kernel<DO_NMS, DO_DECODE, SIGMOID>;
  if (smem > 48 * 1024) B200BEV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B200BEV_CUDA_TRY(cudaMemsetAsync(a.done, 0, sizeof(int) * (size_t)a.B, st));
  DecodeArgs aa = a;
  if (debug_env("B200BEV_DECODE_TRACE")) {
    B200BEV_CUDA_TRY(cudaMalloc(&aa.dbg, 32 * sizeof(unsigned long long)));
    B200BEV_CUDA_TRY(cudaMemsetAsync(aa.dbg, 0, 32 * sizeof(unsigned long long), st));
  }
  kern<<<dim3(a.C, a.B), kDecodeThreads, smem, st>>>(aa);
  if (aa.dbg) {
    unsigned long long h[32];
    B200BEV_CUDA_TRY(cudaStreamSynchronize(st));
    B200BEV_CUDA_TRY(cudaMemcpy(h, aa.dbg, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "decode trace stage1 (class C-1 of sample 0), clk since start:");
    for (int i = 1; i < 16 && h[i]; ++i) fprintf(stderr, " %llu", h[i] - h[0]);
    fprintf(stderr, "\n  stage 2 (merge CTA of sample 0), clk since its start:");
    for (int i = 17; i < 32 && h[i]; ++i) fprintf(stderr, " %llu", h[i] - h[16]);
    fprintf(stderr, "  [stage-2 start is %lld clk after stage-1 start of that CTA]\n", (long long)(h[16] - h[0]));
    cudaFree(aa.dbg);
  }
  return launch_status();
}

int check_topk_shape(int B, int C, int H, int W, int K, size_t ws_bytes, const void* ws) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K <= 0 || !ws) return B200BEV_ERR_INVALID_ARGUMENT;
  if ((long long)K > (long long)H * W) return B200BEV_ERR_K_OUT_OF_RANGE;
  if (K > kMaxK || C > 65535 || B > 65535) return B200BEV_ERR_UNSUPPORTED;
  if (decode_smem_bytes(C, H * W, K, next_pow2(K)) > 200 * 1024) return B200BEV_ERR_UNSUPPORTED;
  if (ws_bytes < b200bev_centernet_workspace_bytes(B, C, K) || ((uintptr_t)ws & 15)) return B200BEV_ERR_WORKSPACE;
  return B200BEV_OK;
}

}  // namespace
}  // namespace b200bev

using namespace b200bev;

extern "C" B200BEV_API size_t b200bev_centernet_workspace_bytes(int B, int C, int K) {
  if (B <= 0 || C <= 0 || K <= 0) return 0;
  const size_t cand = (size_t)B * C * K * sizeof(unsigned long long);
  return cand + (((size_t)B * sizeof(int) + 15) & ~(size_t)15);
}

extern "C" B200BEV_API int b200bev_centernet_nms(const float* heat, int B, int C, int H, int W, float* out, void* stream) {
  if (!heat || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return B200BEV_ERR_INVALID_ARGUMENT;
  const long long items = (long long)B * C * ceil_div(W, 30) * ceil_div(H, 8);
  long long blocks = (items + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  centernet_nms_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(heat, out, B * C, H, W);
  return launch_status();
}

extern "C" B200BEV_API int b200bev_centernet_topk(const float* scores, int B, int C, int H, int W, int K, float* topk_score,
                                      int64_t* topk_ind, int64_t* topk_classes, int64_t* topk_ys, int64_t* topk_xs,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!scores || !topk_score || !topk_classes) return B200BEV_ERR_INVALID_ARGUMENT;
  const int rc = check_topk_shape(B, C, H, W, K, workspace_bytes, workspace);
  if (rc) return rc;
  DecodeArgs a{};
  a.heat = scores;
  a.B = B; a.C = C; a.H = H; a.W = W; a.K = K; a.P = next_pow2(K);
  a.scores = topk_score;
  a.labels = topk_classes;
  a.ind = topk_ind;
  a.ys = topk_ys;
  a.xs = topk_xs;
  a.cand = reinterpret_cast<unsigned long long*>(workspace);
  a.done = reinterpret_cast<int*>(a.cand + (size_t)B * C * K);
  return launch_topk<false, false>(a, (cudaStream_t)stream);
}

namespace {
int decode_entry(bool logits, const float* heatmap, const float* offset, const float* size, const float* rot,
                 const float* vel, int B, int C, int H, int W, int K, float voxel_size, float x_origin, float y_origin,
                 float z_value, float score_thresh, float* boxes, float* scores, int64_t* labels, float* velocities,
                 int64_t* ys, int64_t* xs, int64_t* ind, int32_t* count, void* workspace, size_t workspace_bytes,
                 void* stream) {
  if (!heatmap || !offset || !size || !rot || !vel || !boxes || !scores || !labels || !velocities || !count)
    return B200BEV_ERR_INVALID_ARGUMENT;
  const int rc = check_topk_shape(B, C, H, W, K, workspace_bytes, workspace);
  if (rc) return rc;
  DecodeArgs a{};
  a.heat = heatmap; a.off = offset; a.size = size; a.rot = rot; a.vel = vel;
  a.B = B; a.C = C; a.H = H; a.W = W; a.K = K; a.P = next_pow2(K);
  a.voxel = voxel_size; a.x0 = x_origin; a.y0 = y_origin; a.z = z_value; a.thresh = score_thresh;
  a.boxes = boxes; a.scores = scores; a.labels = labels; a.vels = velocities;
  a.ys = ys; a.xs = xs; a.ind = ind; a.count = count;
  a.cand = reinterpret_cast<unsigned long long*>(workspace);
  a.done = reinterpret_cast<int*>(a.cand + (size_t)B * C * K);
  return logits ? launch_topk<true, true, true>(a, (cudaStream_t)stream) : launch_topk<true, true>(a, (cudaStream_t)stream);
}
}  // namespace

extern "C" B200BEV_API int b200bev_centernet_decode(const float* heatmap, const float* offset, const float* size, const float* rot,
                                        const float* vel, int B, int C, int H, int W, int K, float voxel_size,
                                        float x_origin, float y_origin, float z_value, float score_thresh,
                                        float* boxes, float* scores, int64_t* labels, float* velocities, int64_t* ys,
                                        int64_t* xs, int64_t* ind, int32_t* count, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  return decode_entry(false, heatmap, offset, size, rot, vel, B, C, H, W, K, voxel_size, x_origin, y_origin, z_value,
                      score_thresh, boxes, scores, labels, velocities, ys, xs, ind, count, workspace, workspace_bytes, stream);
}

extern "C" B200BEV_API int b200bev_centernet_decode_logits(const float* heatmap_logits, const float* offset, const float* size,
                                               const float* rot, const float* vel, int B, int C, int H, int W, int K,
                                               float voxel_size, float x_origin, float y_origin, float z_value,
                                               float score_thresh, float* boxes, float* scores, int64_t* labels,
                                               float* velocities, int64_t* ys, int64_t* xs, int64_t* ind, int32_t* count,
                                               void* workspace, size_t workspace_bytes, void* stream) {
  return decode_entry(true, heatmap_logits, offset, size, rot, vel, B, C, H, W, K, voxel_size, x_origin, y_origin, z_value,
                      score_thresh, boxes, scores, labels, velocities, ys, xs, ind, count, workspace, workspace_bytes, stream);
}
