// Shared pieces of the canonical-fp32 B x B tile kernels (batch_hard.cu, batch_all.cu): a warp computes
// 8 anchors x 4 columns per step, lane l accumulating chain l of all 32 entries, and a transposing butterfly
// hands entry e to lane e with exactly the reduction tree of dif_canon.cuh.
#pragma once
#include "dif_canon.cuh"
#include "dif_common.cuh"

namespace dif {

constexpr int BH_WARPS = 4;
constexpr int BH_TI = 8;                      // anchors per warp
constexpr int BH_TJ = 4;                      // columns per warp step
constexpr int BH_RB = BH_WARPS * BH_TI;       // anchors per block
constexpr int BH_CB = 32;                     // columns staged per tile
constexpr int BH_MAX_KD = 16;                 // D <= 512

struct BhRec {            // per (split, anchor)
  float pos_val; int pos_idx; int pos_cnt;   // cosine: min over positives; euclid: max over positives
  float neg_val; int neg_idx; int neg_cnt;   // cosine: max over negatives; euclid: min over negatives
  float all_max; int all_idx; int all_cnt;   // max over every column (euclid filler, losses.py:70)
  float row_sum; int n_pos; float pos_sum;   // sum of dist over the positive columns (batch-all)
};

struct BhRow {            // per anchor, written by bh_merge_kernel, read by bh_grad_kernel
  float pos_val, neg_val;   // extreme over REAL positives / negatives
  int pos_idx, neg_idx;     // first index, -1 if none
  int pos_cnt, neg_cnt;     // real tie counts
  float coef_pos, coef_neg; // dL/d(dist) applied to each tied real positive / negative column
  float all_max; int all_idx; int all_cnt;
  float coef_gmax;          // dL/d(dist) applied to every position holding the global max (euclid filler gradient)
};

template <bool MIN>
__device__ __forceinline__ void fold(float v, int j, float& val, int& idx, int& cnt) {
  if (MIN ? (v < val) : (v > val)) {
    val = v;
    idx = j;
    cnt = 1;
  } else if (v == val) {
    ++cnt;
    idx = (idx < 0 || j < idx) ? j : idx;
  }
}
template <bool MIN>
__device__ __forceinline__ void merge(float v, int j, int c, float& val, int& idx, int& cnt) {
  if (c == 0) return;
  if (cnt == 0 || (MIN ? (v < val) : (v > val))) {
    val = v;
    idx = j;
    cnt = c;
  } else if (v == val) {
    cnt += c;
    idx = j < idx ? j : idx;
  }
}

// lane l holds chain-l partial sums of 32 entries in v[0..31]; returns the canonical total of entry `lane`.
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int e = 0; e < o; ++e) {
      const float send = up ? v[e] : v[e + o];
      const float keep = up ? v[e + o] : v[e];
      v[e] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, o));
    }
  }
  return v[0];
}

// Load `n` rows [row0, row0+n) of x into smem (zero beyond B); cosine: normalise in place; writes the
// canonical sum of squares (euclid) or inverse norm (cosine) of each row to aux[n] in smem.
template <bool COSINE>
__device__ __forceinline__ void stage_rows(const float* __restrict__ x, int B, int D, int row0, int n,
                                           float* __restrict__ dst, float* __restrict__ aux) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kd = (D + 31) / 32;
  // Phase A: copy this warp's rows (warp, warp + 4, ...) with 16 independent loads in flight per step.  A small
  // batch is latency bound: one row at a time cost a full global-load round trip per row (8 per warp and tile).
  for (int rb = warp; rb < n; rb += 4 * BH_WARPS) {
    for (int c0 = 0; c0 < kd; c0 += 4) {
      float t[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = rb + u * BH_WARPS, gr = row0 + r;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int d = (c0 + cc) * 32 + lane;
          t[u][cc] = (r < n && gr < B && d < D) ? x[(size_t)gr * D + d] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = rb + u * BH_WARPS;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int d = (c0 + cc) * 32 + lane;
          if (r < n && d < D) dst[r * D + d] = t[u][cc];
        }
      }
    }
  }
  // Phase B: each lane re-reads only what it wrote itself (d = lane, lane + 32, ...): no barrier needed.
  for (int r = warp; r < n; r += BH_WARPS) {
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float t = dst[r * D + d];
      acc = __fmaf_rn(t, t, acc);
    }
    const float ss = canon_tree(acc);
    if (COSINE) {
      const float inv = canon_inv_norm(ss);
      for (int d = lane; d < D; d += 32) dst[r * D + d] = __fmul_rn(dst[r * D + d], inv);
      if (lane == 0) aux[r] = inv;
    } else {
      if (lane == 0) aux[r] = ss;
    }
  }
}


// Tensor-core miner (bh_tc.cu): writes ONE merged record per anchor (recs [B]) and aux [B] (inverse norm |
// sum of squares) - the same data bh_mine_kernel + the split merge produce, bit for bit.
template <bool COSINE>
int bh_mine_tensor(const float* emb, const int32_t* labels, int B, int D, BhRec* recs, float* aux,
                   unsigned long long* gmax_key, cudaStream_t st);

// One warp step: canonical dot products of anchors (warp*8 .. +7) with columns (j0 .. j0+3) of the staged
// tiles; returns the value of entry (lane >> 2, lane & 3).
__device__ __forceinline__ float tile_step_dot(const float* __restrict__ sa, const float* __restrict__ sb, int D, int kd,
                                               int warp, int lane, int j0) {
  float v[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) v[e] = 0.f;
  for (int c = 0; c < kd; ++c) {
    const int d = c * 32 + lane;
    float a[BH_TI], b[BH_TJ];
#pragma unroll
    for (int i = 0; i < BH_TI; ++i) a[i] = d < D ? sa[(warp * BH_TI + i) * D + d] : 0.f;
#pragma unroll
    for (int j = 0; j < BH_TJ; ++j) b[j] = d < D ? sb[(j0 + j) * D + d] : 0.f;
#pragma unroll
    for (int i = 0; i < BH_TI; ++i)
#pragma unroll
      for (int j = 0; j < BH_TJ; ++j) v[i * BH_TJ + j] = __fmaf_rn(a[i], b[j], v[i * BH_TJ + j]);
  }
  return transpose_reduce32(v, lane);
}

}  // namespace dif
