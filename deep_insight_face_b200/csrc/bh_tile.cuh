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

// compact per-anchor record for the "who mined me" scan: x = positive column (-1 none, -2 tied: see BhRow),
// y = negative column, z / w = the two coefficients as float bits
__device__ __forceinline__ int4 make_compact(const BhRow& r) {
  int4 c;
  c.x = r.coef_pos == 0.f ? -1 : (r.pos_cnt == 1 ? r.pos_idx : -2);
  c.y = r.coef_neg == 0.f ? -1 : (r.neg_cnt == 1 ? r.neg_idx : -2);
  c.z = __float_as_int(r.coef_pos);
  c.w = __float_as_int(r.coef_neg);
  return c;
}


// Per-anchor loss and cotangent scale from the hardest positive / negative (x = hn - hp for cosine similarities,
// hp - hn for distances).  Hard margin (the reference, losses.py:47-51,81-85): max(x + alpha, 0), gradient 1 when
// x + alpha >= 0 (tf.maximum).  Soft margin (arXiv 1703.07737 eq. 4, not in the reference): log(1 + exp(x)),
// gradient sigmoid(x).  `basic` is the reference's own expression so the hard loss keeps its fp32 rounding.
__device__ __forceinline__ void bh_loss_rule(float basic, float x, int soft, float dl, float& loss, float& g) {
  if (!soft) {
    loss = fmaxf(basic, 0.f);
    g = basic >= 0.f ? dl : 0.f;
  } else {
    loss = x > 0.f ? x + log1pf(expf(-x)) : log1pf(expf(x));
    g = dl / (1.f + expf(-x));
  }
}


// One anchor of the multi-block finalize (tensor-core path): hardest values with their fillers, loss, gradient
// coefficients, the row / compact records of the gradient gather, and this anchor's terms of the statistics:
// sums = {sum dists, hardest_pos, hardest_neg, filler share, positions holding the max}.  gmax = max(dists), only
// read by the squared-L2 loss (its filler, losses.py:70) - so the cosine loss can run this inside the re-rank kernel.
template <bool COSINE>
__device__ __forceinline__ void bh_finalize_row(const BhRec& q, int i, int B, float alpha, int soft,
                                                const float* __restrict__ dloss, float gmax, float* __restrict__ loss,
                                                int32_t* __restrict__ pos_idx_out, int32_t* __restrict__ neg_idx_out,
                                                BhRow* __restrict__ rows, int4* __restrict__ compact, double (&sums)[5]) {
  BhRow r;
  r.pos_val = q.pos_val; r.neg_val = q.neg_val; r.pos_idx = q.pos_idx; r.neg_idx = q.neg_idx;
  r.pos_cnt = q.pos_cnt; r.neg_cnt = q.neg_cnt; r.all_max = q.all_max; r.all_idx = q.all_idx; r.all_cnt = q.all_cnt;
  const int n_pos = q.n_pos, n_non = B - n_pos;
  const float fill_p = COSINE ? 1.f : 0.f, fill_n = COSINE ? -1.f : gmax;
  float hp = r.pos_val, hn = r.neg_val;
  int tie_p = r.pos_cnt, tie_n = r.neg_cnt, pidx = r.pos_idx, nidx = r.neg_idx;
  if (n_non > 0) {
    if (r.pos_cnt == 0 || (COSINE ? fill_p < hp : fill_p > hp)) { hp = fill_p; tie_p = n_non; pidx = -1; r.pos_cnt = 0; }
    else if (fill_p == hp) tie_p += n_non;
  }
  if (n_pos > 0) {
    if (r.neg_cnt == 0 || (COSINE ? fill_n > hn : fill_n < hn)) { hn = fill_n; tie_n = n_pos; nidx = -1; r.neg_cnt = 0; }
    else if (fill_n == hn) tie_n += n_pos;
  }
  const float basic = COSINE ? __fadd_rn(__fsub_rn(hn, hp), alpha) : __fsub_rn(__fadd_rn(hp, alpha), hn);
  float li, g;
  bh_loss_rule(basic, COSINE ? hn - hp : hp - hn, soft, dloss ? dloss[i] : 1.f / (float)B, li, g);
  loss[i] = li;
  if (pos_idx_out) pos_idx_out[i] = pidx;
  if (neg_idx_out) neg_idx_out[i] = nidx;
  r.coef_pos = r.pos_cnt > 0 ? (COSINE ? -g : g) / (float)tie_p : 0.f;
  r.coef_neg = r.neg_cnt > 0 ? (COSINE ? g : -g) / (float)tie_n : 0.f;
  r.pos_idx = pidx;
  r.neg_idx = nidx;
  r.coef_gmax = (!COSINE && r.all_max == gmax) ? 1.f : 0.f;   // flag: the gradient kernel substitutes the real share
  rows[i] = r;
  compact[i] = make_compact(r);
  sums[0] = (double)q.row_sum;
  sums[1] = (double)hp;
  sums[2] = (double)hn;
  sums[3] = (!COSINE && n_pos > 0 && hn == gmax) ? (double)(-g) * (double)n_pos / (double)tie_n : 0.0;
  sums[4] = r.all_max == gmax ? (double)r.all_cnt : 0.0;
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
// `fin` (cosine only, rows != NULL): finalize every anchor inside the re-rank launch - loss, mined indices, the row /
// compact records of the gradient gather - and leave one partial-sum record per 8 anchors in fin.partials.
struct BhFinalize {
  float alpha; int soft;
  const float* dloss;
  float* loss; int32_t* pos_idx; int32_t* neg_idx;
  BhRow* rows; int4* compact; double* partials;
};
constexpr int kBhFusedFinalizeRows = 8;   // anchors per partial-sum record of the fused finalize
template <bool COSINE>
int bh_mine_tensor(const float* emb, const int32_t* labels, int B, int D, BhRec* recs, float* aux,
                   unsigned long long* gmax_key, const BhFinalize& fin, cudaStream_t st);

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
