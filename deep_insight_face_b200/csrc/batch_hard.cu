// Batch-hard triplet losses, forward + backward, in canonical fp32 arithmetic.
//   deep_insight_face/common/losses.py:33-51   BatchHardTripletLoss            (cosine,   DIF_LOSS_BH_COSINE)
//   deep_insight_face/common/losses.py:54-85   BatchHardTripletLossEuclidean   (sq. L2,   DIF_LOSS_BH_EUCLIDEAN)
//   deep_insight_face/common/losses.py:88-128  ...AutoAlpha                    (same kernel, alpha = state)
//
// Three launches per call, the B x B matrix never exists in memory:
//   bh_mine_kernel   grid (row blocks x column splits).  A warp owns 8 anchors x 4 columns at a time: lane l
//                    accumulates chain l (d = l, l+32, ...) of all 32 entries, a transposing butterfly hands
//                    entry e to lane e with exactly the reduction tree of dif_canon.cuh, and the lane folds
//                    the entry into the running masked min/max (value, first index, tie count) of its anchor.
//                    Rows are normalised on the fly for the cosine variant (l2_normalize, losses.py:39).
//   bh_merge_kernel  one block: merges the per-split records, global max / means (losses.py:70,72-80),
//                    per-anchor loss, mined indices and gradient coefficients.
//   bh_grad_kernel   one warp per row r: gathers  sum_j (G_rj + G_jr) * d(dist_rj)/d(x_r)  from the mined
//                    columns of row r and from the rows that mined r (no atomics, fixed order), then the
//                    l2_normalize backward.  Exact ties follow TensorFlow's reduce_min/max gradient: the
//                    cotangent is split evenly over every tied position, filler positions included.
#include <algorithm>

#include "bh_tile.cuh"
#include "canon_mm.cuh"
#include "dif_ptx.cuh"

namespace dif {

template <bool COSINE>
__global__ void __launch_bounds__(BH_WARPS * 32) bh_mine_kernel(const float* __restrict__ x,
                                                                const int32_t* __restrict__ labels, int B, int D,
                                                                int cols_per_split, BhRec* __restrict__ recs,
                                                                float* __restrict__ row_aux) {
  extern __shared__ float sm[];
  float* sa = sm;                       // [BH_RB][D]
  float* sb = sa + BH_RB * D;           // [BH_CB][D]
  float* aux_a = sb + BH_CB * D;        // [BH_RB]
  float* aux_b = aux_a + BH_RB;         // [BH_CB]
  int* lab_b = reinterpret_cast<int*>(aux_b + BH_CB);  // [BH_CB]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * BH_RB;
  const int split = blockIdx.y;
  const int c_begin = split * cols_per_split, c_end = min(B, c_begin + cols_per_split);
  const int kd = (D + 31) / 32;

  stage_rows<COSINE>(x, B, D, row0, BH_RB, sa, aux_a);
  __syncthreads();
  if (split == 0 && threadIdx.x < BH_RB && row0 + threadIdx.x < B) row_aux[row0 + threadIdx.x] = aux_a[threadIdx.x];

  const int my_i = warp * BH_TI + (lane >> 2);   // anchor (block-local) this lane reports on
  const int gi = row0 + my_i;
  const int my_lab = gi < B ? labels[gi] : -1;
  const float my_aux = aux_a[my_i];

  float pos_val = COSINE ? INFINITY : -INFINITY, neg_val = COSINE ? -INFINITY : INFINITY, all_max = -INFINITY;
  int pos_idx = -1, neg_idx = -1, all_idx = -1, pos_cnt = 0, neg_cnt = 0, all_cnt = 0, n_pos = 0;
  float row_sum = 0.f, pos_sum = 0.f;

  for (int c0 = c_begin; c0 < c_end; c0 += BH_CB) {
    __syncthreads();   // previous tile fully consumed
    stage_rows<COSINE>(x, B, D, c0, BH_CB, sb, aux_b);
    if (threadIdx.x < BH_CB) lab_b[threadIdx.x] = (c0 + threadIdx.x < B) ? labels[c0 + threadIdx.x] : -2;
    __syncthreads();
    const int steps = min(BH_CB, c_end - c0);
    for (int j0 = 0; j0 < steps; j0 += BH_TJ) {
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
      const float dot = transpose_reduce32(v, lane);   // entry (lane >> 2, lane & 3)
      const int jl = j0 + (lane & 3);
      const int gj = c0 + jl;
      if (gj < c_end && gi < B) {
        const float dist = COSINE ? dot : __fsub_rn(__fadd_rn(my_aux, aux_b[jl]), __fmul_rn(2.f, dot));
        row_sum += dist;
        fold<false>(dist, gj, all_max, all_idx, all_cnt);
        if (lab_b[jl] == my_lab) {
          ++n_pos;
          pos_sum += dist;
          fold<COSINE>(dist, gj, pos_val, pos_idx, pos_cnt);
        } else {
          fold<!COSINE>(dist, gj, neg_val, neg_idx, neg_cnt);
        }
      }
    }
  }
  // combine the 4 lanes of an anchor
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    const float pv = __shfl_xor_sync(0xffffffffu, pos_val, o);
    const int pi = __shfl_xor_sync(0xffffffffu, pos_idx, o), pc = __shfl_xor_sync(0xffffffffu, pos_cnt, o);
    const float nv = __shfl_xor_sync(0xffffffffu, neg_val, o);
    const int ni = __shfl_xor_sync(0xffffffffu, neg_idx, o), nc = __shfl_xor_sync(0xffffffffu, neg_cnt, o);
    const float av = __shfl_xor_sync(0xffffffffu, all_max, o);
    const int ai = __shfl_xor_sync(0xffffffffu, all_idx, o), ac = __shfl_xor_sync(0xffffffffu, all_cnt, o);
    merge<COSINE>(pv, pi, pc, pos_val, pos_idx, pos_cnt);
    merge<!COSINE>(nv, ni, nc, neg_val, neg_idx, neg_cnt);
    merge<false>(av, ai, ac, all_max, all_idx, all_cnt);
    // fixed order: (lane0 + lane1) + (lane2 + lane3)
    row_sum = __fadd_rn(row_sum, __shfl_xor_sync(0xffffffffu, row_sum, o));
    pos_sum = __fadd_rn(pos_sum, __shfl_xor_sync(0xffffffffu, pos_sum, o));
    n_pos += __shfl_xor_sync(0xffffffffu, n_pos, o);
  }
  if ((lane & 3) == 0 && gi < B) {
    BhRec r;
    r.pos_val = pos_val; r.pos_idx = pos_idx; r.pos_cnt = pos_cnt;
    r.neg_val = neg_val; r.neg_idx = neg_idx; r.neg_cnt = neg_cnt;
    r.all_max = all_max; r.all_idx = all_idx; r.all_cnt = all_cnt;
    r.row_sum = row_sum; r.n_pos = n_pos; r.pos_sum = pos_sum;
    recs[(size_t)split * B + gi] = r;
  }
}

__device__ __forceinline__ double block_sum(double v, double* red) {
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

// three sums for the price of one barrier pair (same per-sum order as block_sum)
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* red) {
  for (int o = 16; o >= 1; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  __syncthreads();
  const int w = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
  if ((threadIdx.x & 31) == 0) {
    red[w] = a;
    red[32 + w] = b;
    red[64 + w] = c;
  }
  __syncthreads();
  a = b = c = 0.0;
  for (int i = 0; i < nw; ++i) {
    a += red[i];
    b += red[32 + i];
    c += red[64 + i];
  }
}

// One block.  stats = {mean(dists), mean(hardest_pos), mean(hardest_neg), max(dists)}
// Body shared by bh_merge_kernel (one block) and the fused small-batch kernel (every block redoes it into its own
// shared memory; `write_out` selects the block that also writes loss / indices / statistics).
template <bool COSINE>
__device__ __forceinline__ void bh_merge_body(const BhRec* __restrict__ recs, int n_splits, int B, float alpha, int soft,
                                              const float* __restrict__ dloss, float* __restrict__ loss,
                                              int32_t* __restrict__ pos_idx_out, int32_t* __restrict__ neg_idx_out,
                                              float* __restrict__ stats, BhRow* rows, int4* compact, bool write_out) {
  __shared__ double red[96];
  __shared__ unsigned long long gmax_key;   // orderable(value) << 32 | ~first row
  __shared__ int gmax_cnt_s;
  if (threadIdx.x == 0) {
    gmax_key = 0ull;
    gmax_cnt_s = 0;
  }
  __syncthreads();
  // phase 1: merge the splits of every anchor
  double sum_d = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    BhRow r;
    r.pos_val = COSINE ? INFINITY : -INFINITY; r.neg_val = COSINE ? -INFINITY : INFINITY; r.all_max = -INFINITY;
    r.pos_idx = r.neg_idx = r.all_idx = -1;
    r.pos_cnt = r.neg_cnt = r.all_cnt = 0;
    r.coef_pos = r.coef_neg = r.coef_gmax = 0.f;
    int n_pos = 0;
    float rs = 0.f;
    for (int s = 0; s < n_splits; ++s) {
      const BhRec& q = recs[(size_t)s * B + i];
      merge<COSINE>(q.pos_val, q.pos_idx, q.pos_cnt, r.pos_val, r.pos_idx, r.pos_cnt);
      merge<!COSINE>(q.neg_val, q.neg_idx, q.neg_cnt, r.neg_val, r.neg_idx, r.neg_cnt);
      merge<false>(q.all_max, q.all_idx, q.all_cnt, r.all_max, r.all_idx, r.all_cnt);
      rs = __fadd_rn(rs, q.row_sum);
      n_pos += q.n_pos;
    }
    r.coef_pos = __int_as_float(n_pos);   // parked here until phase 2
    rows[i] = r;
    sum_d += (double)rs;
    if (r.all_cnt > 0) atomicMax(&gmax_key, ((unsigned long long)float_orderable(r.all_max) << 32) | (0xFFFFFFFFu - (unsigned)i));
  }
  const double tot_d = block_sum(sum_d, red);
  __syncthreads();
  const uint32_t gmax_o = (uint32_t)(gmax_key >> 32);
  const float gmax = __uint_as_float((gmax_o & 0x80000000u) ? (gmax_o ^ 0x80000000u) : ~gmax_o);
  for (int i = threadIdx.x; i < B; i += blockDim.x)
    if (rows[i].all_max == gmax) atomicAdd(&gmax_cnt_s, rows[i].all_cnt);
  __syncthreads();
  const int gmax_cnt = gmax_cnt_s;

  // phase 2: hardest values with fillers, loss, gradient coefficients
  double sum_hp = 0.0, sum_hn = 0.0, gm_share = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    BhRow r = rows[i];
    const int n_pos = __float_as_int(r.coef_pos), n_non = B - n_pos;
    // filler values: cosine where(pos, S, 1) / where(pos, -1, S); euclid where(pos, D, 0) / where(pos, gmax, D)
    const float fill_p = COSINE ? 1.f : 0.f, fill_n = COSINE ? -1.f : gmax;
    float hp = r.pos_val, hn = r.neg_val;
    int tie_p = r.pos_cnt, tie_n = r.neg_cnt;
    int pidx = r.pos_idx, nidx = r.neg_idx;
    if (n_non > 0) {   // positive-side fillers sit in the non-positive columns
      if (r.pos_cnt == 0 || (COSINE ? fill_p < hp : fill_p > hp)) { hp = fill_p; tie_p = n_non; pidx = -1; r.pos_cnt = 0; }
      else if (fill_p == hp) tie_p += n_non;
    }
    if (n_pos > 0) {   // negative-side fillers sit in the positive columns
      if (r.neg_cnt == 0 || (COSINE ? fill_n > hn : fill_n < hn)) { hn = fill_n; tie_n = n_pos; nidx = -1; r.neg_cnt = 0; }
      else if (fill_n == hn) tie_n += n_pos;
    }
    const float basic = COSINE ? __fadd_rn(__fsub_rn(hn, hp), alpha) : __fsub_rn(__fadd_rn(hp, alpha), hn);
    float li, g;
    bh_loss_rule(basic, COSINE ? hn - hp : hp - hn, soft, dloss ? dloss[i] : 1.f / (float)B, li, g);
    if (write_out) {
      loss[i] = li;
      if (pos_idx_out) pos_idx_out[i] = pidx;
      if (neg_idx_out) neg_idx_out[i] = nidx;
    }
    sum_hp += (double)hp;
    sum_hn += (double)hn;
    // d loss / d dist at each tied position: cosine  +g/tie_n (neg), -g/tie_p (pos); euclid  +g/tie_p (pos), -g/tie_n (neg)
    r.coef_pos = r.pos_cnt > 0 ? (COSINE ? -g : g) / (float)tie_p : 0.f;
    r.coef_neg = r.neg_cnt > 0 ? (COSINE ? g : -g) / (float)tie_n : 0.f;
    r.pos_idx = pidx;
    r.neg_idx = nidx;
    if (!COSINE && n_pos > 0 && hn == gmax)   // share of the cotangent that lands on the max(dists) filler
      gm_share += (double)(-g) * (double)n_pos / (double)tie_n;
    rows[i] = r;
  }
  double tot_hp = sum_hp, tot_hn = sum_hn, tot_gm = gm_share;
  block_sum3(tot_hp, tot_hn, tot_gm, red);
  if (threadIdx.x == 0 && write_out) {
    stats[0] = B > 0 ? (float)(tot_d / ((double)B * (double)B)) : 0.f;
    stats[1] = B > 0 ? (float)(tot_hp / (double)B) : 0.f;
    stats[2] = B > 0 ? (float)(tot_hn / (double)B) : 0.f;
    stats[3] = gmax;
  }
  // every position holding the global max receives an equal part of gm_share
  const float cg = (!COSINE && gmax_cnt > 0) ? (float)(tot_gm / (double)gmax_cnt) : 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    rows[i].coef_gmax = (rows[i].all_max == gmax) ? cg : 0.f;
    compact[i] = make_compact(rows[i]);
  }
}

template <bool COSINE>
__global__ void __launch_bounds__(1024) bh_merge_kernel(const BhRec* __restrict__ recs, int n_splits, int B,
                                                        float alpha, int soft, const float* __restrict__ dloss,
                                                        float* __restrict__ loss, int32_t* __restrict__ pos_idx_out,
                                                        int32_t* __restrict__ neg_idx_out, float* __restrict__ stats,
                                                        BhRow* __restrict__ rows, int4* __restrict__ compact) {
  bh_merge_body<COSINE>(recs, n_splits, B, alpha, soft, dloss, loss, pos_idx_out, neg_idx_out, stats, rows, compact, true);
}

// Multi-block form of the second half of bh_merge_kernel for the tensor-core path (one merged record per anchor,
// max(dists) already reduced into gmax_key by the re-rank kernel).  Per-block partial sums go to `partials`
// [gridDim.x][5] = {sum dists, sum hardest_pos, sum hardest_neg, filler share, positions holding the max};
// bh_stats_kernel folds them in a fixed order.
template <bool COSINE>
__global__ void __launch_bounds__(256) bh_finalize_kernel(const BhRec* __restrict__ recs, int B, float alpha, int soft,
                                                          const float* __restrict__ dloss,
                                                          const unsigned long long* __restrict__ gmax_key,
                                                          float* __restrict__ loss, int32_t* __restrict__ pos_idx_out,
                                                          int32_t* __restrict__ neg_idx_out, BhRow* __restrict__ rows,
                                                          int4* __restrict__ compact, double* __restrict__ partials) {
  __shared__ double red[32];
  const uint32_t gmax_o = (uint32_t)(*gmax_key >> 32);
  const float gmax = __uint_as_float((gmax_o & 0x80000000u) ? (gmax_o ^ 0x80000000u) : ~gmax_o);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double sums[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (i < B) bh_finalize_row<COSINE>(recs[i], i, B, alpha, soft, dloss, gmax, loss, pos_idx_out, neg_idx_out, rows, compact, sums);
  const double s_d = sums[0], s_hp = sums[1], s_hn = sums[2], s_gm = sums[3], s_cnt = sums[4];
  const double t_d = block_sum(s_d, red), t_hp = block_sum(s_hp, red), t_hn = block_sum(s_hn, red);
  const double t_gm = block_sum(s_gm, red), t_cnt = block_sum(s_cnt, red);
  if (threadIdx.x == 0) {
    double* o = partials + (size_t)blockIdx.x * 5;
    o[0] = t_d; o[1] = t_hp; o[2] = t_hn; o[3] = t_gm; o[4] = t_cnt;
  }
}

// one warp: partials -> stats [4] and the per-position share of the max(dists) filler gradient
__device__ __forceinline__ void bh_stats_body(const double* __restrict__ partials, int n_parts, int B,
                                              const unsigned long long* __restrict__ gmax_key, float* __restrict__ stats,
                                              float* __restrict__ cg_out) {
  const int lane = threadIdx.x & 31;
  double t[5] = {0, 0, 0, 0, 0};
  for (int p = lane; p < n_parts; p += 32)
    for (int k = 0; k < 5; ++k) t[k] += partials[(size_t)p * 5 + k];
  for (int o = 16; o >= 1; o >>= 1)
    for (int k = 0; k < 5; ++k) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
  if (lane == 0) {
    const uint32_t gmax_o = (uint32_t)(*gmax_key >> 32);
    if (stats) {
      stats[0] = (float)(t[0] / ((double)B * (double)B));
      stats[1] = (float)(t[1] / (double)B);
      stats[2] = (float)(t[2] / (double)B);
      stats[3] = __uint_as_float((gmax_o & 0x80000000u) ? (gmax_o ^ 0x80000000u) : ~gmax_o);
    }
    *cg_out = t[4] > 0.0 ? (float)(t[3] / t[4]) : 0.f;
  }
}

__global__ void bh_stats_kernel(const double* __restrict__ partials, int n_parts, int B,
                                const unsigned long long* __restrict__ gmax_key, float* __restrict__ stats,
                                float* __restrict__ cg_out) {
  bh_stats_body(partials, n_parts, B, gmax_key, stats, cg_out);
}

// canonical dist(r, j) recomputed by one warp (tie resolution)
template <bool COSINE>
__device__ __forceinline__ float warp_dist(const float* __restrict__ x, const float* __restrict__ aux, int D, int r,
                                           int j) {
  const float* a = x + (size_t)r * D;
  const float* b = x + (size_t)j * D;
  float acc = 0.f;
  if (COSINE) {
    const float ia = aux[r], ib = aux[j];
    for (int d = (int)(threadIdx.x & 31u); d < D; d += 32) acc = __fmaf_rn(__fmul_rn(a[d], ia), __fmul_rn(b[d], ib), acc);
    return canon_tree(acc);
  }
  for (int d = (int)(threadIdx.x & 31u); d < D; d += 32) acc = __fmaf_rn(a[d], b[d], acc);
  return __fsub_rn(__fadd_rn(aux[r], aux[j]), __fmul_rn(2.f, canon_tree(acc)));
}

// acc[c] += w * (value of row j at d = lane + 32c); cosine: the normalised row; euclid: (x_r - x_j) * 2
template <bool COSINE>
__device__ __forceinline__ void axpy_row(float (&acc)[BH_MAX_KD], float w, const float* __restrict__ x,
                                         const float* __restrict__ aux, int D, int r, int j) {
  const int lane = threadIdx.x & 31;
  const float* b = x + (size_t)j * D;
  const float* a = x + (size_t)r * D;
  const float ib = COSINE ? aux[j] : 0.f;
#pragma unroll
  for (int c = 0; c < BH_MAX_KD; ++c) {
    const int d = c * 32 + lane;
    if (d < D) acc[c] += COSINE ? w * (b[d] * ib) : w * 2.f * (a[d] - b[d]);
  }
}

constexpr int BH_GRAD_WARPS = 8;
constexpr int BH_GRAD_CHUNK = 1024;   // compact records staged per pass (16 KB)

// bh_grad_map_kernel's scan also keeps, per row of the block, the (anchor, weight) pairs of the anchors that mined it
constexpr int BH_MAP_LIST = 32;
struct BhMapLists {
  int cnt[BH_GRAD_WARPS];
  int anchor[BH_GRAD_WARPS][BH_MAP_LIST];
  float weight[BH_GRAD_WARPS][BH_MAP_LIST];
  int any_tied;
  // squared-L2: the share of the max(dists) filler gradient is folded from the finalize kernel's partial sums by the
  // (one or two) warps whose row holds the maximum, not by every block
  const double* partials;
  int n_parts;
  const unsigned long long* gmax_key;
  float cg[BH_GRAD_WARPS];
};

// MODE 0: every block stages the compact records of all anchors and its warps scan them (O(B) per row);
// MODE 2: the rows that mined r come from the bitmaps bh_grad_map_kernel built in shared memory (O(in-degree) per row).
// Both visit the contributing anchors in ascending order, so the two variants give bit-identical gradients.
template <bool COSINE, int MODE>
__device__ __forceinline__ void bh_grad_body(const float* __restrict__ x, const int32_t* __restrict__ labels, int B, int D,
                                             const float* __restrict__ aux,   // inv norm | sum sq
                                             const BhRow* rows, const int4* compact, const float* __restrict__ cg_dev,
                                             float* __restrict__ demb, const unsigned* s_bits = nullptr,
                                             const BhMapLists* lists = nullptr) {
  __shared__ int4 s_c[MODE != 0 ? 1 : BH_GRAD_CHUNK];
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * BH_GRAD_WARPS + (threadIdx.x >> 5);
  const bool active = r < B && (threadIdx.x >> 5) < BH_GRAD_WARPS;   // (the cluster step's block has spare warps)
  float acc[BH_MAX_KD];
#pragma unroll
  for (int c = 0; c < BH_MAX_KD; ++c) acc[c] = 0.f;
  BhRow me;
  me.pos_val = me.neg_val = me.coef_pos = me.coef_neg = me.all_max = me.coef_gmax = 0.f;
  me.pos_idx = me.neg_idx = me.all_idx = -1;
  me.pos_cnt = me.neg_cnt = me.all_cnt = 0;
  if (active) me = rows[r];
  const int my_lab = active ? labels[r] : -1;
  if (cg_dev) me.coef_gmax = me.coef_gmax != 0.f ? *cg_dev : 0.f;   // multi-block finalize leaves a flag here
  if (MODE == 2 && !COSINE && active && me.coef_gmax != 0.f) {      // warp-uniform: this row holds max(dists)
    BhMapLists* wl = const_cast<BhMapLists*>(lists);
    bh_stats_body(wl->partials, wl->n_parts, B, wl->gmax_key, nullptr, &wl->cg[threadIdx.x >> 5]);
    __syncwarp();
    me.coef_gmax = wl->cg[threadIdx.x >> 5];
  }

  // MODE 2, the usual case (no ties anywhere, D <= 128, a short list): the warp's whole work is a list of (row, weight)
  // pairs - own positive, own negative, then the anchors that mined r in ascending order, the order of the code below -
  // so the rows are fetched four at a time instead of one dependent load after the other
  bool gathered = false;
  if (MODE == 2) {
    const int wid = threadIdx.x >> 5;
    const int n_inv = lists->cnt[wid];
    const bool untied = active && D <= 128 && !lists->any_tied && (me.coef_pos == 0.f || me.pos_cnt == 1) &&
                        (me.coef_neg == 0.f || me.neg_cnt == 1);
    if (untied) {   // warp-uniform
      const float* xr = x + (size_t)r * D;
      float a_r[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) a_r[c] = (!COSINE && c * 32 + lane < D) ? xr[c * 32 + lane] : 0.f;
      // n list entries, one per lane: rows fetched four at a time, added in list order (the expressions of axpy_row)
      auto gather = [&](int n, int my_j, float my_w) {
        for (int e0 = 0; e0 < n; e0 += 4) {
          float v[4][4], w[4], ib[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int src = min(e0 + u, n - 1);
            const int j = __shfl_sync(0xffffffffu, my_j, src);
            w[u] = __shfl_sync(0xffffffffu, my_w, src);
            ib[u] = COSINE ? aux[j] : 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) v[u][c] = c * 32 + lane < D ? x[(size_t)j * D + c * 32 + lane] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (e0 + u < n) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (c * 32 + lane < D) acc[c] += COSINE ? w[u] * (v[u][c] * ib[u]) : w[u] * 2.f * (a_r[c] - v[u][c]);
            }
        }
      };
      const int has_p = me.coef_pos != 0.f ? 1 : 0, has_n = me.coef_neg != 0.f ? 1 : 0;
      int my_j = 0;
      float my_w = 0.f;
      if (lane < has_p) {
        my_j = me.pos_idx;
        my_w = me.coef_pos;
      } else if (lane < has_p + has_n) {
        my_j = me.neg_idx;
        my_w = me.coef_neg;
      }
      if (n_inv <= BH_MAP_LIST - 2) {
        // the scan's entries, ranked by anchor: lane t >= has_p + has_n takes the entry whose rank is t - has_p - has_n
        const int want = lane - has_p - has_n;
        const int a_in = lane < n_inv ? lists->anchor[wid][lane] : 0x7fffffff;
        const float w_in = lane < n_inv ? lists->weight[wid][lane] : 0.f;
        int rank_in = 0;
        if (lane < n_inv)
          for (int e = 0; e < n_inv; ++e) rank_in += lists->anchor[wid][e] < a_in ? 1 : 0;
        for (int e = 0; e < n_inv; ++e) {
          const int re = __shfl_sync(0xffffffffu, rank_in, e), ae = __shfl_sync(0xffffffffu, a_in, e);
          const float we = __shfl_sync(0xffffffffu, w_in, e);
          if (re == want) {
            my_j = ae;
            my_w = we;
          }
        }
        gather(has_p + has_n + n_inv, my_j, my_w);
      } else {
        // a hub row (the hardest negative of dozens or hundreds of anchors: squared-L2 batches have them): own entries,
        // then the bitmap 32 words at a time, its set bits handed to the lanes in ascending order, 32 anchors per
        // round; every lane fetches its own anchor's record, then the rows are gathered as above
        gather(has_p + has_n, my_j, my_w);
        const int W = (B + 31) >> 5;
        const unsigned* bm = s_bits + wid * W;
        int* slot = const_cast<int*>(lists->anchor[wid]);   // (the scan's truncated list is not needed any more)
        for (int w0 = 0; w0 < W; w0 += 32) {
          const int wi = w0 + lane;
          const unsigned word = wi < W ? bm[wi] : 0u;
          const int cnt = __popc(word);
          int incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
          }
          const int total = __shfl_sync(0xffffffffu, incl, 31), first = incl - cnt;
          for (int base = 0; base < total; base += 32) {
            __syncwarp();
            unsigned u = word;
            for (int k = first; u; ++k) {
              const int bit = __ffs((int)u) - 1;
              u &= u - 1;
              if (k >= base && k < base + 32) slot[k - base] = wi * 32 + bit;
            }
            __syncwarp();
            const int n = min(32, total - base);
            int a = 0;
            float w = 0.f;
            if (lane < n) {
              a = slot[lane];
              const int4 c = compact[a];
              if (c.x == r) w += __int_as_float(c.z);
              if (c.y == r) w += __int_as_float(c.w);
            }
            gather(n, a, w);
          }
        }
      }
      gathered = true;
    }
  }
  // --- own row: tied real positive / negative columns
  if (!gathered && active && me.coef_pos != 0.f) {
    if (me.pos_cnt == 1) axpy_row<COSINE>(acc, me.coef_pos, x, aux, D, r, me.pos_idx);
    else
      for (int j = 0; j < B; ++j)
        if (labels[j] == my_lab && warp_dist<COSINE>(x, aux, D, r, j) == me.pos_val)
          axpy_row<COSINE>(acc, me.coef_pos, x, aux, D, r, j);
  }
  if (!gathered && active && me.coef_neg != 0.f) {
    if (me.neg_cnt == 1) axpy_row<COSINE>(acc, me.coef_neg, x, aux, D, r, me.neg_idx);
    else
      for (int j = 0; j < B; ++j)
        if (labels[j] != my_lab && warp_dist<COSINE>(x, aux, D, r, j) == me.neg_val)
          axpy_row<COSINE>(acc, me.coef_neg, x, aux, D, r, j);
  }
  if (MODE == 2) {
    // --- rows that mined r, from this warp's bitmap over the anchors (bit a set = anchor a mined r, or a is tied):
    // set bits come out in ascending anchor order, the order of the staged scan below
    if (active && !gathered) {
      const int W = (B + 31) >> 5;
      const unsigned* bm = s_bits + (threadIdx.x >> 5) * W;
      const unsigned* tm = s_bits + BH_GRAD_WARPS * W;
      for (int w0 = 0; w0 < W; w0 += 32) {
        const int wi = w0 + lane;
        const unsigned mb = wi < W ? bm[wi] : 0u, tb = wi < W ? tm[wi] : 0u;
        unsigned any = __ballot_sync(0xffffffffu, (mb | tb) != 0u);
        while (any) {
          const int src = __ffs((int)any) - 1;
          any &= any - 1;
          const unsigned tw = __shfl_sync(0xffffffffu, tb, src);
          unsigned u = __shfl_sync(0xffffffffu, mb, src) | tw;
          while (u) {
            const int bit = __ffs((int)u) - 1;
            u &= u - 1;
            const int jj = (w0 + src) * 32 + bit;
            const int4 c = compact[jj];
            float w = 0.f;
            if (c.x == r) w += __int_as_float(c.z);
            if (c.y == r) w += __int_as_float(c.w);
            if ((tw >> bit) & 1u) {
              const BhRow o = rows[jj];
              const float dv = warp_dist<COSINE>(x, aux, D, jj, r);
              const bool same = labels[jj] == my_lab;
              if (c.x == -2 && same && dv == o.pos_val) w += o.coef_pos;
              if (c.y == -2 && !same && dv == o.neg_val) w += o.coef_neg;
            }
            if (w != 0.f) axpy_row<COSINE>(acc, w, x, aux, D, r, jj);
          }
        }
      }
    }
  } else
  // --- rows that mined r: the block stages the compact records once and its 8 warps scan them
  for (int chunk0 = 0; chunk0 < B; chunk0 += BH_GRAD_CHUNK) {
    const int n_chunk = min(BH_GRAD_CHUNK, B - chunk0);
    __syncthreads();
    for (int t = threadIdx.x; t < n_chunk; t += blockDim.x) s_c[t] = compact[chunk0 + t];
    __syncthreads();
    if (!active) continue;
    for (int j0 = 0; j0 < n_chunk; j0 += 32) {
      const int jl = j0 + lane;
      float w0 = 0.f;
      bool tied_p = false, tied_n = false;
      if (jl < n_chunk) {
        const int4 c = s_c[jl];
        if (c.x == r) w0 += __int_as_float(c.z);
        if (c.y == r) w0 += __int_as_float(c.w);
        tied_p = c.x == -2;
        tied_n = c.y == -2;
      }
      unsigned hit = __ballot_sync(0xffffffffu, w0 != 0.f || tied_p || tied_n);
      while (hit) {
        const int src = __ffs((int)hit) - 1;
        hit &= hit - 1;
        const int jj = chunk0 + j0 + src;
        float w = __shfl_sync(0xffffffffu, w0, src);
        const bool tp = __shfl_sync(0xffffffffu, (int)tied_p, src) != 0, tn = __shfl_sync(0xffffffffu, (int)tied_n, src) != 0;
        if (tp || tn) {
          const BhRow o = rows[jj];
          const float dv = warp_dist<COSINE>(x, aux, D, jj, r);
          const bool same = labels[jj] == my_lab;
          if (tp && same && dv == o.pos_val) w += o.coef_pos;
          if (tn && !same && dv == o.neg_val) w += o.coef_neg;
        }
        // d dist(jj, r) / d x_r : cosine n_jj ; euclid 2 (x_r - x_jj)  -> same helper with roles (r, jj)
        if (w != 0.f) axpy_row<COSINE>(acc, w, x, aux, D, r, jj);
      }
    }
  }
  if (!active) return;
  // --- euclid: gradient through the max(dists) filler (both (r, j) and (j, r) hold the max)
  if (!COSINE && me.coef_gmax != 0.f) {
    if (me.all_cnt == 1) axpy_row<COSINE>(acc, 2.f * me.coef_gmax, x, aux, D, r, me.all_idx);
    else
      for (int j = 0; j < B; ++j)
        if (warp_dist<COSINE>(x, aux, D, r, j) == me.all_max) axpy_row<COSINE>(acc, 2.f * me.coef_gmax, x, aux, D, r, j);
  }
  // --- write: cosine applies the l2_normalize backward  dx = inv * (dn - n * (n . dn))
  if (COSINE) {
    const float inv = aux[r];
    const float* xr = x + (size_t)r * D;
    float dotp = 0.f, ss = 0.f;
#pragma unroll
    for (int c = 0; c < BH_MAX_KD; ++c) {
      const int d = c * 32 + lane;
      if (d < D) {
        dotp += acc[c] * (xr[d] * inv);
        ss += xr[d] * xr[d];
      }
    }
    for (int o = 16; o >= 1; o >>= 1) {
      dotp += __shfl_xor_sync(0xffffffffu, dotp, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    // below the epsilon, l2_normalize is x * rsqrt(eps): a plain scaling
    const bool clamped = ss < kNormEps;
#pragma unroll
    for (int c = 0; c < BH_MAX_KD; ++c) {
      const int d = c * 32 + lane;
      if (d < D) demb[(size_t)r * D + d] = clamped ? inv * acc[c] : inv * (acc[c] - (xr[d] * inv) * dotp);
    }
  } else {
#pragma unroll
    for (int c = 0; c < BH_MAX_KD; ++c) {
      const int d = c * 32 + lane;
      if (d < D) demb[(size_t)r * D + d] = acc[c];
    }
  }
}

template <bool COSINE>
__global__ void __launch_bounds__(BH_GRAD_WARPS * 32) bh_grad_kernel(const float* __restrict__ x,
                                                                     const int32_t* __restrict__ labels, int B, int D,
                                                                     const float* __restrict__ aux,
                                                                     const BhRow* __restrict__ rows,
                                                                     const int4* __restrict__ compact,
                                                                     const float* __restrict__ cg_dev,
                                                                     float* __restrict__ demb) {
  bh_grad_body<COSINE, 0>(x, labels, B, D, aux, rows, compact, cg_dev, demb);
}

// Tensor-core path: statistics + "who mined me" + gradient gather in ONE launch (round 2 had a one-block kernel build
// inverse CSR lists first: 15 us of a 68 us step at B = 4096).  Every block folds the finalize kernel's partial sums
// itself (a fixed order, so every block gets the same filler share; block 0 also writes the statistics), scans the
// compact records of all anchors once (16 B each, L2 resident) and sets bit a of row r's bitmap in shared memory when
// anchor a mined one of the block's 8 rows, plus a bitmap of the tied anchors; each warp then walks the set bits of
// its row in ascending anchor order - the summation order of the staged scan, so the gradients stay bit-identical.
constexpr int BH_MAP_MAX_B = 32768;   // (8 + 1) bitmaps of B bits: 36 KB of shared memory
template <bool COSINE>
__global__ void __launch_bounds__(BH_GRAD_WARPS * 32) bh_grad_map_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ labels, int B, int D, const float* __restrict__ aux,
    const BhRow* __restrict__ rows, const int4* __restrict__ compact, const double* __restrict__ partials, int n_parts,
    const unsigned long long* __restrict__ gmax_key, float* __restrict__ stats, float* __restrict__ demb) {
  extern __shared__ unsigned s_map[];   // [BH_GRAD_WARPS + 1][W]
  __shared__ float s_cg;
  __shared__ BhMapLists s_lists;
  const int W = (B + 31) >> 5;
  const int r0 = blockIdx.x * BH_GRAD_WARPS;
  for (int i = threadIdx.x; i < (BH_GRAD_WARPS + 1) * W; i += blockDim.x) s_map[i] = 0u;
  if (threadIdx.x < BH_GRAD_WARPS) s_lists.cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_lists.any_tied = 0;
  if (threadIdx.x == 0) {
    s_lists.partials = partials;
    s_lists.n_parts = n_parts;
    s_lists.gmax_key = gmax_key;
  }
  if ((threadIdx.x >> 5) == 0 && blockIdx.x == 0) bh_stats_body(partials, n_parts, B, gmax_key, stats, &s_cg);
  __syncthreads();
#pragma unroll 4
  for (int a = threadIdx.x; a < B; a += BH_GRAD_WARPS * 32) {
    const int4 c = compact[a];
    const unsigned ux = (unsigned)(c.x - r0), uy = (unsigned)(c.y - r0), bit = 1u << (a & 31);
    if (ux < (unsigned)BH_GRAD_WARPS) {
      atomicOr(&s_map[ux * W + (a >> 5)], bit);
      // (a hub row - the hardest negative of hundreds of anchors - must not serialise the block on its counter: once the
      // list is past its cap the row walks its bitmap anyway)
      if (*reinterpret_cast<volatile int*>(&s_lists.cnt[ux]) <= BH_MAP_LIST) {
        const int slot = atomicAdd(&s_lists.cnt[ux], 1);
        if (slot < BH_MAP_LIST) {
          s_lists.anchor[ux][slot] = a;
          s_lists.weight[ux][slot] = __int_as_float(c.z);
        }
      }
    }
    if (uy < (unsigned)BH_GRAD_WARPS) {
      atomicOr(&s_map[uy * W + (a >> 5)], bit);
      // (a hub row - the hardest negative of hundreds of anchors - must not serialise the block on its counter: once the
      // list is past its cap the row walks its bitmap anyway)
      if (*reinterpret_cast<volatile int*>(&s_lists.cnt[uy]) <= BH_MAP_LIST) {
        const int slot = atomicAdd(&s_lists.cnt[uy], 1);
        if (slot < BH_MAP_LIST) {
          s_lists.anchor[uy][slot] = a;
          s_lists.weight[uy][slot] = __int_as_float(c.w);
        }
      }
    }
    if (c.x == -2 || c.y == -2) {
      atomicOr(&s_map[BH_GRAD_WARPS * W + (a >> 5)], bit);
      s_lists.any_tied = 1;
    }
  }
  __syncthreads();
  bh_grad_body<COSINE, 2>(x, labels, B, D, aux, rows, compact, nullptr, demb, s_map, &s_lists);
}

// Small batches (B <= 256): merge + gradient in one launch.  Every block redoes the (tiny) merge of all anchors
// into its own shared memory and then runs the gradient gather for its 8 rows; block 0 also writes the loss,
// the mined indices and the statistics.  Saves one launch and the one-block merge kernel on the C1-size step.
constexpr int BH_FUSED_MAX_B = 256;
template <bool COSINE>
__global__ void __launch_bounds__(BH_GRAD_WARPS * 32) bh_merge_grad_kernel(
    const BhRec* __restrict__ recs, int n_splits, const float* __restrict__ x, const int32_t* __restrict__ labels, int B,
    int D, float alpha, int soft, const float* __restrict__ dloss, const float* __restrict__ aux, float* __restrict__ loss,
    int32_t* __restrict__ pos_idx_out, int32_t* __restrict__ neg_idx_out, float* __restrict__ stats,
    float* __restrict__ demb) {
  __shared__ BhRow s_rows[BH_FUSED_MAX_B];
  __shared__ int4 s_compact[BH_FUSED_MAX_B];
  bh_merge_body<COSINE>(recs, n_splits, B, alpha, soft, dloss, loss, pos_idx_out, neg_idx_out, stats, s_rows, s_compact,
                        blockIdx.x == 0);
  __syncthreads();
  bh_grad_body<COSINE, 0>(x, labels, B, D, aux, s_rows, s_compact, nullptr, demb);
}

// Small batches in ONE launch (B <= 128: the reference's own configuration is P x K = 18 x 4 = 72 rows of 128
// floats, common/losses.py:33-85 at training/triplet.py's batch shape).  At this size a step is a chain of
// global-memory round trips, so the whole step runs inside one thread-block CLUSTER of ceil(B / 8) CTAs and nothing
// but the inputs and the results touches global memory:
//   1. every CTA stages the whole batch in its shared memory (36 KB at C1) and normalises it (cosine);
//   2. CTA g mines anchors 8g..8g+7: its eight warps take interleaved 4-column steps of the canonical tile code
//      (bh_tile.cuh), the per-warp partial records are merged in shared memory;
//   3. the eight merged records (48 B each) are stored into EVERY CTA's shared memory (distributed shared memory,
//      st.shared::cluster) and one cluster barrier publishes them;
//   4. every CTA finalises all anchors from its copy (bh_merge_body with one "split": fillers, loss, coefficients,
//      statistics; CTA 0 writes the outputs) and runs the gradient gather for its own 8 rows.
// Same arithmetic, same records, same gradient code as the two-launch path (bit-identical indices and gradients; the
// statistics' float sums are folded in a different - fixed - order).
constexpr int BH_CL_WARPS = 8;    // (sixteen warps were tried: staging 0.7 us faster, mining and exchange 1.7 us slower)
// Canonical dot products out of a chain-major operand [32 chains][n_rows][4] (row index XOR-swizzled by the chain, see
// canon_mm.cuh) for TWO anchors (gi0, gi0 + 1) against rows lane, lane + 32, ... (NQ of them), HALF of the chains:
// half 0 = the sixteen even chains = the first sixteen of the bit-reversed order, whose counter tree ends in y_0 of the
// canonical butterfly; half 1 = the odd chains, y_1.
// The canonical total is y_0 + y_1.  Two anchors per thread halve the shared-memory traffic of the column operand
// (the mining bound at C1: every B-row quad was read by all eight warps).
template <int NQ>
__device__ __forceinline__ void chain_major_half_dots(const float* __restrict__ s_x, int n_rows, int gi0, int half, int lane,
                                                      float (&y)[2][4]) {
  float st[5][2][NQ];
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const int n = half * 16 + m;   // position in the bit-reversed order; the low four bits drive the counter
    const int l = ((n & 1) << 4) | ((n & 2) << 2) | (n & 4) | ((n & 8) >> 2) | ((n & 16) >> 4);
    const int sw = (l >> 2) & 7;
    float4 a[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) a[i] = *reinterpret_cast<const float4*>(s_x + ((size_t)l * n_rows + ((gi0 + i) ^ sw)) * 4);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int row = min(lane + 32 * q, n_rows - 1);
      const float4 bq = *reinterpret_cast<const float4*>(s_x + ((size_t)l * n_rows + (row ^ sw)) * 4);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float sacc = __fmaf_rn(a[i].x, bq.x, 0.f);
        sacc = __fmaf_rn(a[i].y, bq.y, sacc);
        sacc = __fmaf_rn(a[i].z, bq.z, sacc);
        sacc = __fmaf_rn(a[i].w, bq.w, sacc);
        int lvl = 0;
#pragma unroll
        for (int mm = m; mm & 1; mm >>= 1) {
          sacc = __fadd_rn(st[lvl][i][q], sacc);
          ++lvl;
        }
        st[lvl][i][q] = sacc;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int q = 0; q < NQ; ++q) y[i][q] = st[4][i][q];
}

constexpr int BH_CL_MAX_B = 128;
template <bool COSINE>
__global__ void __launch_bounds__(BH_CL_WARPS * 32) bh_cluster_step_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ labels, int B, int D, float alpha, int soft,
    const float* __restrict__ dloss, float* __restrict__ loss, int32_t* __restrict__ pos_idx_out,
    int32_t* __restrict__ neg_idx_out, float* __restrict__ stats, float* __restrict__ demb, int sliced,
    unsigned long long* __restrict__ trace /* development aid: [CTA][6] globaltimer stamps, NULL = off */) {
  extern __shared__ float sm[];
  auto stamp = [&](int k) {
    if (trace && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      trace[blockIdx.x * 6 + k] = t;
    }
  };
  stamp(0);
  const int G = (int)gridDim.x, Bp = G * BH_TI;
  // D <= 128: the mining operand is kept chain-major, [32 chains][Bp rows][4] (element d = l + 32 k of a row at
  // (l, row ^ swizzle(l), k)), so that ONE thread can run all 32 canonical chains of an entry with one 16-byte load per
  // chain and row (see canon_mm.cuh); wider rows keep the row-major warp tiles
  const bool chain_major = D <= 128;
  const int Dx = chain_major ? 128 : D;
  float* s_x = sm;                                         // [Bp][Dx] mining operand (cosine: normalised)
  float* s_raw = (COSINE || chain_major) ? s_x + (size_t)Bp * Dx : s_x;   // [Bp][D] the rows as given (the gradient
                                                           // code recomputes x * inv and must find x where it looks)
  float* s_aux = s_raw + (size_t)Bp * D;                   // [Bp] inverse norm | sum of squares
  int* s_lab = reinterpret_cast<int*>(s_aux + Bp);         // [Bp]
  BhRec* s_recs = reinterpret_cast<BhRec*>(s_lab + Bp);    // [Bp] merged record of every anchor (filled by all CTAs)
  BhRec* s_part = s_recs + Bp;                             // [BH_CL_WARPS][BH_TI] this CTA's per-warp partials
  __shared__ BhRow s_rows[BH_CL_MAX_B];
  __shared__ int4 s_compact[BH_CL_MAX_B];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = (int)blockIdx.x;
  const int kd = (D + 31) / 32;

  // 1. stage: ONE bulk copy (TMA engine, no registers, a single round trip) brings the whole batch into shared
  //    memory; then warp w normalises rows w, w + 8, ... out of it.  (Loading rows in register batches cost one
  //    global round trip per batch: 6.4 us of the step at C1.)
  __shared__ __align__(8) uint64_t s_bar;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
    const uint32_t bytes = (uint32_t)((size_t)B * D * 4);      // host: a multiple of 16, x 16-byte aligned
    mbar_arrive_expect_tx(&s_bar, bytes);
    if (!sliced) bulk_load_1d(s_raw, x, bytes, &s_bar);
  }
  if (sliced) {
    // the batch lives in page-locked HOST memory (dif_batch_hard_host): G whole-batch reads would cross PCIe G times,
    // so CTA g fetches only its own 8 rows and the copy engine multicasts them into every CTA of the cluster
    cluster_sync_all();   // every CTA's barrier is armed before the first slice can land
    if (threadIdx.x == 0) {
      const size_t off = (size_t)g * BH_TI * D;
      const uint32_t slice = (uint32_t)((size_t)min(BH_TI, B - g * BH_TI) * D * 4);   // host: D % 4 == 0
      bulk_load_1d_multicast(s_raw + off, x + off, slice, &s_bar, (uint16_t)((1u << G) - 1u));
    }
  }
  for (int r = B + warp; r < Bp; r += BH_CL_WARPS)             // rows past the batch are zero
    for (int d = lane; d < D; d += 32) s_raw[(size_t)r * D + d] = 0.f;
  for (int r = threadIdx.x; r < Bp; r += blockDim.x) s_lab[r] = r < B ? labels[r] : -2;
  __syncthreads();                                             // the barrier is initialised for everyone
  mbar_wait(&s_bar, 0);
  if (chain_major) {
    // three rows per trip: their load -> fma -> butterfly -> sqrt / divide chains run side by side (one row at a time
    // the nine rows of a warp were nine dependent chains in a row)
    constexpr int U = 3;
    for (int r0 = warp; r0 < Bp; r0 += U * BH_CL_WARPS) {
      float v[U][4], acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = min(r0 + u * BH_CL_WARPS, Bp - 1);
        acc[u] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int d = c * 32 + lane;
          v[u][c] = (c < kd && d < D) ? s_raw[(size_t)r * D + d] : 0.f;
          if (c < kd) acc[u] = __fmaf_rn(v[u][c], v[u][c], acc[u]);   // chain `lane`: d = lane, lane + 32, ... (dif_canon.cuh)
        }
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = __fadd_rn(acc[u], __shfl_xor_sync(0xffffffffu, acc[u], o));
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = r0 + u * BH_CL_WARPS;
        if (r >= Bp) break;   // warp-uniform
        const float ss = acc[u];
        const float inv = canon_inv_norm(ss);
        // lane = chain; (padding columns d >= D hold zeros: fma(0, 0, acc) = acc)
        float4 q;
        q.x = COSINE ? __fmul_rn(v[u][0], inv) : v[u][0];
        q.y = COSINE ? __fmul_rn(v[u][1], inv) : v[u][1];
        q.z = COSINE ? __fmul_rn(v[u][2], inv) : v[u][2];
        q.w = COSINE ? __fmul_rn(v[u][3], inv) : v[u][3];
        *reinterpret_cast<float4*>(s_x + ((size_t)lane * Bp + (r ^ ((lane >> 2) & 7))) * 4) = q;
        if (lane == 0) s_aux[r] = COSINE ? inv : ss;
      }
    }
  } else
  for (int r = warp; r < Bp; r += BH_CL_WARPS) {
    float v[BH_MAX_KD];
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < BH_MAX_KD; ++c) {
      const int d = c * 32 + lane;
      v[c] = (c < kd && d < D) ? s_raw[(size_t)r * D + d] : 0.f;
      if (c < kd) acc = __fmaf_rn(v[c], v[c], acc);            // chain `lane`: d = lane, lane + 32, ... as in dif_canon.cuh
    }
    const float ss = canon_tree(acc);
    const float inv = canon_inv_norm(ss);
    if (COSINE) {
#pragma unroll
      for (int c = 0; c < BH_MAX_KD; ++c) {
        const int d = c * 32 + lane;
        if (c < kd && d < D) s_x[(size_t)r * D + d] = __fmul_rn(v[c], inv);
      }
    }
    if (lane == 0) s_aux[r] = COSINE ? inv : ss;
  }
  __syncthreads();
  stamp(1);

  // 2. mine anchors 8g .. 8g + 7
  if (chain_major) {
    // warp w = anchor 8g + w, lane t = columns t, t + 32, ... : every thread runs the 32 chains of its (up to four)
    // entries in bit-reversed order and folds them with the counter tree of canon_mm.cuh - 128 fma + 31 add per entry,
    // no shuffles, no dependent tile steps (the warp-tile code below spends 4.5 us here at B = 72).  The anchor's
    // record is then one butterfly over the warp.
    const int gi = g * BH_TI + warp;
    const int my_lab = gi < B ? s_lab[gi] : -1;
    const float my_aux = s_aux[gi];
    const int nq = (B + 31) >> 5;   // <= 4
    float dots[4] = {0.f, 0.f, 0.f, 0.f};
    // (the column count is a template argument: straight-line code, so the loads of later chains overlap the fma
    // chains of earlier ones - with a run-time bound every chain waited for its own loads: 5.8 us instead of 4.5)
    {
      // warp (pair p, half h) runs half of the chains for anchors 2p, 2p + 1; the halves meet in shared memory
      __shared__ float s_y[2][BH_TI][BH_CL_MAX_B];
      const int pair = warp & 3, half = warp >> 2;
      float y[2][4];
      if (nq <= 1) chain_major_half_dots<1>(s_x, Bp, g * BH_TI + 2 * pair, half, lane, y);
      else if (nq == 2) chain_major_half_dots<2>(s_x, Bp, g * BH_TI + 2 * pair, half, lane, y);
      else if (nq == 3) chain_major_half_dots<3>(s_x, Bp, g * BH_TI + 2 * pair, half, lane, y);
      else chain_major_half_dots<4>(s_x, Bp, g * BH_TI + 2 * pair, half, lane, y);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nq) s_y[half][2 * pair + i][lane + 32 * q] = y[i][q];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < nq) dots[q] = __fadd_rn(s_y[0][warp][lane + 32 * q], s_y[1][warp][lane + 32 * q]);
    }
    float pos_val = COSINE ? INFINITY : -INFINITY, neg_val = COSINE ? -INFINITY : INFINITY, all_max = -INFINITY;
    int pos_idx = -1, neg_idx = -1, all_idx = -1, pos_cnt = 0, neg_cnt = 0, all_cnt = 0, n_pos = 0;
    float row_sum = 0.f, pos_sum = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int gj = lane + 32 * q;
      if (q < nq && gj < B && gi < B) {
        const float dot = dots[q];
        const float dist = COSINE ? dot : __fsub_rn(__fadd_rn(my_aux, s_aux[gj]), __fmul_rn(2.f, dot));
        row_sum += dist;
        fold<false>(dist, gj, all_max, all_idx, all_cnt);
        if (s_lab[gj] == my_lab) {
          ++n_pos;
          pos_sum += dist;
          fold<COSINE>(dist, gj, pos_val, pos_idx, pos_cnt);
        } else {
          fold<!COSINE>(dist, gj, neg_val, neg_idx, neg_cnt);
        }
      }
    }
#pragma unroll
    for (int o = 1; o <= 16; o <<= 1) {
      const float pv = __shfl_xor_sync(0xffffffffu, pos_val, o);
      const int pi = __shfl_xor_sync(0xffffffffu, pos_idx, o), pc = __shfl_xor_sync(0xffffffffu, pos_cnt, o);
      const float nv = __shfl_xor_sync(0xffffffffu, neg_val, o);
      const int ni = __shfl_xor_sync(0xffffffffu, neg_idx, o), nc = __shfl_xor_sync(0xffffffffu, neg_cnt, o);
      const float av = __shfl_xor_sync(0xffffffffu, all_max, o);
      const int ai = __shfl_xor_sync(0xffffffffu, all_idx, o), ac = __shfl_xor_sync(0xffffffffu, all_cnt, o);
      merge<COSINE>(pv, pi, pc, pos_val, pos_idx, pos_cnt);
      merge<!COSINE>(nv, ni, nc, neg_val, neg_idx, neg_cnt);
      merge<false>(av, ai, ac, all_max, all_idx, all_cnt);
      row_sum = __fadd_rn(row_sum, __shfl_xor_sync(0xffffffffu, row_sum, o));
      pos_sum = __fadd_rn(pos_sum, __shfl_xor_sync(0xffffffffu, pos_sum, o));
      n_pos += __shfl_xor_sync(0xffffffffu, n_pos, o);
    }
    if (lane == 0) {   // already the anchor's merged record: the eight-warp merge below has nothing to add
      BhRec r;
      r.pos_val = pos_val; r.pos_idx = pos_idx; r.pos_cnt = pos_cnt;
      r.neg_val = neg_val; r.neg_idx = neg_idx; r.neg_cnt = neg_cnt;
      r.all_max = all_max; r.all_idx = all_idx; r.all_cnt = all_cnt;
      r.row_sum = row_sum; r.n_pos = n_pos; r.pos_sum = pos_sum;
      s_part[warp] = r;
    }
  } else {
    // warp w takes the column steps 4w, 4w + 32, ...
    const int my_i = lane >> 2, gi = g * BH_TI + my_i;
    const int my_lab = gi < B ? s_lab[gi] : -1;
    const float my_aux = s_aux[gi];
    float pos_val = COSINE ? INFINITY : -INFINITY, neg_val = COSINE ? -INFINITY : INFINITY, all_max = -INFINITY;
    int pos_idx = -1, neg_idx = -1, all_idx = -1, pos_cnt = 0, neg_cnt = 0, all_cnt = 0, n_pos = 0;
    float row_sum = 0.f, pos_sum = 0.f;
    for (int j0 = warp * BH_TJ; j0 < B; j0 += BH_CL_WARPS * BH_TJ) {
      const float dot = tile_step_dot(s_x + (size_t)g * BH_TI * D, s_x, D, kd, 0, lane, j0);
      const int gj = j0 + (lane & 3);
      if (gj < B && gi < B) {
        const float dist = COSINE ? dot : __fsub_rn(__fadd_rn(my_aux, s_aux[gj]), __fmul_rn(2.f, dot));
        row_sum += dist;
        fold<false>(dist, gj, all_max, all_idx, all_cnt);
        if (s_lab[gj] == my_lab) {
          ++n_pos;
          pos_sum += dist;
          fold<COSINE>(dist, gj, pos_val, pos_idx, pos_cnt);
        } else {
          fold<!COSINE>(dist, gj, neg_val, neg_idx, neg_cnt);
        }
      }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {   // the 4 lanes of an anchor
      const float pv = __shfl_xor_sync(0xffffffffu, pos_val, o);
      const int pi = __shfl_xor_sync(0xffffffffu, pos_idx, o), pc = __shfl_xor_sync(0xffffffffu, pos_cnt, o);
      const float nv = __shfl_xor_sync(0xffffffffu, neg_val, o);
      const int ni = __shfl_xor_sync(0xffffffffu, neg_idx, o), nc = __shfl_xor_sync(0xffffffffu, neg_cnt, o);
      const float av = __shfl_xor_sync(0xffffffffu, all_max, o);
      const int ai = __shfl_xor_sync(0xffffffffu, all_idx, o), ac = __shfl_xor_sync(0xffffffffu, all_cnt, o);
      merge<COSINE>(pv, pi, pc, pos_val, pos_idx, pos_cnt);
      merge<!COSINE>(nv, ni, nc, neg_val, neg_idx, neg_cnt);
      merge<false>(av, ai, ac, all_max, all_idx, all_cnt);
      row_sum = __fadd_rn(row_sum, __shfl_xor_sync(0xffffffffu, row_sum, o));
      pos_sum = __fadd_rn(pos_sum, __shfl_xor_sync(0xffffffffu, pos_sum, o));
      n_pos += __shfl_xor_sync(0xffffffffu, n_pos, o);
    }
    if ((lane & 3) == 0) {
      BhRec r;
      r.pos_val = pos_val; r.pos_idx = pos_idx; r.pos_cnt = pos_cnt;
      r.neg_val = neg_val; r.neg_idx = neg_idx; r.neg_cnt = neg_cnt;
      r.all_max = all_max; r.all_idx = all_idx; r.all_cnt = all_cnt;
      r.row_sum = row_sum; r.n_pos = n_pos; r.pos_sum = pos_sum;
      s_part[warp * BH_TI + my_i] = r;
    }
  }
  __syncthreads();
  stamp(2);
  // merge the eight warps' partials of each anchor (fixed warp order), then hand the record to every CTA
  if (!chain_major && threadIdx.x < BH_TI) {
    const int a = threadIdx.x;
    BhRec m = s_part[a];
    for (int w = 1; w < BH_CL_WARPS; ++w) {
      const BhRec& q = s_part[w * BH_TI + a];
      merge<COSINE>(q.pos_val, q.pos_idx, q.pos_cnt, m.pos_val, m.pos_idx, m.pos_cnt);
      merge<!COSINE>(q.neg_val, q.neg_idx, q.neg_cnt, m.neg_val, m.neg_idx, m.neg_cnt);
      merge<false>(q.all_max, q.all_idx, q.all_cnt, m.all_max, m.all_idx, m.all_cnt);
      m.row_sum = __fadd_rn(m.row_sum, q.row_sum);
      m.pos_sum = __fadd_rn(m.pos_sum, q.pos_sum);
      m.n_pos += q.n_pos;
    }
    s_part[a] = m;
  }
  __syncthreads();
  {
    constexpr int kWords = (int)(sizeof(BhRec) / 4) * BH_TI;   // the CTA's eight merged records
    const uint32_t* src = reinterpret_cast<const uint32_t*>(s_part);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_recs + (size_t)g * BH_TI);
    for (int t = threadIdx.x; t < kWords * G; t += blockDim.x) {
      const int cta = t / kWords, w = t - cta * kWords;
      st_cluster_u32(dst + w, (uint32_t)cta, src[w]);
    }
  }
  cluster_sync_all();   // release / acquire: every CTA now holds all Bp merged records
  stamp(3);

  // 4. finalize every anchor locally (one "split"), CTA 0 writes loss / indices / statistics; then the gradient
  bh_merge_body<COSINE>(s_recs, 1, B, alpha, soft, dloss, loss, pos_idx_out, neg_idx_out, stats, s_rows, s_compact, g == 0);
  __syncthreads();
  stamp(4);
  // the gradient gather reads rows, labels and norms out of shared memory: at this size its cost is the latency of
  // its dependent loads (mined index -> row -> ...), a few cycles here against an L2 round trip each
  if (demb) bh_grad_body<COSINE, 0>(s_raw, s_lab, B, D, s_aux, s_rows, s_compact, nullptr, demb);
  stamp(5);
}

// one-hot [B, C] -> int32 class ids (tf.argmax(labels, axis=1): first maximum), losses.py:35
__global__ void argmax_rows_kernel(const float* __restrict__ onehot, int B, int C, int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= B) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float v = onehot[(size_t)r * C + c];
    if (v > best || (v == best && c < bi)) {
      best = v;
      bi = c;
    }
  }
  for (int o = 16; o >= 1; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if (lane == 0) out[r] = bi == 0x7fffffff ? 0 : bi;
}

// workspace cache (per thread): split records, per-row info, row aux
struct BhWorkspace {
  BhRec* recs = nullptr;
  size_t rec_cap = 0;
  BhRow* rows = nullptr;
  float* aux = nullptr;
  int4* compact = nullptr;
  double* partials = nullptr;              // [ceil(rows / 256)][5]
  unsigned long long* gmax_key = nullptr;  // [1] + float cg right behind it
  size_t row_cap = 0;
  int ensure(size_t n_rec, size_t n_rows) {
    if (n_rec > rec_cap) {
      retire_device_block(recs);
      recs = nullptr;
      rec_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&recs, n_rec * sizeof(BhRec)));
      rec_cap = n_rec;
    }
    if (n_rows > row_cap) {
      retire_device_block(rows);
      retire_device_block(aux);
      retire_device_block(compact);
      retire_device_block(partials);
      rows = nullptr;
      aux = nullptr;
      compact = nullptr;
      partials = nullptr;
      row_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&rows, n_rows * sizeof(BhRow)));
      DIF_CUDA_OK(cudaMalloc((void**)&aux, n_rows * sizeof(float)));
      DIF_CUDA_OK(cudaMalloc((void**)&compact, n_rows * sizeof(int4)));
      DIF_CUDA_OK(cudaMalloc((void**)&partials, ((n_rows + 7) / 8) * 5 * sizeof(double)));
      if (!gmax_key) DIF_CUDA_OK(cudaMalloc((void**)&gmax_key, 16));
      row_cap = n_rows;
    }
    return DIF_OK;
  }
};
static thread_local BhWorkspace g_ws;

struct BhHostStage {
  void* h = nullptr;
  void* d = nullptr;
  size_t bytes = 0;
  cudaStream_t st = nullptr;
  int ensure(size_t need) {
    if (!st) DIF_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    if (need <= bytes) return DIF_OK;
    if (h) cudaFreeHost(h);
    if (d) cudaFree(d);
    h = d = nullptr;
    bytes = 0;
    DIF_CUDA_OK(cudaMallocHost(&h, need));
    DIF_CUDA_OK(cudaMalloc(&d, need));
    bytes = need;
    return DIF_OK;
  }
};
static thread_local BhHostStage g_bh_stage;

static thread_local bool g_bh_host_input = false;   // set by dif_batch_hard_host around its zero-copy call
static int g_bh_force_path = 0;   // 0 auto, 1 CUDA-core miner, 2 tensor-core miner (tests)

template <bool COSINE>
static int run_batch_hard(const float* emb, const int32_t* labels, int B, int D, float alpha, int soft, float* loss,
                          int32_t* pos_idx, int32_t* neg_idx, float* stats, const float* dloss, float* demb,
                          cudaStream_t st) {
  // column splits: enough blocks to cover the machine, at least one step of 4 columns each
  const int row_blocks = (B + BH_RB - 1) / BH_RB;
  const int sms = std::max(1, device_sm_count());
  // at least one full staged tile (32 columns) per split: finer splits only add staging and merge work
  int splits = std::max(1, std::min((2 * sms + row_blocks - 1) / row_blocks, (B + BH_CB - 1) / BH_CB));
  splits = std::min(splits, 64);
  int cols = (B + splits - 1) / splits;
  cols = (cols + BH_TJ - 1) / BH_TJ * BH_TJ;
  splits = (B + cols - 1) / cols;
  if (int rc = g_ws.ensure((size_t)splits * B, (size_t)B)) return rc;
  const size_t smem = ((size_t)(BH_RB + BH_CB) * D + BH_RB + BH_CB) * sizeof(float) + BH_CB * sizeof(int);
  static bool configured = false;
  if (!configured) {
    DIF_CUDA_OK(cudaFuncSetAttribute(bh_mine_kernel<COSINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  // small batches: the whole step in one cluster launch
  if (g_bh_force_path == 0 || g_bh_force_path == 3) {
    const int G = (B + BH_TI - 1) / BH_TI;
    // mining operand (chain-major, padded to 128 columns, for D <= 128) + the rows as given (shared with the operand
    // only for the squared-L2 loss on wide rows) + norms, labels, records
    const size_t cl_rows = D <= 128 ? (size_t)G * BH_TI * (128 + D) : (size_t)G * BH_TI * D * (COSINE ? 2 : 1);
    const size_t cl_smem = (cl_rows + 2 * (size_t)G * BH_TI) * 4 + ((size_t)G * BH_TI + BH_CL_WARPS * BH_TI) * sizeof(BhRec);
    if (B <= BH_CL_MAX_B && G <= 16 && D <= 32 * BH_MAX_KD && cl_smem <= 160 * 1024 && ((size_t)B * D * 4) % 16 == 0 &&
        (reinterpret_cast<uintptr_t>(emb) & 15u) == 0) {
      static int cluster_ok = -1;   // per instantiation: can a cluster of this kernel be scheduled at all?
      auto kern = bh_cluster_step_kernel<COSINE>;
      if (cluster_ok < 0) {
        cluster_ok = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) == cudaSuccess &&
                     cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        if (!cluster_ok) cudaGetLastError();
      }
      if (cluster_ok) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)G);
        cfg.blockDim = dim3(BH_CL_WARPS * 32);
        cfg.dynamicSmemBytes = cl_smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)G;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        static const bool profile = getenv("DIF_BH_PROFILE") != nullptr;   // development aid
        static unsigned long long* trace_d = nullptr;
        if (profile && !trace_d) cudaMalloc((void**)&trace_d, 16 * 6 * 8);
        static const bool force_sliced = getenv("DIF_BH_SLICED") != nullptr;   // A/B switch
        const int sliced = ((g_bh_host_input || force_sliced) && D % 4 == 0) ? 1 : 0;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, emb, labels, B, D, alpha, soft, dloss, loss, pos_idx, neg_idx, stats, demb,
                                                 sliced, profile ? trace_d : (unsigned long long*)nullptr);
        if (e == cudaSuccess) {
          count_launch();
          if (profile) {
            unsigned long long th[16 * 6];
            cudaStreamSynchronize(st);
            cudaMemcpy(th, trace_d, sizeof(th), cudaMemcpyDeviceToHost);
            unsigned long long t0 = ~0ull;
            for (int c = 0; c < G; ++c) t0 = std::min(t0, th[c * 6]);
            static const char* nm[6] = {"entry", "staged", "mined", "exchanged", "finalized", "gradient"};
            fprintf(stderr, "bh cluster step (us after the first CTA started, CTA 0 / max over CTAs):");
            for (int k = 0; k < 6; ++k) {
              unsigned long long mx = 0;
              for (int c = 0; c < G; ++c) mx = std::max(mx, th[c * 6 + k]);
              fprintf(stderr, " %s %.1f/%.1f", nm[k], (double)(th[k] - t0) * 1e-3, (double)(mx - t0) * 1e-3);
            }
            fprintf(stderr, "\n");
          }
          return DIF_OK;
        }
        cudaGetLastError();   // (a cluster this large does not fit this device / partition: use the two-launch path)
        cluster_ok = 0;
      }
    }
  }
  // large batches: tensor-core filter + canonical re-rank (bh_tc.cu), then a multi-block finalize
  // (measured, D = 128: cosine B = 256 30.8 us either way, B = 384 32.6 us on the tensor cores vs 39.0; squared-L2 B = 384
  // 36.7 vs 32.3 - its epilogue and finalize are heavier - so the two losses switch at different sizes)
  const bool tensor_path = (g_bh_force_path == 2 || (g_bh_force_path == 0 && B >= (COSINE ? 320 : 512))) && D >= 32 && D % 4 == 0;
  const float* cg_dev = nullptr;
  if (tensor_path) {
    DIF_CUDA_OK(cudaMemsetAsync(g_ws.gmax_key, 0, 16, st));
    // cosine: the re-rank launch finalizes as well (one partial-sum record per 8 anchors); squared-L2 needs max(dists) first
    BhFinalize fin{alpha, soft, dloss, loss, pos_idx, neg_idx, COSINE ? g_ws.rows : nullptr, g_ws.compact, g_ws.partials};
    if (int rc = bh_mine_tensor<COSINE>(emb, labels, B, D, g_ws.recs, g_ws.aux, g_ws.gmax_key, fin, st)) return rc;
    int fb = (B + kBhFusedFinalizeRows - 1) / kBhFusedFinalizeRows;
    if (!COSINE) {
      fb = (B + 255) / 256;
      bh_finalize_kernel<COSINE><<<fb, 256, 0, st>>>(g_ws.recs, B, alpha, soft, dloss, g_ws.gmax_key, loss, pos_idx, neg_idx,
                                                    g_ws.rows, g_ws.compact, g_ws.partials);
      DIF_LAUNCH_OK();
    }
    float* cg = reinterpret_cast<float*>(g_ws.gmax_key + 1);
    cg_dev = cg;
    if (demb && B <= BH_MAP_MAX_B) {
      const size_t map_smem = (size_t)(BH_GRAD_WARPS + 1) * ((B + 31) / 32) * sizeof(unsigned);
      bh_grad_map_kernel<COSINE><<<(B + BH_GRAD_WARPS - 1) / BH_GRAD_WARPS, BH_GRAD_WARPS * 32, map_smem, st>>>(
          emb, labels, B, D, g_ws.aux, g_ws.rows, g_ws.compact, g_ws.partials, fb, g_ws.gmax_key, stats, demb);
      DIF_LAUNCH_OK();
      return DIF_OK;
    }
    bh_stats_kernel<<<1, 32, 0, st>>>(g_ws.partials, fb, B, g_ws.gmax_key, stats, cg);
    DIF_LAUNCH_OK();
  } else {
    bh_mine_kernel<COSINE><<<dim3(row_blocks, splits), BH_WARPS * 32, smem, st>>>(emb, labels, B, D, cols, g_ws.recs, g_ws.aux);
    DIF_LAUNCH_OK();
    if (demb && B <= BH_FUSED_MAX_B) {
      bh_merge_grad_kernel<COSINE><<<(B + BH_GRAD_WARPS - 1) / BH_GRAD_WARPS, BH_GRAD_WARPS * 32, 0, st>>>(
          g_ws.recs, splits, emb, labels, B, D, alpha, soft, dloss, g_ws.aux, loss, pos_idx, neg_idx, stats, demb);
      DIF_LAUNCH_OK();
      return DIF_OK;
    }
    const int merge_threads = std::min(1024, std::max(64, (B + 31) / 32 * 32));
    bh_merge_kernel<COSINE><<<1, merge_threads, 0, st>>>(g_ws.recs, splits, B, alpha, soft, dloss, loss, pos_idx, neg_idx, stats,
                                                       g_ws.rows, g_ws.compact);
    DIF_LAUNCH_OK();
  }
  if (demb) {
    bh_grad_kernel<COSINE><<<(B + BH_GRAD_WARPS - 1) / BH_GRAD_WARPS, BH_GRAD_WARPS * 32, 0, st>>>(
        emb, labels, B, D, g_ws.aux, g_ws.rows, g_ws.compact, cg_dev, demb);
    DIF_LAUNCH_OK();
  }
  return DIF_OK;
}


// =============================================================================================
// Batch-all triplet loss, deep_insight_face/common/losses.py:131-148 (cosine):
//   pos_loss_i = sum_j (1 - where(pos, S, 1)) / n_pos_i;  hp_i = min_j where(pos, S, 1)
//   valid_ij   = !pos_ij && (hp_i - S_ij < alpha);  neg_loss_i = sum_j where(valid, S, 0) / (n_valid_i + 1)
// Pass 1 is bh_mine_kernel<true> (hp, n_pos, sum of positive similarities); pass 2 needs hp_i and counts the
// valid negatives; the backward pass recomputes S tile by tile and accumulates (G + G^T) N in registers
// (the `valid` mask is piecewise constant, so no gradient flows through hp).
// =============================================================================================
struct BaRow {
  float hp;      // min over positives and the 1.0 fillers of the non-positive columns
  float cp;      // dL/dS applied to every positive column of this anchor   (-g / n_pos)
  float cn;      // dL/dS applied to every valid negative column            (+g / (n_valid + 1))
  int n_pos;
};

__global__ void __launch_bounds__(256) ba_merge1_kernel(const BhRec* __restrict__ recs, int n_splits, int B,
                                                        BaRow* __restrict__ rows, float* __restrict__ pos_loss) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float hp = INFINITY, ps = 0.f;
  int idx = -1, cnt = 0, n_pos = 0;
  for (int s = 0; s < n_splits; ++s) {
    const BhRec& q = recs[(size_t)s * B + i];
    merge<true>(q.pos_val, q.pos_idx, q.pos_cnt, hp, idx, cnt);
    ps = __fadd_rn(ps, q.pos_sum);
    n_pos += q.n_pos;
  }
  if (n_pos < B) hp = fminf(hp, 1.f);
  BaRow r;
  r.hp = hp;
  r.cp = r.cn = 0.f;
  r.n_pos = n_pos;
  rows[i] = r;
  pos_loss[i] = ((float)n_pos - ps) / (float)n_pos;   // sum over positives of (1 - S)
}

__global__ void __launch_bounds__(BH_WARPS * 32) ba_valid_kernel(const float* __restrict__ x,
                                                                 const int32_t* __restrict__ labels, int B, int D,
                                                                 int cols_per_split, const BaRow* __restrict__ rows,
                                                                 float alpha, float2* __restrict__ vrec) {
  extern __shared__ float sm[];
  float* sa = sm;
  float* sb = sa + BH_RB * D;
  float* aux_a = sb + BH_CB * D;
  float* aux_b = aux_a + BH_RB;
  int* lab_b = reinterpret_cast<int*>(aux_b + BH_CB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * BH_RB, split = blockIdx.y;
  const int c_begin = split * cols_per_split, c_end = min(B, c_begin + cols_per_split);
  const int kd = (D + 31) / 32;
  stage_rows<true>(x, B, D, row0, BH_RB, sa, aux_a);
  __syncthreads();
  const int gi = row0 + warp * BH_TI + (lane >> 2);
  const int my_lab = gi < B ? labels[gi] : -1;
  const float my_hp = gi < B ? rows[gi].hp : 0.f;
  float sum_v = 0.f;
  int n_v = 0;
  for (int c0 = c_begin; c0 < c_end; c0 += BH_CB) {
    __syncthreads();
    stage_rows<true>(x, B, D, c0, BH_CB, sb, aux_b);
    if (threadIdx.x < BH_CB) lab_b[threadIdx.x] = (c0 + threadIdx.x < B) ? labels[c0 + threadIdx.x] : -2;
    __syncthreads();
    const int steps = min(BH_CB, c_end - c0);
    for (int j0 = 0; j0 < steps; j0 += BH_TJ) {
      const float sim = tile_step_dot(sa, sb, D, kd, warp, lane, j0);
      const int jl = j0 + (lane & 3), gj = c0 + jl;
      if (gj < c_end && gi < B && lab_b[jl] != my_lab && __fsub_rn(my_hp, sim) < alpha) {
        sum_v += sim;
        ++n_v;
      }
    }
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    sum_v = __fadd_rn(sum_v, __shfl_xor_sync(0xffffffffu, sum_v, o));
    n_v += __shfl_xor_sync(0xffffffffu, n_v, o);
  }
  if ((lane & 3) == 0 && gi < B) vrec[(size_t)split * B + gi] = make_float2(sum_v, __int_as_float(n_v));
}

__global__ void __launch_bounds__(256) ba_merge2_kernel(const float2* __restrict__ vrec, int n_splits, int B,
                                                        const float* __restrict__ pos_loss,
                                                        const float* __restrict__ dloss, float* __restrict__ loss,
                                                        BaRow* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float sv = 0.f;
  int nv = 0;
  for (int s = 0; s < n_splits; ++s) {
    const float2 q = vrec[(size_t)s * B + i];
    sv = __fadd_rn(sv, q.x);
    nv += __float_as_int(q.y);
  }
  loss[i] = pos_loss[i] + sv / ((float)nv + 1.f);
  const float g = dloss ? dloss[i] : 1.f / (float)B;
  rows[i].cp = -g / (float)rows[i].n_pos;
  rows[i].cn = g / ((float)nv + 1.f);
}

template <int KD>
__global__ void __launch_bounds__(BH_WARPS * 32) ba_grad_kernel(const float* __restrict__ x,
                                                                const int32_t* __restrict__ labels, int B, int D,
                                                                const BaRow* __restrict__ rows, float alpha,
                                                                float* __restrict__ demb) {
  extern __shared__ float sm[];
  float* sa = sm;
  float* sb = sa + BH_RB * D;
  float* aux_a = sb + BH_CB * D;
  float* aux_b = aux_a + BH_RB;
  int* lab_b = reinterpret_cast<int*>(aux_b + BH_CB);
  float* hp_b = reinterpret_cast<float*>(lab_b + BH_CB);
  float* cp_b = hp_b + BH_CB;
  float* cn_b = cp_b + BH_CB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * BH_RB;
  const int kd = (D + 31) / 32;
  stage_rows<true>(x, B, D, row0, BH_RB, sa, aux_a);
  __syncthreads();
  const int gi = row0 + warp * BH_TI + (lane >> 2);
  const int my_lab = gi < B ? labels[gi] : -1;
  BaRow me;
  me.hp = me.cp = me.cn = 0.f;
  me.n_pos = 0;
  if (gi < B) me = rows[gi];
  float acc[BH_TI][KD];
#pragma unroll
  for (int i = 0; i < BH_TI; ++i)
#pragma unroll
    for (int c = 0; c < KD; ++c) acc[i][c] = 0.f;

  for (int c0 = 0; c0 < B; c0 += BH_CB) {
    __syncthreads();
    stage_rows<true>(x, B, D, c0, BH_CB, sb, aux_b);
    if (threadIdx.x < BH_CB) {
      const int gj = c0 + threadIdx.x;
      lab_b[threadIdx.x] = gj < B ? labels[gj] : -2;
      hp_b[threadIdx.x] = gj < B ? rows[gj].hp : 0.f;
      cp_b[threadIdx.x] = gj < B ? rows[gj].cp : 0.f;
      cn_b[threadIdx.x] = gj < B ? rows[gj].cn : 0.f;
    }
    __syncthreads();
    const int steps = min(BH_CB, B - c0);
    for (int j0 = 0; j0 < steps; j0 += BH_TJ) {
      const float sim = tile_step_dot(sa, sb, D, kd, warp, lane, j0);
      const int jl = j0 + (lane & 3), gj = c0 + jl;
      float w = 0.f;
      if (gj < B && gi < B) {
        if (lab_b[jl] == my_lab) {
          w = me.cp + cp_b[jl];
        } else {
          if (__fsub_rn(me.hp, sim) < alpha) w += me.cn;          // j is a valid negative of anchor i
          if (__fsub_rn(hp_b[jl], sim) < alpha) w += cn_b[jl];    // i is a valid negative of anchor j
        }
      }
      // acc[i][:] += w(i, j) * n_j for the 32 entries of this step
#pragma unroll
      for (int j = 0; j < BH_TJ; ++j) {
        float nb[KD];
#pragma unroll
        for (int c = 0; c < KD; ++c) {
          const int d = c * 32 + lane;
          nb[c] = (c < kd && d < D) ? sb[(j0 + j) * D + d] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < BH_TI; ++i) {
          const float we = __shfl_sync(0xffffffffu, w, i * BH_TJ + j);
#pragma unroll
          for (int c = 0; c < KD; ++c) acc[i][c] = fmaf(we, nb[c], acc[i][c]);
        }
      }
    }
  }
  // l2_normalize backward per anchor: dx = inv * (dn - n (n . dn))
#pragma unroll
  for (int i = 0; i < BH_TI; ++i) {
    const int r = row0 + warp * BH_TI + i;
    if (r >= B) continue;   // warp-uniform
    const float* n = sa + (warp * BH_TI + i) * D;
    float dotp = 0.f;
#pragma unroll
    for (int c = 0; c < KD; ++c) {
      const int d = c * 32 + lane;
      if (c < kd && d < D) dotp += acc[i][c] * n[d];
    }
    for (int o = 16; o >= 1; o >>= 1) dotp += __shfl_xor_sync(0xffffffffu, dotp, o);
    const float inv = aux_a[warp * BH_TI + i];
    const bool clamped = inv >= 0.99e6f;
#pragma unroll
    for (int c = 0; c < KD; ++c) {
      const int d = c * 32 + lane;
      if (c < kd && d < D) demb[(size_t)r * D + d] = clamped ? inv * acc[i][c] : inv * (acc[i][c] - n[d] * dotp);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Large batches (B >= 256, D <= 128): the three passes above each recompute S with warp tiles (and the gradient runs
// on B / 32 blocks).  Here S is produced ONCE by the one-thread-per-entry canonical kernel (canon_mm.cuh, the same
// products in the same order, so `hp - S < alpha` decides exactly as before), a warp per anchor reduces its row, and
// the gradient (G + G^T) N is a warp per anchor streaming N: B=1024 516 -> ~40 us.
// ---------------------------------------------------------------------------------------------
struct BaDotEpi {
  __device__ __forceinline__ float operator()(int, int, float dot) const { return dot; }
};

__global__ void __launch_bounds__(256) ba_norm_kernel(const float* __restrict__ x, int B, int D, float* __restrict__ xn,
                                                      float* __restrict__ inv_out) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= B) return;
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float t = x[(size_t)r * D + d];
    acc = __fmaf_rn(t, t, acc);
  }
  const float inv = canon_inv_norm(canon_tree(acc));
  for (int d = lane; d < D; d += 32) xn[(size_t)r * D + d] = __fmul_rn(x[(size_t)r * D + d], inv);
  if (lane == 0) inv_out[r] = inv;
}

constexpr int BA_ROW_WARPS = 8;
__global__ void __launch_bounds__(BA_ROW_WARPS * 32) ba_row_kernel(const float* __restrict__ S, int lds,
                                                                   const int32_t* __restrict__ labels, int B, float alpha,
                                                                   const float* __restrict__ dloss, float* __restrict__ loss,
                                                                   BaRow* __restrict__ rows) {
  const int i = blockIdx.x * BA_ROW_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= B) return;
  const float* row = S + (size_t)i * lds;
  const int my_lab = labels[i];
  float hp = INFINITY, ps = 0.f;
  int n_pos = 0;
  for (int j = lane; j < B; j += 32) {
    if (labels[j] == my_lab) {
      const float v = row[j];
      hp = fminf(hp, v);
      ps += v;
      ++n_pos;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    hp = fminf(hp, __shfl_xor_sync(0xffffffffu, hp, o));
    ps += __shfl_xor_sync(0xffffffffu, ps, o);
    n_pos += __shfl_xor_sync(0xffffffffu, n_pos, o);
  }
  if (n_pos < B) hp = fminf(hp, 1.f);   // the 1.0 fillers of the non-positive columns
  float sv = 0.f;
  int nv = 0;
  for (int j = lane; j < B; j += 32) {
    const float v = row[j];
    if (labels[j] != my_lab && __fsub_rn(hp, v) < alpha) {
      sv += v;
      ++nv;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    sv += __shfl_xor_sync(0xffffffffu, sv, o);
    nv += __shfl_xor_sync(0xffffffffu, nv, o);
  }
  if (lane == 0) {
    const float g = dloss ? dloss[i] : 1.f / (float)B;
    loss[i] = ((float)n_pos - ps) / (float)n_pos + sv / ((float)nv + 1.f);
    BaRow r;
    r.hp = hp;
    r.cp = -g / (float)n_pos;
    r.cn = g / ((float)nv + 1.f);
    r.n_pos = n_pos;
    rows[i] = r;
  }
}

// dN = (G + G^T) N as a split-K product: a warp owns 4 anchors and one range of columns j, keeps the weights of a
// 128-column chunk in shared memory, streams the rows n_j (each load feeds 4 anchors) and leaves a partial sum;
// ba_grad_reduce_kernel folds the ranges in a fixed order and applies the l2_normalize backward.
constexpr int BA_GRAD_CHUNK = 128;
constexpr int BA_GRAD_ROWS = 4;      // anchors per warp
template <int KQ>   // 32-float chunks per row: 4 (D <= 128), 8 (<= 256), 16 (<= 512)
__global__ void __launch_bounds__(BA_ROW_WARPS * 32) ba_grad_fast_kernel(const float* __restrict__ S, int lds,
                                                                         const float* __restrict__ xn,
                                                                         const int32_t* __restrict__ labels, int B, int D,
                                                                         const BaRow* __restrict__ rows, float alpha,
                                                                         int cols_per_split, float* __restrict__ part) {
  __shared__ float s_w[BA_ROW_WARPS][BA_GRAD_ROWS][BA_GRAD_CHUNK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = (blockIdx.x * BA_ROW_WARPS + warp) * BA_GRAD_ROWS;
  if (i0 >= B) return;   // warp-uniform; warps never synchronise with each other
  const int c_begin = blockIdx.y * cols_per_split, c_end = min(B, c_begin + cols_per_split);
  int my_lab[BA_GRAD_ROWS];
  BaRow me[BA_GRAD_ROWS];
#pragma unroll
  for (int a = 0; a < BA_GRAD_ROWS; ++a) {
    const int i = min(i0 + a, B - 1);
    my_lab[a] = labels[i];
    me[a] = rows[i];
  }
  constexpr int U = 16 / KQ;   // rows n_j in flight per trip
  unsigned long long acc[BA_GRAD_ROWS][KQ / 2];   // (d, d + 1) pairs
#pragma unroll
  for (int a = 0; a < BA_GRAD_ROWS; ++a)
#pragma unroll
    for (int c = 0; c < KQ / 2; ++c) acc[a][c] = 0ull;
  for (int j0 = c_begin; j0 < c_end; j0 += BA_GRAD_CHUNK) {
    const int n = min(BA_GRAD_CHUNK, c_end - j0);
    __syncwarp();
    // w(i, j) = G_ij + G_ji: positives -1/n_pos each way, valid negatives +1/(n_valid + 1) each way
#pragma unroll
    for (int q = 0; q < BA_GRAD_CHUNK / 32; ++q) {
      const int jj = lane + 32 * q, j = j0 + jj;
      BaRow o;
      o.hp = o.cp = o.cn = 0.f;
      o.n_pos = 0;
      int lab_j = -2;
      float sim[BA_GRAD_ROWS];
      if (jj < n) {
        o = rows[j];
        lab_j = labels[j];
#pragma unroll
        for (int a = 0; a < BA_GRAD_ROWS; ++a) sim[a] = S[(size_t)min(i0 + a, B - 1) * lds + j];
      }
#pragma unroll
      for (int a = 0; a < BA_GRAD_ROWS; ++a) {
        float w = 0.f;
        if (jj < n && i0 + a < B) {
          if (lab_j == my_lab[a]) {
            w = me[a].cp + o.cp;
          } else {
            if (__fsub_rn(me[a].hp, sim[a]) < alpha) w += me[a].cn;   // j is a valid negative of anchor i
            if (__fsub_rn(o.hp, sim[a]) < alpha) w += o.cn;           // i is a valid negative of anchor j
          }
        }
        s_w[warp][a][jj] = w;
      }
    }
    __syncwarp();
    // lane t owns the column pairs d = 64 c + 2 t, + 1 (D is even): one 8-byte load per pair, one packed fma
    // (fma.rn.f32x2: both halves plain IEEE fma, two per issue slot of the fp32 pipe) per anchor and pair
    for (int jj = 0; jj < n; jj += U) {
      unsigned long long v[U][KQ / 2];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + min(jj + u, n - 1);
#pragma unroll
        for (int c = 0; c < KQ / 2; ++c)
          v[u][c] = c * 64 + 2 * lane < D ? *reinterpret_cast<const unsigned long long*>(xn + (size_t)j * D + c * 64 + 2 * lane) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int a = 0; a < BA_GRAD_ROWS; ++a) {
          const unsigned long long w2 = dup2(jj + u < n ? s_w[warp][a][jj + u] : 0.f);
#pragma unroll
          for (int c = 0; c < KQ / 2; ++c) acc[a][c] = fma2(w2, v[u][c], acc[a][c]);
        }
    }
  }
#pragma unroll
  for (int a = 0; a < BA_GRAD_ROWS; ++a)
    if (i0 + a < B)
#pragma unroll
      for (int c = 0; c < KQ / 2; ++c)
        if (c * 64 + 2 * lane < D)
          *reinterpret_cast<unsigned long long*>(part + ((size_t)blockIdx.y * B + i0 + a) * D + c * 64 + 2 * lane) = acc[a][c];
}

// dx = inv * (dn - n (n . dn)), dn = the column ranges' partial sums in ascending order
template <int KQ>
__global__ void __launch_bounds__(BA_ROW_WARPS * 32) ba_grad_reduce_kernel(const float* __restrict__ part, int n_splits,
                                                                           const float* __restrict__ xn,
                                                                           const float* __restrict__ inv, int B, int D,
                                                                           float* __restrict__ demb) {
  const int i = blockIdx.x * BA_ROW_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= B) return;
  float acc[KQ], nrm[KQ], dotp = 0.f;
#pragma unroll
  for (int c = 0; c < KQ; ++c) acc[c] = 0.f;
  for (int s = 0; s < n_splits; ++s)
#pragma unroll
    for (int c = 0; c < KQ; ++c)
      if (c * 32 + lane < D) acc[c] += part[((size_t)s * B + i) * D + c * 32 + lane];
#pragma unroll
  for (int c = 0; c < KQ; ++c) {
    nrm[c] = c * 32 + lane < D ? xn[(size_t)i * D + c * 32 + lane] : 0.f;
    dotp += acc[c] * nrm[c];
  }
  for (int o = 16; o >= 1; o >>= 1) dotp += __shfl_xor_sync(0xffffffffu, dotp, o);
  const float iv = inv[i];
  const bool clamped = iv >= 0.99e6f;
#pragma unroll
  for (int c = 0; c < KQ; ++c)
    if (c * 32 + lane < D) demb[(size_t)i * D + c * 32 + lane] = clamped ? iv * acc[c] : iv * (acc[c] - nrm[c] * dotp);
}

struct BaWorkspace {
  BaRow* rows = nullptr;
  float* pos_loss = nullptr;
  size_t row_cap = 0;
  float2* vrec = nullptr;
  size_t v_cap = 0;
  float *xn = nullptr, *S = nullptr;   // large-batch path: normalised rows [B][D], similarity matrix [B][lds]
  float* part = nullptr;               // [column ranges][B][D] partial gradients
  size_t xn_cap = 0, s_cap = 0, part_cap = 0;
  int ensure_matrix(size_t n_xn, size_t n_s, size_t n_part) {
    if (n_part > part_cap) {
      retire_device_block(part);
      part = nullptr; part_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&part, n_part * sizeof(float)));
      part_cap = n_part;
    }
    if (n_xn > xn_cap) {
      retire_device_block(xn);
      xn = nullptr; xn_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&xn, n_xn * sizeof(float)));
      xn_cap = n_xn;
    }
    if (n_s > s_cap) {
      retire_device_block(S);
      S = nullptr; s_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&S, n_s * sizeof(float)));
      s_cap = n_s;
    }
    return DIF_OK;
  }
  int ensure(size_t n_rows, size_t n_v) {
    if (n_rows > row_cap) {
      retire_device_block(rows); retire_device_block(pos_loss);
      rows = nullptr; pos_loss = nullptr; row_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&rows, n_rows * sizeof(BaRow)));
      DIF_CUDA_OK(cudaMalloc((void**)&pos_loss, n_rows * sizeof(float)));
      row_cap = n_rows;
    }
    if (n_v > v_cap) {
      retire_device_block(vrec);
      vrec = nullptr; v_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&vrec, n_v * sizeof(float2)));
      v_cap = n_v;
    }
    return DIF_OK;
  }
};
static thread_local BaWorkspace g_ba;

}  // namespace dif

using namespace dif;

extern "C" {

int dif_batch_hard(const float* emb, const int32_t* labels, int B, int D, int variant, float alpha, float* loss,
                   int32_t* pos_idx, int32_t* neg_idx, float* stats, const float* dloss, float* demb, int precision,
                   void* stream) {
  DIF_REQUIRE(emb && labels && loss && stats, DIF_ERR_INVALID, "dif_batch_hard: null argument");
  DIF_REQUIRE(B >= 1 && B <= 65536 && D >= 1 && D <= 32 * BH_MAX_KD, DIF_ERR_INVALID,
              "dif_batch_hard: B %d (1..65536), D %d (1..%d)", B, D, 32 * BH_MAX_KD);
  const int soft = (variant & DIF_LOSS_SOFT_MARGIN) ? 1 : 0;
  variant &= ~DIF_LOSS_SOFT_MARGIN;
  DIF_REQUIRE(variant == DIF_LOSS_BH_COSINE || variant == DIF_LOSS_BH_EUCLIDEAN, DIF_ERR_INVALID,
              "dif_batch_hard: variant %d (batch-all is dif_batch_all)", variant);
  DIF_REQUIRE(precision == DIF_PREC_TF32X3, DIF_ERR_INVALID,
              "dif_batch_hard: only the fp32-exact path (precision 0) is implemented");
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return variant == DIF_LOSS_BH_COSINE
             ? run_batch_hard<true>(emb, labels, B, D, alpha, soft, loss, pos_idx, neg_idx, stats, dloss, demb, st)
             : run_batch_hard<false>(emb, labels, B, D, alpha, soft, loss, pos_idx, neg_idx, stats, dloss, demb, st);
}

int dif_batch_hard_host(const float* emb_host, const int32_t* labels_host, int B, int D, int variant, float alpha,
                        float* loss_host, int32_t* pos_idx_host, int32_t* neg_idx_host, float* stats_host,
                        const float* dloss_host, float* demb_host, int precision) {
  DIF_REQUIRE(emb_host && labels_host && loss_host && stats_host && B >= 1 && D >= 1, DIF_ERR_INVALID,
              "dif_batch_hard_host: invalid argument");
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t eb = al((size_t)B * D * 4), lb = al((size_t)B * 4);
  // input block: emb | labels | dloss ; output block: loss | pos | neg | stats | demb
  const size_t in_bytes = eb + 2 * lb, out_bytes = 3 * lb + 256 + eb;
  if (int rc = g_bh_stage.ensure(in_bytes + out_bytes)) return rc;
  char* h = (char*)g_bh_stage.h;
  char* d = (char*)g_bh_stage.d;
  // (a caller that works in the staging block itself - dif_batch_hard_host_buffers - needs no copies)
  if ((const char*)emb_host != h) memcpy(h, emb_host, (size_t)B * D * 4);
  if ((const char*)labels_host != h + eb) memcpy(h + eb, labels_host, (size_t)B * 4);
  if (dloss_host && (const char*)dloss_host != h + eb + lb) memcpy(h + eb + lb, dloss_host, (size_t)B * 4);
  cudaStream_t st = g_bh_stage.st;
  // The reference's own batch (18 x 4 rows of 128 floats) is one cluster launch of ~16 us: two DMA copies around it
  // would cost more than the step.  The page-locked staging block is device-visible (unified addressing), so the
  // kernel reads its 36 KB of input and writes its 37 KB of results across PCIe itself: one launch, one wait.
  static const bool staged_only = getenv("DIF_BH_HOST_STAGED") != nullptr;   // A/B switch
  const bool zero_copy = !staged_only && B <= BH_CL_MAX_B && (g_bh_force_path == 0 || g_bh_force_path == 3);
  if (zero_copy) d = h;
  else DIF_CUDA_OK(cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, st));
  char* o = d + in_bytes;
  struct HostInputScope {
    explicit HostInputScope(bool on) { g_bh_host_input = on && getenv("DIF_BH_HOST_UNSLICED") == nullptr; }
    ~HostInputScope() { g_bh_host_input = false; }
  } scope(zero_copy);
  if (int rc = dif_batch_hard((const float*)d, (const int32_t*)(d + eb), B, D, variant, alpha, (float*)o,
                              (int32_t*)(o + lb), (int32_t*)(o + 2 * lb), (float*)(o + 3 * lb),
                              dloss_host ? (const float*)(d + eb + lb) : nullptr,
                              demb_host ? (float*)(o + 3 * lb + 256) : nullptr, precision, st))
    return rc;
  const size_t back = demb_host ? out_bytes : 3 * lb + 256;
  if (!zero_copy) DIF_CUDA_OK(cudaMemcpyAsync(h + in_bytes, o, back, cudaMemcpyDeviceToHost, st));
  DIF_CUDA_OK(cudaStreamSynchronize(st));
  char* ho = h + in_bytes;
  if ((char*)loss_host != ho) memcpy(loss_host, ho, (size_t)B * 4);
  if (pos_idx_host && (char*)pos_idx_host != ho + lb) memcpy(pos_idx_host, ho + lb, (size_t)B * 4);
  if (neg_idx_host && (char*)neg_idx_host != ho + 2 * lb) memcpy(neg_idx_host, ho + 2 * lb, (size_t)B * 4);
  if ((char*)stats_host != ho + 3 * lb) memcpy(stats_host, ho + 3 * lb, 16);
  if (demb_host && (char*)demb_host != ho + 3 * lb + 256) memcpy(demb_host, ho + 3 * lb + 256, (size_t)B * D * 4);
  return DIF_OK;
}

int dif_batch_hard_host_buffers(int B, int D, void** bufs) {
  DIF_REQUIRE(bufs && B >= 1 && D >= 1 && B <= 65536 && D <= 32 * BH_MAX_KD, DIF_ERR_INVALID,
              "dif_batch_hard_host_buffers: invalid argument");
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t eb = al((size_t)B * D * 4), lb = al((size_t)B * 4);
  const size_t in_bytes = eb + 2 * lb, out_bytes = 3 * lb + 256 + eb;
  if (int rc = g_bh_stage.ensure(in_bytes + out_bytes)) return rc;
  char* h = (char*)g_bh_stage.h;
  char* ho = h + in_bytes;
  bufs[0] = h;                  // emb    [B * D] float
  bufs[1] = h + eb;             // labels [B] int32
  bufs[2] = h + eb + lb;        // dloss  [B] float
  bufs[3] = ho;                 // loss   [B] float
  bufs[4] = ho + lb;            // pos_idx [B] int32
  bufs[5] = ho + 2 * lb;        // neg_idx [B] int32
  bufs[6] = ho + 3 * lb;        // stats  [4] float
  bufs[7] = ho + 3 * lb + 256;  // demb   [B * D] float
  return DIF_OK;
}

int dif_batch_all(const float* emb, const int32_t* labels, int B, int D, float alpha, float* loss, const float* dloss,
                  float* demb, void* stream) {
  DIF_REQUIRE(emb && labels && loss, DIF_ERR_INVALID, "dif_batch_all: null argument");
  DIF_REQUIRE(B >= 1 && B <= 65536 && D >= 1 && D <= 32 * BH_MAX_KD, DIF_ERR_INVALID,
              "dif_batch_all: B %d (1..65536), D %d (1..%d)", B, D, 32 * BH_MAX_KD);
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int row_blocks = (B + BH_RB - 1) / BH_RB;
  const int sms = std::max(1, device_sm_count());
  // at least one full staged tile (32 columns) per split: finer splits only add staging and merge work
  int splits = std::max(1, std::min((2 * sms + row_blocks - 1) / row_blocks, (B + BH_CB - 1) / BH_CB));
  splits = std::min(splits, 64);
  int cols = (B + splits - 1) / splits;
  cols = (cols + BH_TJ - 1) / BH_TJ * BH_TJ;
  splits = (B + cols - 1) / cols;
  static const bool ba_tiles_only = getenv("DIF_BA_TILES") != nullptr;   // A/B switch
  if (B >= 256 && B <= 16384 && D <= kCanonMmMaxD && D % 2 == 0 && !ba_tiles_only) {   // (even D: 8-byte row pairs)
    const int lds = (B + 3) & ~3;
    if (int rc = g_ws.ensure(1, (size_t)B)) return rc;
    if (int rc = g_ba.ensure((size_t)B, 1)) return rc;
    // column ranges of the gradient: enough (32-anchor block, range) pairs for ~4 blocks per SM, whole 128-column chunks
    const int gblocks = (B + BA_ROW_WARPS * BA_GRAD_ROWS - 1) / (BA_ROW_WARPS * BA_GRAD_ROWS);
    int ks = std::max(1, std::min(16, (4 * sms + gblocks - 1) / gblocks));
    int gcols = ((B + ks - 1) / ks + BA_GRAD_CHUNK - 1) / BA_GRAD_CHUNK * BA_GRAD_CHUNK;
    ks = (B + gcols - 1) / gcols;
    if (int rc = g_ba.ensure_matrix((size_t)B * D, (size_t)B * lds, demb ? (size_t)ks * B * D : 0)) return rc;
    ba_norm_kernel<<<(B + 7) / 8, 256, 0, st>>>(emb, B, D, g_ba.xn, g_ws.aux);
    DIF_LAUNCH_OK();
    if (int rc = canon_mm_launch(g_ba.xn, B, D, BaDotEpi{}, g_ba.S, lds, st)) return rc;
    ba_row_kernel<<<(B + BA_ROW_WARPS - 1) / BA_ROW_WARPS, BA_ROW_WARPS * 32, 0, st>>>(g_ba.S, lds, labels, B, alpha, dloss, loss,
                                                                                      g_ba.rows);
    DIF_LAUNCH_OK();
    if (demb) {
      const dim3 gg(gblocks, ks), rg((B + BA_ROW_WARPS - 1) / BA_ROW_WARPS);
      const int th = BA_ROW_WARPS * 32;
      if (D <= 128) {
        ba_grad_fast_kernel<4><<<gg, th, 0, st>>>(g_ba.S, lds, g_ba.xn, labels, B, D, g_ba.rows, alpha, gcols, g_ba.part);
        ba_grad_reduce_kernel<4><<<rg, th, 0, st>>>(g_ba.part, ks, g_ba.xn, g_ws.aux, B, D, demb);
      } else if (D <= 256) {
        ba_grad_fast_kernel<8><<<gg, th, 0, st>>>(g_ba.S, lds, g_ba.xn, labels, B, D, g_ba.rows, alpha, gcols, g_ba.part);
        ba_grad_reduce_kernel<8><<<rg, th, 0, st>>>(g_ba.part, ks, g_ba.xn, g_ws.aux, B, D, demb);
      } else {
        ba_grad_fast_kernel<16><<<gg, th, 0, st>>>(g_ba.S, lds, g_ba.xn, labels, B, D, g_ba.rows, alpha, gcols, g_ba.part);
        ba_grad_reduce_kernel<16><<<rg, th, 0, st>>>(g_ba.part, ks, g_ba.xn, g_ws.aux, B, D, demb);
      }
      count_launch(1);   // (DIF_LAUNCH_OK counts the other one)
      DIF_LAUNCH_OK();
    }
    return DIF_OK;
  }
  if (int rc = g_ws.ensure((size_t)splits * B, (size_t)B)) return rc;
  if (int rc = g_ba.ensure((size_t)B, (size_t)splits * B)) return rc;
  const size_t smem = ((size_t)(BH_RB + BH_CB) * D + BH_RB + BH_CB) * sizeof(float) + BH_CB * sizeof(int) + 3 * BH_CB * sizeof(float);
  static bool configured = false;
  if (!configured) {
    DIF_CUDA_OK(cudaFuncSetAttribute(bh_mine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DIF_CUDA_OK(cudaFuncSetAttribute(ba_valid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DIF_CUDA_OK(cudaFuncSetAttribute(ba_grad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DIF_CUDA_OK(cudaFuncSetAttribute(ba_grad_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DIF_CUDA_OK(cudaFuncSetAttribute(ba_grad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  bh_mine_kernel<true><<<dim3(row_blocks, splits), BH_WARPS * 32, smem, st>>>(emb, labels, B, D, cols, g_ws.recs, g_ws.aux);
  DIF_LAUNCH_OK();
  ba_merge1_kernel<<<(B + 255) / 256, 256, 0, st>>>(g_ws.recs, splits, B, g_ba.rows, g_ba.pos_loss);
  DIF_LAUNCH_OK();
  ba_valid_kernel<<<dim3(row_blocks, splits), BH_WARPS * 32, smem, st>>>(emb, labels, B, D, cols, g_ba.rows, alpha, g_ba.vrec);
  DIF_LAUNCH_OK();
  ba_merge2_kernel<<<(B + 255) / 256, 256, 0, st>>>(g_ba.vrec, splits, B, g_ba.pos_loss, dloss, loss, g_ba.rows);
  DIF_LAUNCH_OK();
  if (demb) {
    const int kd = (D + 31) / 32;
    if (kd <= 4) ba_grad_kernel<4><<<row_blocks, BH_WARPS * 32, smem, st>>>(emb, labels, B, D, g_ba.rows, alpha, demb);
    else if (kd <= 8) ba_grad_kernel<8><<<row_blocks, BH_WARPS * 32, smem, st>>>(emb, labels, B, D, g_ba.rows, alpha, demb);
    else ba_grad_kernel<16><<<row_blocks, BH_WARPS * 32, smem, st>>>(emb, labels, B, D, g_ba.rows, alpha, demb);
    DIF_LAUNCH_OK();
  }
  return DIF_OK;
}

int dif_batch_hard_set_path(int path) {
  DIF_REQUIRE(path >= 0 && path <= 3, DIF_ERR_INVALID,
              "dif_batch_hard_set_path: 0 auto, 1 CUDA-core miner (two launches), 2 tensor-core miner, 3 one-launch cluster step");
  g_bh_force_path = path;
  return DIF_OK;
}

int dif_labels_from_onehot(const float* onehot, int B, int C, int32_t* labels, void* stream) {
  DIF_REQUIRE(onehot && labels && B >= 1 && C >= 1, DIF_ERR_INVALID, "dif_labels_from_onehot: invalid argument");
  argmax_rows_kernel<<<(B + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(onehot, B, C, labels);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

}  // extern "C"
