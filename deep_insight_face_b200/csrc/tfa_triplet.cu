// tensorflow_addons TripletHardLoss / TripletSemiHardLoss, the third-party losses the reference compiles its
// triplet models with (deep_insight_face/networks/triplet.py:196,209,211 `tfa.losses.TripletHardLoss()`,
// `tfa.losses.TripletSemiHardLoss()`; sparse int labels, training/triplet.py:72).  Arithmetic restated from the
// published tensorflow_addons/losses/{triplet,metric_learning}.py [ext]:
//   P        = pairwise_distance: sqrt(max(|a|^2 + |b|^2 - 2ab, 0)), entries <= 0 -> exactly 0, diagonal 0
//   hard     : hp_b = max over same-identity a != b of P_ba;  hn_b = min over other identities of P_bk, evaluated
//              as min((P_b - rowmax_b) * mask) + rowmax_b;  mean_b max(hp - hn + margin, 0) | log1p(exp(hp - hn))
//   semihard : for every positive pair (b, a): the closest negative farther than P_ba (same rowmax form), else the
//              farthest negative;  sum max(margin + (P_ba - sh_ba), 0) / #positive pairs
// Layout: P [B][ldp] fp32 (canonical arithmetic of dif_canon.cuh, so selections match oracle/tfa_oracle.py bit
// for bit) in a library-owned workspace; one block per anchor turns its row into a loss term and a short sorted
// list of (column, weight = dL/dP * 1/P | 2) - a dense row of Cf only past TFA_LIST_CAP non-zeros - and the backward
// pass gathers dX_r = scale * sum_j (W_rj + W_jr)(x_r - x_j) from the lists (own list, then the anchors that list r).
// Every reduction has a fixed order: results are run-to-run identical.
#include <algorithm>
#include <cmath>

#include "../../include/dif_b200.h"
#include "bh_tile.cuh"
#include "canon_mm.cuh"

namespace dif {

constexpr int TFA_THREADS = 128;
constexpr int TFA_WARPS = TFA_THREADS / 32;
constexpr int TFA_MAX_B = 8192;   // row kernel smem: 11 B per column (90 KB), u16 column ids

// ---------------------------------------------------------------- K1: P = tfa pairwise_distance(emb)
__global__ void __launch_bounds__(BH_WARPS * 32) tfa_pdist_kernel(const float* __restrict__ x, int B, int D,
                                                                  int cols_per_split, int squared,
                                                                  float* __restrict__ P, int ldp) {
  extern __shared__ float sm[];
  float* sa = sm;                       // [BH_RB][D]
  float* sb = sa + BH_RB * D;           // [BH_CB][D]
  float* aux_a = sb + BH_CB * D;        // [BH_RB] canonical sum of squares
  float* aux_b = aux_a + BH_RB;         // [BH_CB]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * BH_RB;
  const int c_begin = blockIdx.y * cols_per_split, c_end = min(B, c_begin + cols_per_split);
  const int kd = (D + 31) / 32;
  stage_rows<false>(x, B, D, row0, BH_RB, sa, aux_a);
  __syncthreads();
  const int gi = row0 + warp * BH_TI + (lane >> 2);
  const float my_sq = aux_a[warp * BH_TI + (lane >> 2)];
  for (int c0 = c_begin; c0 < c_end; c0 += BH_CB) {
    __syncthreads();
    stage_rows<false>(x, B, D, c0, BH_CB, sb, aux_b);
    __syncthreads();
    const int steps = min(BH_CB, c_end - c0);
    for (int j0 = 0; j0 < steps; j0 += BH_TJ) {
      const float dot = tile_step_dot(sa, sb, D, kd, warp, lane, j0);
      const int jl = j0 + (lane & 3);
      const int gj = c0 + jl;
      if (gj < c_end && gi < B) {
        float sq = __fsub_rn(__fadd_rn(my_sq, aux_b[jl]), __fmul_rn(2.f, dot));
        sq = fmaxf(sq, 0.f);
        const bool err = sq <= 0.f;                       // error_mask of metric_learning.pairwise_distance
        float d = squared ? sq : __fsqrt_rn(__fadd_rn(sq, err ? 1e-16f : 0.f));
        if (err || gj == gi) d = 0.f;
        P[(size_t)gi * ldp + gj] = d;
      }
    }
  }
}

// ---------------------------------------------------------------- K1 (D <= 128): the same matrix, one thread per entry
// (canon_mm.cuh: 32 fma chains folded by a bit-reversed counter tree, upper triangle + transposed store)
__global__ void __launch_bounds__(256) tfa_sqnorm_kernel(const float* __restrict__ x, int B, int D, float* __restrict__ sq) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= B) return;
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float t = x[(size_t)r * D + d];
    acc = __fmaf_rn(t, t, acc);
  }
  acc = canon_tree(acc);
  if (lane == 0) sq[r] = acc;
}

struct TfaDistEpi {
  const float* sq;   // [B] canonical sum of squares
  int squared;
  __device__ __forceinline__ float from_norms(float sq_i, float sq_j, float dot, bool diagonal) const {
    float v = __fsub_rn(__fadd_rn(sq_i, sq_j), __fmul_rn(2.f, dot));       // (a + b == b + a: symmetric bit for bit)
    v = fmaxf(v, 0.f);
    const bool err = v <= 0.f;                           // error_mask of metric_learning.pairwise_distance
    const float d = squared ? v : __fsqrt_rn(__fadd_rn(v, err ? 1e-16f : 0.f));
    return (err || diagonal) ? 0.f : d;
  }
  __device__ __forceinline__ float operator()(int gi, int gj, float dot) const {
    return from_norms(sq[gi], sq[gj], dot, gj == gi);
  }
};

// ---------------------------------------------------------------- block-wide count
__device__ __forceinline__ int block_count(int n, int* s_c) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = n;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < TFA_WARPS; ++w) t += s_c[w];
  return t;
}

struct TfaRow {
  double loss_sum;   // hard: this anchor's loss; semihard: sum over its positives
  int n_pos;         // same-identity samples other than the anchor
  int pad;
};

// ---------------------------------------------------------------- K2: one block per anchor
// srow = P_b, sflag: 0 diagonal, 1 positive, 2 negative; Cf_b (unscaled: the common factor dloss / B or
// dloss / #pairs is applied by the gradient kernel) is built in smem and written once.
// Coefficient rows are sparse (hard: the tied extremes; semi-hard: the positives and their selected negatives), so a
// row leaves the kernel as a short list of (column, w = dL/dP * dP/d(squared distance) factor) sorted by column;
// only a row with more than TFA_LIST_CAP non-zeros is written densely into Cf (count = -1).
constexpr int TFA_LIST_CAP = 32;
struct TfaLists {
  int* count;     // [B]  entries of the row's list, -1 = the row is dense in Cf
  int* cols;      // [B][TFA_LIST_CAP] ascending
  float* wts;     // [B][TFA_LIST_CAP]
};

// dL/dP -> the weight of (x_i - x_j): 1 / P for the L2 metric, 2 for squared-L2, nothing through P == 0
__device__ __forceinline__ float tfa_weight(float cf, float p, int squared) {
  return (p > 0.f && cf != 0.f) ? (squared ? 2.f * cf : cf / p) : 0.f;
}

// Block-wide: the non-zero weights of one row, produced by `weight_of(j)`, become the row's sorted list - or, past
// the cap, a dense row of Cf.
template <typename F>
__device__ __forceinline__ void tfa_emit_row(int b, int B, int ldp, const TfaLists& L, float* __restrict__ Cf, F weight_of) {
  __shared__ int s_n;
  __shared__ int s_col[TFA_LIST_CAP];
  __shared__ float s_w[TFA_LIST_CAP];
  const int t = threadIdx.x;
  __syncthreads();
  if (t == 0) s_n = 0;
  __syncthreads();
  for (int j = t; j < B; j += TFA_THREADS) {
    const float w = weight_of(j);
    if (w != 0.f) {
      const int slot = atomicAdd(&s_n, 1);
      if (slot < TFA_LIST_CAP) {
        s_col[slot] = j;
        s_w[slot] = w;
      }
    }
  }
  __syncthreads();
  const int n = s_n;
  if (n <= TFA_LIST_CAP) {
    if (t < 32) {   // rank sort by column: the lists are consumed in ascending order, whatever order the atomics gave
      const int c = t < n ? s_col[t] : 0x7fffffff;
      const float w = t < n ? s_w[t] : 0.f;
      int rank = 0;
      for (int e = 0; e < n; ++e) rank += s_col[e] < c ? 1 : 0;
      if (t < n) {
        L.cols[(size_t)b * TFA_LIST_CAP + rank] = c;
        L.wts[(size_t)b * TFA_LIST_CAP + rank] = w;
      }
      if (t == 0) L.count[b] = n;
    }
  } else {
    for (int j = t; j < B; j += TFA_THREADS) Cf[(size_t)b * ldp + j] = weight_of(j);
    if (t == 0) L.count[b] = -1;
  }
}

// Semi-hard: values first, like the hard kernel below.  Every sweep of the row is branch-free min / max / count
// arithmetic; "which column, how many ties" is a second membership sweep against the reduced value, and the rarely
// needed facts (ties of the row maximum, of the farthest negative) are computed on demand.
template <int G>
__device__ __forceinline__ void block_min_g(float (&v)[G], float (*s_gv)[G]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
    for (int g = 0; g < G; ++g) v[g] = fminf(v[g], __shfl_xor_sync(0xffffffffu, v[g], o));
  __syncthreads();   // scratch free again
  if (lane == 0)
#pragma unroll
    for (int g = 0; g < G; ++g) s_gv[warp][g] = v[g];
  __syncthreads();
#pragma unroll
  for (int g = 0; g < G; ++g) {
    v[g] = s_gv[0][g];
#pragma unroll
    for (int w = 1; w < TFA_WARPS; ++w) v[g] = fminf(v[g], s_gv[w][g]);
  }
}
// (first index, count) of G membership tests
template <int G>
__device__ __forceinline__ void block_first_count_g(int (&first)[G], int (&cnt)[G], int (*s_gi)[G], int (*s_gc)[G]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
    for (int g = 0; g < G; ++g) {
      first[g] = min(first[g], __shfl_xor_sync(0xffffffffu, first[g], o));
      cnt[g] += __shfl_xor_sync(0xffffffffu, cnt[g], o);
    }
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int g = 0; g < G; ++g) {
      s_gi[warp][g] = first[g];
      s_gc[warp][g] = cnt[g];
    }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < G; ++g) {
    first[g] = s_gi[0][g];
    cnt[g] = s_gc[0][g];
#pragma unroll
    for (int w = 1; w < TFA_WARPS; ++w) {
      first[g] = min(first[g], s_gi[w][g]);
      cnt[g] += s_gc[w][g];
    }
  }
}

__global__ void __launch_bounds__(TFA_THREADS) tfa_semihard_row_kernel(const float* __restrict__ P, int ldp,
                                                                       const int32_t* __restrict__ labels, int B, float margin,
                                                                       int squared, TfaRow* __restrict__ rows,
                                                                       float* __restrict__ Cf, TfaLists L) {
  extern __shared__ float sm[];
  float* srow = sm;                                          // [B]
  float* scf = srow + B;                                     // [B]
  unsigned char* sflag = reinterpret_cast<unsigned char*>(scf + B);   // [B]
  unsigned short* spos = reinterpret_cast<unsigned short*>(sflag + ((B + 1) & ~1));   // [n_pos] <= [B]
  constexpr int G = 4;
  __shared__ float s_gv[TFA_WARPS][G];
  __shared__ int s_gi[TFA_WARPS][G], s_gc[TFA_WARPS][G];
  __shared__ int s_c[TFA_WARPS];
  __shared__ int s_fill;
  const int b = blockIdx.x, t = threadIdx.x;
  const int my_lab = labels[b];
  if (t == 0) s_fill = 0;
  // sweep 1: the row, its flags, row maximum, farthest negative (negatives_inside), number of positives
  float red[G] = {INFINITY, INFINITY, INFINITY, INFINITY};   // [0] = -rowmax, [1] = -max over negatives
  int npos_t = 0;
  for (int j = t; j < B; j += TFA_THREADS) {
    const float v = P[(size_t)b * ldp + j];
    const int f = j == b ? 0 : (labels[j] == my_lab ? 1 : 2);
    srow[j] = v;
    scf[j] = 0.f;
    sflag[j] = (unsigned char)f;
    npos_t += f == 1;
    red[0] = fminf(red[0], -v);
    red[1] = fminf(red[1], f == 2 ? -v : INFINITY);
  }
  block_min_g<G>(red, s_gv);
  const float rmax = -red[0];
  const int n_pos = block_count(npos_t, s_c);
  const int n_neg = B - 1 - n_pos;
  const float in_v = n_neg > 0 ? -red[1] : 0.f;
  const float inside = in_v;                                 // negatives_inside (row minimum 0 as filler)
  // ordered list of this anchor's positives (ascending column)
  if (n_pos <= TFA_THREADS) {
    // a handful of positives (PK batches): append in any order, then rank-sort
    __shared__ unsigned short s_tmp[TFA_THREADS];
    for (int j = t; j < B; j += TFA_THREADS)
      if (sflag[j] == 1) s_tmp[atomicAdd(&s_fill, 1)] = (unsigned short)j;
    __syncthreads();
    if (t < n_pos) {
      const int c = s_tmp[t];
      int rank = 0;
      for (int e = 0; e < n_pos; ++e) rank += s_tmp[e] < c ? 1 : 0;
      spos[rank] = (unsigned short)c;
    }
    __syncthreads();
  } else {
    const int lane = t & 31, warp = t >> 5;
    int filled = 0;
    for (int c0 = 0; c0 < B; c0 += TFA_THREADS) {
      const int j = c0 + t;
      const bool is_pos = j < B && sflag[j] == 1;
      const unsigned m = __ballot_sync(0xffffffffu, is_pos);
      __syncthreads();
      if (lane == 0) s_c[warp] = __popc(m);
      __syncthreads();
      int base = filled, total = 0;
#pragma unroll
      for (int w = 0; w < TFA_WARPS; ++w) {
        if (w < warp) base += s_c[w];
        total += s_c[w];
      }
      if (is_pos) spos[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)j;
      filled += total;
    }
    __syncthreads();
  }
  // on demand (block-uniform): ties of the farthest negative, ties of the row maximum
  int in_i = -1, in_c = -1, rmax_c = -1;
  auto need_inside = [&]() {
    if (in_c >= 0) return;
    int first[G] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff}, cnt[G] = {0, 0, 0, 0};
    for (int k = t; k < B; k += TFA_THREADS) {
      const bool hit = sflag[k] == 2 && srow[k] == in_v;
      cnt[0] += hit;
      first[0] = hit ? min(first[0], k) : first[0];
      cnt[1] += srow[k] == rmax;
    }
    block_first_count_g<G>(first, cnt, s_gi, s_gc);
    in_i = first[0];
    in_c = cnt[0];
    rmax_c = cnt[1];
  };
  double loss_sum = 0.0;
  for (int ia0 = 0; ia0 < n_pos; ia0 += G) {
    int a_g[G];
    float pa_g[G], ov[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      a_g[g] = ia0 + g < n_pos ? spos[ia0 + g] : -1;
      pa_g[g] = a_g[g] >= 0 ? srow[a_g[g]] : INFINITY;   // nothing is farther than +inf: an empty slot stays empty
      ov[g] = INFINITY;
    }
    // sweep A: for each of the four positives, the smallest shifted distance among the negatives farther than it
    for (int k = t; k < B; k += TFA_THREADS) {
      const float v = sflag[k] == 2 ? srow[k] : -INFINITY;   // (-inf is farther than nothing)
      const float sh = __fsub_rn(v, rmax);
#pragma unroll
      for (int g = 0; g < G; ++g) ov[g] = fminf(ov[g], v > pa_g[g] ? sh : INFINITY);
    }
    block_min_g<G>(ov, s_gv);
    // sweep B: which negative (first column) and how many tie with it.  An entry can only matter if its shifted value
    // equals one of the four minima: one OR of four compares sorts out all but a handful of columns
    int oi[G], oc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      oi[g] = 0x7fffffff;
      oc[g] = 0;
    }
    for (int k = t; k < B; k += TFA_THREADS) {
      const float v = sflag[k] == 2 ? srow[k] : -INFINITY;
      const float sh = __fsub_rn(v, rmax);
      if ((sh == ov[0]) | (sh == ov[1]) | (sh == ov[2]) | (sh == ov[3])) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const bool hit = v > pa_g[g] && sh == ov[g];
          oc[g] += hit;
          oi[g] = hit ? min(oi[g], k) : oi[g];
        }
      }
    }
    block_first_count_g<G>(oi, oc, s_gi, s_gc);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (a_g[g] < 0) continue;   // block-uniform
      const int a = a_g[g];
      const float pa = pa_g[g], out_v = ov[g];
      const int out_i = oi[g], out_c = oc[g];
      const bool outside = out_c > 0;                          // mask_final
      const float sh = outside ? __fadd_rn(out_v, rmax) : inside;
      const float lm = __fadd_rn(margin, __fsub_rn(pa, sh));
      loss_sum += (double)fmaxf(lm, 0.f);
      if (Cf && lm >= 0.f) {
        if (!outside || out_v == 0.f) need_inside();
        // the usual case touches two entries: +1 on the positive, -1 on the one selected negative
        if (outside && out_c == 1 && out_v != 0.f) {
          if (t == 0) {
            scf[a] += 1.f;
            scf[out_i] -= 1.f;
          }
        } else if (!outside && in_c == 1) {
          if (t == 0) {
            scf[a] += 1.f;
            scf[in_i] -= 1.f;
          }
        } else if (outside) {
          int n_tied = out_c;
          if (out_v == 0.f) {   // (rare: the selected negative is the row maximum) every unmasked entry ties at 0
            int n_mask = 0;
            for (int k = t; k < B; k += TFA_THREADS) n_mask += (sflag[k] == 2 && srow[k] > pa) ? 1 : 0;
            n_tied += B - block_count(n_mask, s_c);
          }
          const float c_neg = 1.f / (float)n_tied;
          const float c_max = out_v == 0.f ? (1.f - (float)out_c / (float)n_tied) / (float)rmax_c : 0.f;
          __syncthreads();   // thread 0's two-entry updates above are visible
          for (int k = t; k < B; k += TFA_THREADS) {
            float cf = scf[k];
            if (k == a) cf += 1.f;
            if (sflag[k] == 2 && srow[k] > pa && __fsub_rn(srow[k], rmax) == out_v) cf -= c_neg;
            if (out_v == 0.f && srow[k] == rmax) cf -= c_max;
            scf[k] = cf;
          }
          __syncthreads();
        } else {
          const float c_neg = in_c > 0 ? 1.f / (float)in_c : 0.f;
          __syncthreads();
          for (int k = t; k < B; k += TFA_THREADS) {
            float cf = scf[k];
            if (k == a) cf += 1.f;
            if (sflag[k] == 2 && srow[k] == in_v) cf -= c_neg;
            scf[k] = cf;
          }
          __syncthreads();
        }
      }
    }
  }
  if (t == 0) {
    rows[b].loss_sum = loss_sum;
    rows[b].n_pos = n_pos;
  }
  if (Cf) tfa_emit_row(b, B, ldp, L, Cf, [&](int j) { return tfa_weight(scf[j], srow[j], squared); });
}

// ---------------------------------------------------------------- K2 (hard): two light passes per anchor
// The generic kernel above carries (value, first index, tie count) triples through every reduction.  The hard loss
// only needs VALUES first - rowmax, the farthest positive, the closest negative (the shifted minimum of
// _masked_minimum is monotone in P, so min((P - rowmax) * mask) = fl(min P - rowmax)) - and then one membership pass:
// which columns tie with those values (first index, count), appended to the row's weight list as they are found.
struct TfaHardRed {
  float rmax, hp, nmin;
  int n_pos;
};
__global__ void __launch_bounds__(TFA_THREADS) tfa_hard_row_kernel(const float* __restrict__ P, int ldp,
                                                                   const int32_t* __restrict__ labels, int B, float margin,
                                                                   int soft, int squared, TfaRow* __restrict__ rows,
                                                                   int32_t* __restrict__ pos_idx, int32_t* __restrict__ neg_idx,
                                                                   float* __restrict__ Cf, TfaLists L) {
  extern __shared__ __align__(16) float sm[];   // (16-byte row stores below)
  float* srow = sm;                                                    // [B]
  unsigned char* sflag = reinterpret_cast<unsigned char*>(srow + B);   // [B] 0 diagonal, 1 positive, 2 negative
  __shared__ TfaHardRed s_red[TFA_WARPS];
  __shared__ int s_cnt[3], s_first[2], s_n;
  __shared__ int s_col[TFA_LIST_CAP];
  __shared__ unsigned char s_bits[TFA_LIST_CAP];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int my_lab = labels[b];
  if (t < 3) s_cnt[t] = 0;
  if (t < 2) s_first[t] = 0x7fffffff;
  if (t == 0) s_n = 0;
  // pass 1: values only
  TfaHardRed me{-INFINITY, -INFINITY, INFINITY, 0};
  auto visit1 = [&](int j, float v, int lab_j) {
    const int f = j == b ? 0 : (lab_j == my_lab ? 1 : 2);
    me.rmax = fmaxf(me.rmax, v);
    me.hp = fmaxf(me.hp, f == 1 ? v : -INFINITY);
    me.nmin = fminf(me.nmin, f == 2 ? v : INFINITY);
    me.n_pos += f == 1;
    return f;
  };
  // four columns per trip where the alignment allows (the kernel is instruction bound: 46 per entry one at a time)
  const int B4 = ((reinterpret_cast<uintptr_t>(labels) & 15u) == 0) ? (B & ~3) : 0;
  {
    const float4* P4 = reinterpret_cast<const float4*>(P + (size_t)b * ldp);   // ldp % 4 == 0, P from cudaMalloc
    const int4* L4 = reinterpret_cast<const int4*>(labels);
    for (int q = t; q < (B4 >> 2); q += TFA_THREADS) {
      const float4 v = P4[q];
      const int4 l = L4[q];
      const int j = q << 2;
      const int f0 = visit1(j, v.x, l.x), f1 = visit1(j + 1, v.y, l.y), f2 = visit1(j + 2, v.z, l.z), f3 = visit1(j + 3, v.w, l.w);
      *reinterpret_cast<float4*>(srow + j) = v;
      *reinterpret_cast<uint32_t*>(sflag + j) = (uint32_t)f0 | ((uint32_t)f1 << 8) | ((uint32_t)f2 << 16) | ((uint32_t)f3 << 24);
    }
  }
  for (int j = B4 + t; j < B; j += TFA_THREADS) {
    const float v = P[(size_t)b * ldp + j];
    srow[j] = v;
    sflag[j] = (unsigned char)visit1(j, v, labels[j]);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    me.rmax = fmaxf(me.rmax, __shfl_xor_sync(0xffffffffu, me.rmax, o));
    me.hp = fmaxf(me.hp, __shfl_xor_sync(0xffffffffu, me.hp, o));
    me.nmin = fminf(me.nmin, __shfl_xor_sync(0xffffffffu, me.nmin, o));
    me.n_pos += __shfl_xor_sync(0xffffffffu, me.n_pos, o);
  }
  if (lane == 0) s_red[warp] = me;
  __syncthreads();
  me = s_red[0];
#pragma unroll
  for (int w = 1; w < TFA_WARPS; ++w) {
    me.rmax = fmaxf(me.rmax, s_red[w].rmax);
    me.hp = fmaxf(me.hp, s_red[w].hp);
    me.nmin = fminf(me.nmin, s_red[w].nmin);
    me.n_pos += s_red[w].n_pos;
  }
  const float rmax = me.rmax, hp_v = me.hp;
  const int n_pos = me.n_pos, n_neg = B - 1 - n_pos;
  const bool has_pos = n_pos > 0, has_neg = n_neg > 0;
  const float vmin = has_neg ? __fsub_rn(me.nmin, rmax) : 0.f;   // unmasked entries contribute (P - rowmax) * 0
  const bool through_max = vmin == 0.f;                         // reduce_min then ties with every unmasked entry
  // pass 2: membership.  bit 0: a positive at hp; bit 1: a negative whose shifted value is vmin; bit 2: a row maximum
  // (only when the gradient reaches the rowmax term)
  int c_p = 0, c_n = 0, c_m = 0, f_p = 0x7fffffff, f_n = 0x7fffffff;
  auto visit2 = [&](int j, float v, int f) {
    const bool pt = f == 1 && v == hp_v;
    const bool nt = f == 2 && __fsub_rn(v, rmax) == vmin;
    const bool mt = through_max && v == rmax;
    if (pt | nt | mt) {
      c_p += pt;
      c_n += nt;
      c_m += mt;
      if (pt) f_p = min(f_p, j);
      if (nt) f_n = min(f_n, j);
      if (Cf) {
        const int slot = atomicAdd(&s_n, 1);
        if (slot < TFA_LIST_CAP) {
          s_col[slot] = j;
          s_bits[slot] = (unsigned char)((pt ? 1 : 0) | (nt ? 2 : 0) | (mt ? 4 : 0));
        }
      }
    }
  };
  for (int q = t; q < (B4 >> 2); q += TFA_THREADS) {
    const int j = q << 2;
    const float4 v = *reinterpret_cast<const float4*>(srow + j);
    const uint32_t f = *reinterpret_cast<const uint32_t*>(sflag + j);
    visit2(j, v.x, (int)(f & 255u));
    visit2(j + 1, v.y, (int)((f >> 8) & 255u));
    visit2(j + 2, v.z, (int)((f >> 16) & 255u));
    visit2(j + 3, v.w, (int)(f >> 24));
  }
  for (int j = B4 + t; j < B; j += TFA_THREADS) visit2(j, srow[j], sflag[j]);
  if (c_p) atomicAdd(&s_cnt[0], c_p);
  if (c_n) atomicAdd(&s_cnt[1], c_n);
  if (c_m) atomicAdd(&s_cnt[2], c_m);
  if (f_p != 0x7fffffff) atomicMin(&s_first[0], f_p);
  if (f_n != 0x7fffffff) atomicMin(&s_first[1], f_n);
  __syncthreads();
  const int hp_c = s_cnt[0], mn_c = s_cnt[1], rmax_c = s_cnt[2];
  const float hp = has_pos ? hp_v : 0.f;                     // _masked_maximum: row minimum (the 0 diagonal) as filler
  const float hn = __fadd_rn(vmin, rmax);
  const float xd = __fsub_rn(hp, hn);
  float loss, g;
  if (soft) {
    loss = log1pf(expf(xd));
    g = 1.f / (1.f + expf(-xd));
  } else {
    const float basic = __fadd_rn(xd, margin);
    loss = fmaxf(basic, 0.f);
    g = basic >= 0.f ? 1.f : 0.f;
  }
  if (t == 0) {
    rows[b].loss_sum = (double)loss;
    rows[b].n_pos = n_pos;
    if (pos_idx) pos_idx[b] = has_pos ? s_first[0] : -1;
    if (neg_idx) neg_idx[b] = has_neg ? s_first[1] : -1;
  }
  if (!Cf) return;
  // reduce_min tie set: masked entries at vmin, plus every unmasked entry when vmin == 0
  const int m_tied = mn_c;
  const int n_tied = through_max ? m_tied + (B - n_neg) : m_tied;
  const float c_pos = hp_c > 0 ? g / (float)hp_c : 0.f;
  const float c_neg = m_tied > 0 ? g / (float)n_tied : 0.f;
  const float c_max = through_max ? g * (1.f - (float)m_tied / (float)n_tied) / (float)rmax_c : 0.f;
  auto weight_of = [&](int j, int bits) {
    float cf = 0.f;
    if (bits & 1) cf += c_pos;
    if (bits & 2) cf -= c_neg;
    if (bits & 4) cf -= c_max;                               // gradient through the rowmax term of _masked_minimum
    return tfa_weight(cf, srow[j], squared);
  };
  const int n = s_n;
  if (n <= TFA_LIST_CAP) {
    if (t < 32) {   // rank sort by column
      const int c = t < n ? s_col[t] : 0x7fffffff;
      int rank = 0;
      for (int e = 0; e < n; ++e) rank += s_col[e] < c ? 1 : 0;
      if (t < n) {
        L.cols[(size_t)b * TFA_LIST_CAP + rank] = c;
        L.wts[(size_t)b * TFA_LIST_CAP + rank] = weight_of(c, s_bits[t]);
      }
      if (t == 0) L.count[b] = n;
    }
  } else {
    for (int j = t; j < B; j += TFA_THREADS) {
      const float v = srow[j];
      const int f = sflag[j];
      const int bits = ((f == 1 && v == hp_v) ? 1 : 0) | ((f == 2 && __fsub_rn(v, rmax) == vmin) ? 2 : 0) |
                       ((through_max && v == rmax) ? 4 : 0);
      Cf[(size_t)b * ldp + j] = weight_of(j, bits);
    }
    if (t == 0) L.count[b] = -1;
  }
}

// ---------------------------------------------------------------- K3: scalar loss and the backward scale
// All threads of a block of >= 256 threads call this; the first 256 fold the per-anchor terms in a fixed tree, so the
// stand-alone kernel and the gradient kernel that does it for itself (small batches) produce the same bits.
__device__ __forceinline__ void tfa_reduce_rows(const TfaRow* __restrict__ rows, int B, int kind, float dloss, float& loss,
                                                float& scale) {
  __shared__ double s_sum[256];
  __shared__ long long s_cnt[256];
  const int t = threadIdx.x;
  if (t < 256) {
    double sum = 0.0;
    long long cnt = 0;
    for (int b = t; b < B; b += 256) {
      sum += rows[b].loss_sum;
      cnt += rows[b].n_pos;
    }
    s_sum[t] = sum;
    s_cnt[t] = cnt;
  }
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (t < o) {
      s_sum[t] += s_sum[t + o];
      s_cnt[t] += s_cnt[t + o];
    }
    __syncthreads();
  }
  const double denom = kind == DIF_TFA_HARD ? (double)B : (double)s_cnt[0];   // 0 positive pairs: 0 / 0 = NaN as in tfa
  loss = (float)(s_sum[0] / denom);
  scale = (float)((double)dloss / denom);
}

__global__ void __launch_bounds__(256) tfa_finalize_kernel(const TfaRow* __restrict__ rows, int B, int kind, float dloss,
                                                           float* __restrict__ loss, float* __restrict__ scale) {
  float l, sc;
  tfa_reduce_rows(rows, B, kind, dloss, l, sc);
  if (threadIdx.x == 0) {
    loss[0] = l;
    scale[0] = sc;
  }
}

// ---------------------------------------------------------------- K4: dX_r = scale * sum_j (W_rj + W_jr) (x_r - x_j)
// W = the rows' weight lists.  The transposed half needs "which anchors list r": every block scans the columns of all
// lists once (L2 resident) and sets bit a of row r's bitmap in shared memory when anchor a lists one of the block's
// rows (dense anchors go to a bitmap of their own); each warp then gathers its row - own list in ascending column
// order, then the listing anchors in ascending order - so the sums are reproducible without atomics on the result.
constexpr int TFA_GRAD_WARPS = 16;
constexpr int TFA_INV_CAP = 32;   // listing anchors kept per row in shared memory; a busier row walks its bitmap
__global__ void __launch_bounds__(TFA_GRAD_WARPS * 32) tfa_grad_sparse_kernel(TfaLists L, const float* __restrict__ Cf, int ldp,
                                                                              const float* __restrict__ x, int B, int D,
                                                                              const float* __restrict__ scale,
                                                                              float* __restrict__ dX,
                                                                              const TfaRow* __restrict__ fin_rows, int fin_kind,
                                                                              float fin_dloss, float* __restrict__ fin_loss) {
  extern __shared__ unsigned s_map[];   // [TFA_GRAD_WARPS + 1][W]
  // small batches: no separate finalize launch - every block folds the per-anchor terms itself (block 0 writes the loss)
  float fused_scale = 0.f;
  if (fin_rows) {
    float l;
    tfa_reduce_rows(fin_rows, B, fin_kind, fin_dloss, l, fused_scale);
    if (blockIdx.x == 0 && threadIdx.x == 0) fin_loss[0] = l;
  }
  __shared__ int s_cnt[TFA_GRAD_WARPS];
  __shared__ int s_anchor[TFA_GRAD_WARPS][TFA_INV_CAP];
  __shared__ float s_weight[TFA_GRAD_WARPS][TFA_INV_CAP];
  const int W = (B + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = blockIdx.x * TFA_GRAD_WARPS;
  for (int i = threadIdx.x; i < (TFA_GRAD_WARPS + 1) * W; i += blockDim.x) s_map[i] = 0u;
  if (threadIdx.x < TFA_GRAD_WARPS) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  // scan: counts of 4 anchors, then the first four columns of each, are in flight together (most lists are that short)
  auto note = [&](int a, int e, int col) {
    const unsigned u = (unsigned)(col - r0);
    if (u < (unsigned)TFA_GRAD_WARPS) {
      atomicOr(&s_map[u * W + (a >> 5)], 1u << (a & 31));
      const int slot = atomicAdd(&s_cnt[u], 1);
      if (slot < TFA_INV_CAP) {
        s_anchor[u][slot] = a;
        s_weight[u][slot] = L.wts[(size_t)a * TFA_LIST_CAP + e];
      }
    }
  };
  for (int a0 = threadIdx.x; a0 < B; a0 += 4 * TFA_GRAD_WARPS * 32) {
    int n[4];
    int4 c[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int a = a0 + q * TFA_GRAD_WARPS * 32;
      n[q] = a < B ? L.count[a] : 0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int a = a0 + q * TFA_GRAD_WARPS * 32;
      c[q] = n[q] > 0 ? *reinterpret_cast<const int4*>(L.cols + (size_t)a * TFA_LIST_CAP) : make_int4(-1, -1, -1, -1);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int a = a0 + q * TFA_GRAD_WARPS * 32;
      if (n[q] < 0) atomicOr(&s_map[TFA_GRAD_WARPS * W + (a >> 5)], 1u << (a & 31));
      if (n[q] > 0) note(a, 0, c[q].x);
      if (n[q] > 1) note(a, 1, c[q].y);
      if (n[q] > 2) note(a, 2, c[q].z);
      if (n[q] > 3) note(a, 3, c[q].w);
      for (int e = 4; e < n[q]; ++e) note(a, e, L.cols[(size_t)a * TFA_LIST_CAP + e]);
    }
  }
  __syncthreads();
  const int r = r0 + warp;
  if (r >= B) return;
  float xr[BH_MAX_KD], acc[BH_MAX_KD];
#pragma unroll
  for (int c = 0; c < BH_MAX_KD; ++c) {
    const int d = c * 32 + lane;
    xr[c] = d < D ? x[(size_t)r * D + d] : 0.f;
    acc[c] = 0.f;
  }
  auto axpy = [&](float w, int j) {
    const float* xj = x + (size_t)j * D;
#pragma unroll
    for (int c = 0; c < BH_MAX_KD; ++c) {
      const int d = c * 32 + lane;
      if (d < D) acc[c] = __fmaf_rn(w, xr[c] - xj[d], acc[c]);
    }
  };
  // a list held one entry per lane: rows are fetched four at a time, added in list order
  auto gather = [&](int n, int my_j, float my_w) {
    if (D <= 128) {
      int e0 = 0;
      // long lists (hub rows): eight rows in flight per trip
      for (; e0 + 8 <= n; e0 += 8) {
        float v[8][4], w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = __shfl_sync(0xffffffffu, my_j, e0 + u);
          w[u] = __shfl_sync(0xffffffffu, my_w, e0 + u);
#pragma unroll
          for (int c = 0; c < 4; ++c) v[u][c] = c * 32 + lane < D ? x[(size_t)j * D + c * 32 + lane] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[c] = __fmaf_rn(w[u], xr[c] - v[u][c], acc[c]);
      }
      for (; e0 < n; e0 += 4) {
        float v[4][4], w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int src = min(e0 + u, n - 1);
          const int j = __shfl_sync(0xffffffffu, my_j, src);
          w[u] = e0 + u < n ? __shfl_sync(0xffffffffu, my_w, src) : 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) v[u][c] = c * 32 + lane < D ? x[(size_t)j * D + c * 32 + lane] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (e0 + u < n) acc[c] = __fmaf_rn(w[u], xr[c] - v[u][c], acc[c]);
      }
    } else {
      for (int e = 0; e < n; ++e) axpy(__shfl_sync(0xffffffffu, my_w, e), __shfl_sync(0xffffffffu, my_j, e));
    }
  };
  // own row
  const int n_own = L.count[r];
  if (n_own >= 0) {
    gather(n_own, lane < n_own ? L.cols[(size_t)r * TFA_LIST_CAP + lane] : 0,
           lane < n_own ? L.wts[(size_t)r * TFA_LIST_CAP + lane] : 0.f);
  } else {
    for (int j0 = 0; j0 < B; j0 += 32) {
      const float w = j0 + lane < B ? Cf[(size_t)r * ldp + j0 + lane] : 0.f;
      unsigned nz = __ballot_sync(0xffffffffu, w != 0.f);
      while (nz) {
        const int src = __ffs((int)nz) - 1;
        nz &= nz - 1;
        axpy(__shfl_sync(0xffffffffu, w, src), j0 + src);
      }
    }
  }
  // anchors that list r
  const unsigned* dm = s_map + TFA_GRAD_WARPS * W;
  bool any_dense = false;
  for (int wi = lane; wi < W; wi += 32) any_dense |= dm[wi] != 0u;
  const int n_inv = s_cnt[warp];
  if (n_inv <= TFA_INV_CAP && !__any_sync(0xffffffffu, any_dense)) {
    // the usual case: the scan already holds (anchor, weight) of every listing anchor; rank them by anchor
    const int a = lane < n_inv ? s_anchor[warp][lane] : 0x7fffffff;
    const float w = lane < n_inv ? s_weight[warp][lane] : 0.f;
    int rank = 0;
    for (int e = 0; e < n_inv; ++e) rank += s_anchor[warp][e] < a ? 1 : 0;
    __syncwarp();
    if (lane < n_inv) {
      s_anchor[warp][rank] = a;
      s_weight[warp][rank] = w;
    }
    __syncwarp();
    gather(n_inv, lane < n_inv ? s_anchor[warp][lane] : 0, lane < n_inv ? s_weight[warp][lane] : 0.f);
  } else {
    // a row many anchors list (hub rows: the hardest negative of dozens of anchors), or dense anchors around: walk the
    // bitmap 32 words at a time, hand the set bits to the lanes in ascending order, 32 anchors per round; every lane
    // looks its own anchor's weight up (the lookups run side by side), then the rows are gathered as above
    const unsigned* bm = s_map + warp * W;
    for (int w0 = 0; w0 < W; w0 += 32) {
      const int wi = w0 + lane;
      const unsigned db = wi < W ? dm[wi] : 0u;
      const unsigned word = (wi < W ? bm[wi] : 0u) | db;
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
          if (k >= base && k < base + 32) s_anchor[warp][k - base] = wi * 32 + bit;
        }
        __syncwarp();
        const int n = min(32, total - base);
        int a = 0;
        float w = 0.f;
        if (lane < n) {
          a = s_anchor[warp][lane];
          const int na = L.count[a];
          if (na < 0) {
            w = Cf[(size_t)a * ldp + r];
          } else {
            for (int e = 0; e < na; ++e)
              if (L.cols[(size_t)a * TFA_LIST_CAP + e] == r) w = L.wts[(size_t)a * TFA_LIST_CAP + e];
          }
        }
        gather(n, a, w);
      }
    }
  }
  const float sc = fin_rows ? fused_scale : scale[0];
#pragma unroll
  for (int c = 0; c < BH_MAX_KD; ++c) {
    const int d = c * 32 + lane;
    if (d < D) dX[(size_t)r * D + d] = sc * acc[c];
  }
}

struct TfaWorkspace {
  float* P = nullptr;
  float* Cf = nullptr;
  TfaRow* rows = nullptr;
  float* scale = nullptr;
  float* sq = nullptr;   // [rows] canonical sum of squares (fast pairwise kernel)
  TfaLists lists{nullptr, nullptr, nullptr};
  size_t mat_cap = 0, row_cap = 0;
  // (outgrown blocks are parked, not freed: a CUDA graph captured at the smaller size - TfaTripletStep - may still use them)
  int ensure(size_t mat, size_t n_rows, bool want_cf) {
    if (mat > mat_cap) {
      retire_device_block(P);
      retire_device_block(Cf);
      P = Cf = nullptr;
      mat_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&P, mat * sizeof(float)));
      mat_cap = mat;
    }
    if (want_cf && !Cf) DIF_CUDA_OK(cudaMalloc((void**)&Cf, mat_cap * sizeof(float)));
    if (n_rows > row_cap) {
      retire_device_block(rows);
      retire_device_block(sq);
      retire_device_block(lists.count);
      retire_device_block(lists.cols);
      retire_device_block(lists.wts);
      rows = nullptr;
      sq = nullptr;
      lists = TfaLists{nullptr, nullptr, nullptr};
      row_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&rows, n_rows * sizeof(TfaRow)));
      DIF_CUDA_OK(cudaMalloc((void**)&sq, n_rows * sizeof(float)));
      DIF_CUDA_OK(cudaMalloc((void**)&lists.count, n_rows * sizeof(int)));
      DIF_CUDA_OK(cudaMalloc((void**)&lists.cols, n_rows * TFA_LIST_CAP * sizeof(int)));
      DIF_CUDA_OK(cudaMalloc((void**)&lists.wts, n_rows * TFA_LIST_CAP * sizeof(float)));
      row_cap = n_rows;
    }
    if (!scale) DIF_CUDA_OK(cudaMalloc((void**)&scale, 16));
    return DIF_OK;
  }
};
static thread_local TfaWorkspace g_tfa;

}  // namespace dif

using namespace dif;

extern "C" int dif_tfa_triplet(const float* emb, const int32_t* labels, int B, int D, int kind, float margin, float* loss,
                               int32_t* pos_idx, int32_t* neg_idx, float dloss, float* demb, void* stream) {
  DIF_REQUIRE(emb && labels && loss, DIF_ERR_INVALID, "dif_tfa_triplet: null argument");
  DIF_REQUIRE(B >= 1 && B <= TFA_MAX_B && D >= 1 && D <= 32 * BH_MAX_KD, DIF_ERR_INVALID,
              "dif_tfa_triplet: B %d (1..%d), D %d (1..%d)", B, TFA_MAX_B, D, 32 * BH_MAX_KD);
  const int soft = (kind & DIF_TFA_SOFT) ? 1 : 0, squared = (kind & DIF_TFA_SQUARED) ? 1 : 0;
  const int base = kind & ~(DIF_TFA_SOFT | DIF_TFA_SQUARED);
  DIF_REQUIRE(base == DIF_TFA_HARD || base == DIF_TFA_SEMIHARD, DIF_ERR_INVALID, "dif_tfa_triplet: kind %d", kind);
  DIF_REQUIRE(!(soft && base == DIF_TFA_SEMIHARD), DIF_ERR_INVALID, "dif_tfa_triplet: the semi-hard loss has no soft margin");
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ldp = (B + 3) & ~3;
  if (int rc = g_tfa.ensure((size_t)B * ldp, (size_t)B, demb != nullptr)) return rc;
  static bool configured = false;
  if (!configured) {
    DIF_CUDA_OK(cudaFuncSetAttribute(tfa_pdist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DIF_CUDA_OK(cudaFuncSetAttribute(tfa_hard_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    DIF_CUDA_OK(cudaFuncSetAttribute(tfa_semihard_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  // K1
  const int sms = std::max(1, device_sm_count());
  static const bool slow_pdist = getenv("DIF_TFA_SLOW_PDIST") != nullptr;   // A/B switches
  static const bool no_fuse = getenv("DIF_TFA_UNFUSED") != nullptr;
  // small batches (the reference's own: 18 x 4) are launch chains: the gradient blocks fold the loss terms themselves, one
  // launch less.  (Letting every row block compute its own row of the matrix as well - two launches - was measured and
  // dropped: 29.5 us graphed at B = 72 against 24.6, each block repeats 2 B warp-dots.)
  const bool small = B <= 128 && !no_fuse;
  const float* Pm = g_tfa.P;
  if (D <= kCanonMmMaxD && B >= 256 && !slow_pdist) {
    tfa_sqnorm_kernel<<<(B + 7) / 8, 256, 0, st>>>(emb, B, D, g_tfa.sq);
    DIF_LAUNCH_OK();
    if (int rc = canon_mm_launch(emb, B, D, TfaDistEpi{g_tfa.sq, squared}, g_tfa.P, ldp, st)) return rc;
  } else {
    const int row_blocks = (B + BH_RB - 1) / BH_RB;
    int splits = std::max(1, std::min((2 * sms + row_blocks - 1) / row_blocks, (B + BH_CB - 1) / BH_CB));
    int cols = (B + splits - 1) / splits;
    cols = (cols + BH_CB - 1) / BH_CB * BH_CB;
    splits = (B + cols - 1) / cols;
    const size_t smem1 = ((size_t)(BH_RB + BH_CB) * D + BH_RB + BH_CB) * sizeof(float);
    tfa_pdist_kernel<<<dim3(row_blocks, splits), BH_WARPS * 32, smem1, st>>>(emb, B, D, cols, squared, g_tfa.P, ldp);
    DIF_LAUNCH_OK();
  }
  // K2
  const size_t smem2 = (size_t)B * 11 + 32;   // row, coefficient row (fp32), flags (u8), positives list (u16)
  float* cf = demb ? g_tfa.Cf : nullptr;
  if (base == DIF_TFA_HARD)
    tfa_hard_row_kernel<<<B, TFA_THREADS, (size_t)B * 5 + 16, st>>>(Pm, ldp, labels, B, margin, soft, squared, g_tfa.rows,
                                                                   pos_idx, neg_idx, cf, g_tfa.lists);
  else
    tfa_semihard_row_kernel<<<B, TFA_THREADS, smem2, st>>>(Pm, ldp, labels, B, margin, squared, g_tfa.rows, cf, g_tfa.lists);
  DIF_LAUNCH_OK();
  // K3 (folded into K4 for small batches with a backward pass)
  const bool fuse_fin = small && demb != nullptr;
  if (!fuse_fin) {
    tfa_finalize_kernel<<<1, 256, 0, st>>>(g_tfa.rows, B, base, dloss, loss, g_tfa.scale);
    DIF_LAUNCH_OK();
  }
  if (!demb) return DIF_OK;
  // K4
  const size_t map_smem = (size_t)(TFA_GRAD_WARPS + 1) * ((B + 31) / 32) * sizeof(unsigned);
  tfa_grad_sparse_kernel<<<(B + TFA_GRAD_WARPS - 1) / TFA_GRAD_WARPS, TFA_GRAD_WARPS * 32, map_smem, st>>>(
      g_tfa.lists, g_tfa.Cf, ldp, emb, B, D, g_tfa.scale, demb, fuse_fin ? g_tfa.rows : nullptr, base, dloss, loss);
  DIF_LAUNCH_OK();
  return DIF_OK;
}
