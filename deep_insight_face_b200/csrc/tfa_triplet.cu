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
// for bit) and the coefficient matrix Cf = dL/dP [B][ldp], both in a library-owned workspace; the backward pass
// folds Cf + Cf^T and 1/P into P in place and scatters w_ij (x_i - x_j) from the few non-zero entries of each row.
// Every reduction has a fixed order: results are run-to-run identical.
#include <algorithm>
#include <cmath>

#include "../../include/dif_b200.h"
#include "bh_tile.cuh"

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

// ---------------------------------------------------------------- block-wide (value, first index, tie count)
template <bool MIN>
__device__ __forceinline__ void block_extreme(float& v, int& i, int& c, float* s_v, int* s_i, int* s_c) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o), oc = __shfl_xor_sync(0xffffffffu, c, o);
    merge<MIN>(ov, oi, oc, v, i, c);
  }
  __syncthreads();   // scratch free again
  if (lane == 0) {
    s_v[warp] = v;
    s_i[warp] = i;
    s_c[warp] = c;
  }
  __syncthreads();
  v = s_v[0];
  i = s_i[0];
  c = s_c[0];
#pragma unroll
  for (int w = 1; w < TFA_WARPS; ++w) merge<MIN>(s_v[w], s_i[w], s_c[w], v, i, c);
}

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
template <int KIND>
__global__ void __launch_bounds__(TFA_THREADS) tfa_row_kernel(const float* __restrict__ P, int ldp,
                                                              const int32_t* __restrict__ labels, int B, float margin,
                                                              int soft, TfaRow* __restrict__ rows,
                                                              int32_t* __restrict__ pos_idx, int32_t* __restrict__ neg_idx,
                                                              float* __restrict__ Cf) {
  extern __shared__ float sm[];
  float* srow = sm;                                          // [B]
  float* scf = srow + B;                                     // [B]
  unsigned char* sflag = reinterpret_cast<unsigned char*>(scf + B);   // [B]
  __shared__ float s_v[TFA_WARPS];
  __shared__ int s_i[TFA_WARPS], s_c[TFA_WARPS];
  const int b = blockIdx.x, t = threadIdx.x;
  const int my_lab = labels[b];
  float rmax = -INFINITY;
  int rmax_i = -1, rmax_c = 0, npos_t = 0;
  for (int j = t; j < B; j += TFA_THREADS) {
    const float v = P[(size_t)b * ldp + j];
    srow[j] = v;
    scf[j] = 0.f;
    const int f = j == b ? 0 : (labels[j] == my_lab ? 1 : 2);
    sflag[j] = (unsigned char)f;
    npos_t += f == 1;
    fold<false>(v, j, rmax, rmax_i, rmax_c);
  }
  block_extreme<false>(rmax, rmax_i, rmax_c, s_v, s_i, s_c);
  const int n_pos = block_count(npos_t, s_c);
  const int n_neg = B - 1 - n_pos;

  if (KIND == DIF_TFA_HARD) {
    float hp_v = -INFINITY, mn_v = INFINITY;
    int hp_i = -1, hp_c = 0, mn_i = -1, mn_c = 0;
    for (int j = t; j < B; j += TFA_THREADS) {
      if (sflag[j] == 1) fold<false>(srow[j], j, hp_v, hp_i, hp_c);
      if (sflag[j] == 2) fold<true>(__fsub_rn(srow[j], rmax), j, mn_v, mn_i, mn_c);
    }
    block_extreme<false>(hp_v, hp_i, hp_c, s_v, s_i, s_c);
    block_extreme<true>(mn_v, mn_i, mn_c, s_v, s_i, s_c);
    const float hp = hp_c > 0 ? hp_v : 0.f;                  // _masked_maximum: row minimum (the 0 diagonal) as filler
    const float vmin = mn_c > 0 ? mn_v : 0.f;                // unmasked entries contribute (P - rowmax) * 0
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
      if (pos_idx) pos_idx[b] = hp_c > 0 ? hp_i : -1;
      if (neg_idx) neg_idx[b] = mn_c > 0 ? mn_i : -1;
    }
    if (Cf) {
      // reduce_min tie set: masked entries at vmin, plus every unmasked entry when vmin == 0
      const int m_tied = (mn_c > 0 && mn_v == vmin) ? mn_c : 0;
      const int n_tied = vmin == 0.f ? m_tied + (B - n_neg) : m_tied;
      const float c_pos = hp_c > 0 ? g / (float)hp_c : 0.f;
      const float c_neg = m_tied > 0 ? g / (float)n_tied : 0.f;
      const float c_max = vmin == 0.f ? g * (1.f - (float)m_tied / (float)n_tied) / (float)rmax_c : 0.f;
      for (int j = t; j < B; j += TFA_THREADS) {
        float cf = 0.f;
        if (sflag[j] == 1 && srow[j] == hp_v) cf += c_pos;
        if (sflag[j] == 2 && __fsub_rn(srow[j], rmax) == vmin) cf -= c_neg;
        if (srow[j] == rmax) cf -= c_max;                   // gradient through the rowmax term of _masked_minimum
        Cf[(size_t)b * ldp + j] = cf;
      }
    }
    return;
  }

  // ---- semi-hard
  float in_v = -INFINITY;
  int in_i = -1, in_c = 0;
  for (int j = t; j < B; j += TFA_THREADS)
    if (sflag[j] == 2) fold<false>(srow[j], j, in_v, in_i, in_c);
  block_extreme<false>(in_v, in_i, in_c, s_v, s_i, s_c);
  const float inside = in_c > 0 ? in_v : 0.f;                // negatives_inside (row minimum 0 as filler)
  // ordered list of this anchor's positives (ascending column), built chunk by chunk with warp ballots
  unsigned short* spos = reinterpret_cast<unsigned short*>(sflag + ((B + 1) & ~1));   // [n_pos] <= [B]
  {
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
  double loss_sum = 0.0;
  for (int ia = 0; ia < n_pos; ++ia) {
    const int a = spos[ia];
    const float pa = srow[a];
    float out_v = INFINITY;
    int out_i = -1, out_c = 0, n_mask_t = 0;
    for (int k = t; k < B; k += TFA_THREADS) {
      if (sflag[k] == 2 && srow[k] > pa) {
        ++n_mask_t;
        fold<true>(__fsub_rn(srow[k], rmax), k, out_v, out_i, out_c);
      }
    }
    block_extreme<true>(out_v, out_i, out_c, s_v, s_i, s_c);
    const bool outside = out_c > 0;                          // mask_final
    const float sh = outside ? __fadd_rn(out_v, rmax) : inside;
    const float lm = __fadd_rn(margin, __fsub_rn(pa, sh));
    loss_sum += (double)fmaxf(lm, 0.f);
    if (Cf && lm >= 0.f) {
      if (outside) {
        int n_tied = out_c;
        if (out_v == 0.f) n_tied += B - block_count(n_mask_t, s_c);   // every unmasked entry ties at 0
        const float c_neg = 1.f / (float)n_tied;
        const float c_max = out_v == 0.f ? (1.f - (float)out_c / (float)n_tied) / (float)rmax_c : 0.f;
        for (int k = t; k < B; k += TFA_THREADS) {
          float cf = scf[k];
          if (k == a) cf += 1.f;
          if (sflag[k] == 2 && srow[k] > pa && __fsub_rn(srow[k], rmax) == out_v) cf -= c_neg;
          if (out_v == 0.f && srow[k] == rmax) cf -= c_max;
          scf[k] = cf;
        }
      } else {
        const float c_neg = in_c > 0 ? 1.f / (float)in_c : 0.f;
        for (int k = t; k < B; k += TFA_THREADS) {
          float cf = scf[k];
          if (k == a) cf += 1.f;
          if (sflag[k] == 2 && srow[k] == in_v) cf -= c_neg;
          scf[k] = cf;
        }
      }
    }
  }
  if (t == 0) {
    rows[b].loss_sum = loss_sum;
    rows[b].n_pos = n_pos;
  }
  if (Cf)
    for (int j = t; j < B; j += TFA_THREADS) Cf[(size_t)b * ldp + j] = scf[j];   // each thread re-reads its own entries
}

// ---------------------------------------------------------------- K3: scalar loss and the backward scale
__global__ void __launch_bounds__(1024) tfa_finalize_kernel(const TfaRow* __restrict__ rows, int B, int kind, float dloss,
                                                            float* __restrict__ loss, float* __restrict__ scale) {
  __shared__ double s_sum[1024];
  __shared__ long long s_cnt[1024];
  double sum = 0.0;
  long long cnt = 0;
  for (int b = threadIdx.x; b < B; b += 1024) {
    sum += rows[b].loss_sum;
    cnt += rows[b].n_pos;
  }
  s_sum[threadIdx.x] = sum;
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 512; o >= 1; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double denom = kind == DIF_TFA_HARD ? (double)B : (double)s_cnt[0];   // 0 positive pairs: 0 / 0 = NaN as in tfa
    loss[0] = (float)(s_sum[0] / denom);
    scale[0] = (float)((double)dloss / denom);
  }
}

// ---------------------------------------------------------------- K4: P_ij <- (Cf_ij + Cf_ji) * dP_ij/d(sq)-factor
// L2: d P_ij / d x_i = (x_i - x_j) / P_ij;  squared-L2: 2 (x_i - x_j);  nothing where P_ij == 0 (error mask, diagonal)
__global__ void __launch_bounds__(256) tfa_fold_kernel(float* __restrict__ P, const float* __restrict__ Cf, int ldp, int B,
                                                       int squared) {
  __shared__ float tile[32][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += 8) {
    // the mirrored tile of Cf: rows of the j range, columns of the i range, read coalesced
    const int ti = blockIdx.x * 32 + r, tj = blockIdx.y * 32 + threadIdx.x;
    tile[r][threadIdx.x] = (ti < B && tj < B) ? Cf[(size_t)ti * ldp + tj] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = blockIdx.y * 32 + r;
    if (i < B && j < B) {
      const float p = P[(size_t)i * ldp + j];
      const float c = Cf[(size_t)i * ldp + j] + tile[threadIdx.x][r];   // tile[x][r] = Cf[j][i]
      P[(size_t)i * ldp + j] = (p > 0.f && c != 0.f) ? (squared ? 2.f * c : c / p) : 0.f;
    }
  }
}

// ---------------------------------------------------------------- K5: dX_i = scale * sum_j W_ij (x_i - x_j)
template <int KQ>
__global__ void __launch_bounds__(TFA_THREADS) tfa_grad_kernel(const float* __restrict__ Wt, int ldp,
                                                               const float* __restrict__ x, int B, int D,
                                                               const float* __restrict__ scale, float* __restrict__ dX) {
  __shared__ int s_j[TFA_THREADS];
  __shared__ float s_w[TFA_THREADS];
  __shared__ int s_n[TFA_WARPS];
  const int i = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  float xi[KQ], acc[KQ];
#pragma unroll
  for (int q = 0; q < KQ; ++q) {
    const int d = t + q * TFA_THREADS;
    xi[q] = d < D ? x[(size_t)i * D + d] : 0.f;
    acc[q] = 0.f;
  }
  for (int c0 = 0; c0 < B; c0 += TFA_THREADS) {
    const int j = c0 + t;
    const float w = j < B ? Wt[(size_t)i * ldp + j] : 0.f;
    const unsigned nz = __ballot_sync(0xffffffffu, w != 0.f);
    __syncthreads();   // previous list consumed
    if (lane == 0) s_n[warp] = __popc(nz);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int wv = 0; wv < TFA_WARPS; ++wv) {
      if (wv < warp) base += s_n[wv];
      total += s_n[wv];
    }
    if (total == 0) continue;   // block-uniform
    if (w != 0.f) {
      const int at = base + __popc(nz & ((1u << lane) - 1u));
      s_j[at] = j;
      s_w[at] = w;
    }
    __syncthreads();
    for (int e = 0; e < total; ++e) {   // ascending column order: fixed summation order
      const float we = s_w[e];
      const float* xj = x + (size_t)s_j[e] * D;
#pragma unroll
      for (int q = 0; q < KQ; ++q) {
        const int d = t + q * TFA_THREADS;
        if (d < D) acc[q] = __fmaf_rn(we, xi[q] - xj[d], acc[q]);
      }
    }
  }
  const float sc = scale[0];
#pragma unroll
  for (int q = 0; q < KQ; ++q) {
    const int d = t + q * TFA_THREADS;
    if (d < D) dX[(size_t)i * D + d] = sc * acc[q];
  }
}

struct TfaWorkspace {
  float* P = nullptr;
  float* Cf = nullptr;
  TfaRow* rows = nullptr;
  float* scale = nullptr;
  size_t mat_cap = 0, row_cap = 0;
  int ensure(size_t mat, size_t n_rows, bool want_cf) {
    if (mat > mat_cap) {
      cudaFree(P);
      cudaFree(Cf);
      P = Cf = nullptr;
      mat_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&P, mat * sizeof(float)));
      mat_cap = mat;
    }
    if (want_cf && !Cf) DIF_CUDA_OK(cudaMalloc((void**)&Cf, mat_cap * sizeof(float)));
    if (n_rows > row_cap) {
      cudaFree(rows);
      rows = nullptr;
      row_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&rows, n_rows * sizeof(TfaRow)));
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
    DIF_CUDA_OK(cudaFuncSetAttribute(tfa_row_kernel<DIF_TFA_HARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    DIF_CUDA_OK(cudaFuncSetAttribute(tfa_row_kernel<DIF_TFA_SEMIHARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  // K1
  const int row_blocks = (B + BH_RB - 1) / BH_RB;
  const int sms = std::max(1, device_sm_count());
  int splits = std::max(1, std::min((2 * sms + row_blocks - 1) / row_blocks, (B + BH_CB - 1) / BH_CB));
  int cols = (B + splits - 1) / splits;
  cols = (cols + BH_CB - 1) / BH_CB * BH_CB;
  splits = (B + cols - 1) / cols;
  const size_t smem1 = ((size_t)(BH_RB + BH_CB) * D + BH_RB + BH_CB) * sizeof(float);
  tfa_pdist_kernel<<<dim3(row_blocks, splits), BH_WARPS * 32, smem1, st>>>(emb, B, D, cols, squared, g_tfa.P, ldp);
  DIF_LAUNCH_OK();
  // K2
  const size_t smem2 = (size_t)B * 11 + 32;   // row, coefficient row (fp32), flags (u8), positives list (u16)
  float* cf = demb ? g_tfa.Cf : nullptr;
  if (base == DIF_TFA_HARD)
    tfa_row_kernel<DIF_TFA_HARD><<<B, TFA_THREADS, smem2, st>>>(g_tfa.P, ldp, labels, B, margin, soft, g_tfa.rows, pos_idx,
                                                               neg_idx, cf);
  else
    tfa_row_kernel<DIF_TFA_SEMIHARD><<<B, TFA_THREADS, smem2, st>>>(g_tfa.P, ldp, labels, B, margin, 0, g_tfa.rows, nullptr,
                                                                   nullptr, cf);
  DIF_LAUNCH_OK();
  // K3
  tfa_finalize_kernel<<<1, 1024, 0, st>>>(g_tfa.rows, B, base, dloss, loss, g_tfa.scale);
  DIF_LAUNCH_OK();
  if (!demb) return DIF_OK;
  // K4, K5
  const int tb = (B + 31) / 32;
  tfa_fold_kernel<<<dim3(tb, tb), dim3(32, 8), 0, st>>>(g_tfa.P, g_tfa.Cf, ldp, B, squared);
  DIF_LAUNCH_OK();
  const int kq = (D + TFA_THREADS - 1) / TFA_THREADS;
  if (kq <= 1) tfa_grad_kernel<1><<<B, TFA_THREADS, 0, st>>>(g_tfa.P, ldp, emb, B, D, g_tfa.scale, demb);
  else if (kq <= 2) tfa_grad_kernel<2><<<B, TFA_THREADS, 0, st>>>(g_tfa.P, ldp, emb, B, D, g_tfa.scale, demb);
  else tfa_grad_kernel<4><<<B, TFA_THREADS, 0, st>>>(g_tfa.P, ldp, emb, B, D, g_tfa.scale, demb);
  DIF_LAUNCH_OK();
  return DIF_OK;
}
