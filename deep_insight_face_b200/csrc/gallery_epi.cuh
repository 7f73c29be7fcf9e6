// Fused top-k epilogue of the gallery-search tensor-core pass, shared by the per-precision
// translation units (gallery_tc_*.cu) and the host code in gallery.cu.
#pragma once
#include "dif_canon.cuh"
#include "nt_gemm.cuh"

namespace dif {

constexpr int kGalBN = 256;

// Error bound of the tensor-core pass relative to |q| * |g| (cosine: both are 1).
//   3xTF32: dropped lo*lo terms + truncation of lo (2^-20) + <= 3*D/8 fp32 accumulations
//   1xTF32: operand truncation 2 * 2^-10;   bf16: operand rounding 2 * 2^-9
__host__ __device__ inline float mode_eps(int precision) {
  // 3xBF16: |x - b0 - b1| <= 2^-18 |x|, so the three dropped terms stay below 1.1e-5 of sum |a||b| (9x margin)
  return (precision == DIF_PREC_TF32X3 || precision == DIF_PREC_BF16X3) ? 1.0e-4f
                                                                          : (precision == DIF_PREC_BF16 ? 5.0e-3f : 2.5e-3f);
}

// Half-width of the window around the k-th best approximate score inside which a row can still
// belong to the exact top-k.  metric 1: scores are cosines of unit rows; metric 0: scores are
// 2 q.g - |g|^2 with |g|^2 <= gmax_sq.  The second term covers the canonical fp32 rounding.
__host__ __device__ inline float window_eps(int metric, float eps_rel, float q_sq, float gmax_sq) {
  if (metric == 1) return eps_rel + 4e-6f;
  const float qn = sqrtf(fmaxf(q_sq, 0.f)), gn = sqrtf(fmaxf(gmax_sq, 0.f));
  return 2.f * eps_rel * qn * gn + 4e-6f * (qn + gn) * (qn + gn);
}

__device__ __forceinline__ float orderable_to_float(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}
__device__ __forceinline__ float key_score(uint64_t key) { return orderable_to_float((uint32_t)(key >> 32)); }

// ------------------------------------------------------------------------------------------
// Per-query candidate list per gallery split, kept in shared memory ([slot][thread] layout so a
// warp's accesses to one slot are 256 contiguous bytes).
//
// Fast path per 32 columns: a max-reduction of the 32 scores and one compare against the thread's
// drop threshold; the warp leaves it only if some lane has a score above its threshold.  The slow
// path builds each lane's bitmask of such columns and every lane walks its own mask, fetching the
// score with a register select tree (pick32): a single call site of insert(), no per-column code
// replication, and no TMEM re-read (round 1 re-read the union of the warp's hit columns one at a
// time, a few hundred cycles of tcgen05.ld latency each).
//
// Drop threshold of a thread = max(minimum of its own list once full,
//                                  shared lower bound on the query's k-th best score - margin).
// The shared bound (one u32 per query in HBM, atomicMax of every full list's minimum) is what keeps
// a restarted list - a new split, a new item - from re-inserting hundreds of rows: any full list's
// minimum is a lower bound on the global k-th best approximate score A_k, so a row below
// bound - margin (margin > 2 * window_eps) lies outside [A_k - 2 eps, inf) and cannot belong to the
// exact top-k (see rerank_kernel for the rest of the argument).
// ------------------------------------------------------------------------------------------
template <int METRIC>
struct TopkEpi {
  struct Params {
    uint64_t* cand;          // [rows padded][n_splits][kp]
    const float* gnorm;      // [n_rows padded to tile] canonical |g|^2 (METRIC 0)
    unsigned int* bound;     // [rows padded] shared lower bound on the k-th best score (orderable u32, 0 = none)
    unsigned int* maxima;    // [n_splits][q_pad] best score seen so far by each split's list (orderable u32, 0 = none)
    int q_pad;
    int k;                   // result size: the k-th largest of the split maxima is a second lower bound
    const float* q_sq;       // [rows padded] canonical |q|^2 (METRIC 0 margin)
    const unsigned int* gmax;  // orderable max |g|^2
    int n_rows;
    int n_splits;
    int kp;
    float eps_rel;
  };
  // lists [kp][128] u64 + selection scratch [kp][128] f32
  __host__ __device__ static int smem_bytes_kp(int kp) { return kp * GEMM_BM * 12; }
  static int smem_bytes(const Params& pp) { return smem_bytes_kp(pp.kp); }

  const Params& p;
  uint64_t* keys;  // this thread's slot 0; slot s at keys[s * GEMM_BM]
  float* sel;      // selection scratch, slot s at sel[s * GEMM_BM]
  unsigned int* my_bound;
  unsigned int* my_max;    // this (split, query)'s cell of `maxima`
  const unsigned int* max_col;  // maxima + m_row: stride q_pad between splits
  float own_max;
  int tiles_seen;
  float thr;      // drop threshold
  float own_min;  // minimum of the own list once it is full, else -inf
  float shared;   // last value read from the shared bound, minus margin
  float margin;
  int min_slot;
  int fill;

  __device__ TopkEpi(const Params& pp, uint8_t* smem, int row)
      : p(pp), keys(reinterpret_cast<uint64_t*>(smem) + row),
        sel(reinterpret_cast<float*>(smem + (size_t)pp.kp * GEMM_BM * 8) + row), my_bound(nullptr), my_max(nullptr),
        max_col(nullptr), own_max(-INFINITY), tiles_seen(0), thr(-INFINITY), own_min(-INFINITY), shared(-INFINITY),
        margin(0.f), min_slot(0), fill(0) {}

  __device__ __forceinline__ void refresh() {
    const unsigned int b = *reinterpret_cast<volatile unsigned int*>(my_bound);
    shared = fmaxf(shared, b ? orderable_to_float(b) - margin : -INFINITY);
    thr = fmaxf(own_min, shared);
  }
  // Second lower bound on the k-th best score: the maxima of k different splits are k different rows, so the
  // k-th largest split maximum is reached by at least k rows.  Unlike a single list's minimum it tracks the
  // k-th best of EVERYTHING scanned so far for this query, by any CTA.
  __device__ __forceinline__ void refresh_maxima() {
    const int k = p.k;
    if (p.n_splits < k) return;
    int n = 0, ms = 0;
    float mn = INFINITY;
    for (int s0 = 0; s0 < p.n_splits; s0 += 8) {
      unsigned int o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)   // eight independent loads in flight
        o[j] = (s0 + j < p.n_splits) ? *reinterpret_cast<const volatile unsigned int*>(max_col + (size_t)(s0 + j) * p.q_pad) : 0u;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!o[j]) continue;
        const float v = orderable_to_float(o[j]);
        if (n < k) {
          sel[n * GEMM_BM] = v;
          if (v < mn) { mn = v; ms = n; }
          ++n;
        } else if (v > mn) {
          sel[ms * GEMM_BM] = v;
          mn = sel[0];
          ms = 0;
          for (int i = 1; i < k; ++i) {
            const float w = sel[i * GEMM_BM];
            if (w < mn) { mn = w; ms = i; }
          }
        }
      }
    }
    if (n == k) shared = fmaxf(shared, mn - margin);
  }

  __device__ void begin_item(int m_row, int split, int) {
    for (int s = 0; s < p.kp; ++s) keys[s * GEMM_BM] = 0ull;
    own_min = -INFINITY;
    own_max = -INFINITY;
    shared = -INFINITY;
    min_slot = 0;
    fill = 0;
    tiles_seen = 0;
    my_bound = p.bound + m_row;
    max_col = p.maxima + m_row;
    my_max = p.maxima + (size_t)split * p.q_pad + m_row;
    const float eps = window_eps(METRIC, p.eps_rel, METRIC == 0 ? p.q_sq[m_row] : 1.f,
                                 METRIC == 0 ? orderable_to_float(*p.gmax) : 1.f);
    margin = 2.f * eps * 1.001f + 1e-30f;
    refresh();
  }

  __device__ __forceinline__ void begin_tile(int) {
    // the bound improves like log(columns seen): look at tiles 1, 2, 4, 8, 16, ... of the item
    ++tiles_seen;
    if ((tiles_seen & (tiles_seen - 1)) == 0) refresh_maxima();
    refresh();
  }

  __device__ __forceinline__ void insert(float v, int col) {
    const uint64_t key = make_key(v, (uint32_t)col);
    if (v > own_max) {
      own_max = v;
      *reinterpret_cast<volatile unsigned int*>(my_max) = (unsigned int)(key >> 32);
    }
    if (fill < p.kp) {
      keys[fill * GEMM_BM] = key;
      if (++fill < p.kp) return;
    } else {
      keys[min_slot * GEMM_BM] = key;
    }
    uint64_t mn = keys[0];
    int ms = 0;
    for (int s = 1; s < p.kp; ++s) {
      const uint64_t k2 = keys[s * GEMM_BM];
      if (k2 < mn) {
        mn = k2;
        ms = s;
      }
    }
    min_slot = ms;
    own_min = key_score(mn);
    atomicMax(my_bound, (unsigned int)(mn >> 32));
    thr = fmaxf(own_min, shared);  // the shared bound is re-read once per tile, not per insert
  }

  __device__ __forceinline__ float score_of(uint32_t acc_bits, int col) const {
    const float a = __uint_as_float(acc_bits);
    return METRIC == 1 ? a : __fmaf_rn(2.f, a, -__ldg(p.gnorm + col));
  }

  __device__ __forceinline__ void consume(int col0, const uint32_t (&acc)[32], uint32_t taddr,
                                          uint32_t (&pending)[32]) {
    float v[32];
    if (METRIC == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
    } else {
      const float4* gn = reinterpret_cast<const float4*>(p.gnorm + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 g = __ldg(gn + i);
        v[4 * i + 0] = __fmaf_rn(2.f, __uint_as_float(acc[4 * i + 0]), -g.x);
        v[4 * i + 1] = __fmaf_rn(2.f, __uint_as_float(acc[4 * i + 1]), -g.y);
        v[4 * i + 2] = __fmaf_rn(2.f, __uint_as_float(acc[4 * i + 2]), -g.z);
        v[4 * i + 3] = __fmaf_rn(2.f, __uint_as_float(acc[4 * i + 3]), -g.w);
      }
    }
    const int valid = p.n_rows - col0;  // columns >= valid are TMA zero fill and must not compete
    if (valid < 32) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i >= valid) v[i] = -INFINITY;
    }
    float m = v[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) m = fmaxf(m, v[i]);
    if (__any_sync(0xffffffffu, m > thr)) {
      tmem_ld_wait(pending);  // no tcgen05.ld in flight while this path shuffles registers
      uint32_t mask = 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) mask |= (v[i] > thr) ? (1u << i) : 0u;   // (columns past the gallery hold -inf)
      while (mask) {  // per lane: its own hits, in column order; the score comes out of the chunk in registers
        const int i = __ffs((int)mask) - 1;
        mask &= mask - 1;
        const float s = pick32(v, i);
        if (s > thr) insert(s, col0 + i);   // thr rises with every insert into a full list
      }
    }
  }

  __device__ void end_item(int m_row, int split) {
    uint64_t* out = p.cand + ((size_t)m_row * p.n_splits + split) * p.kp;
    for (int s = 0; s < p.kp; ++s) out[s] = keys[s * GEMM_BM];
  }
};

// Implemented once per precision in gallery_tc_*.cu (keeps nvcc's per-file work parallel).
// ares = 1 selects the resident-A schedule (caller has checked that it fits).
int launch_search_tf32x3(int metric, int ctas, int ares, const CUtensorMap* maps, const GemmShape& shape,
                         const TopkEpi<1>::Params& ep, int n_units, cudaStream_t st);
int launch_search_bf16(int metric, int ctas, int ares, const CUtensorMap* maps, const GemmShape& shape,
                       const TopkEpi<1>::Params& ep, int n_units, cudaStream_t st);
int launch_search_tf32x1(int metric, int ctas, int ares, const CUtensorMap* maps, const GemmShape& shape,
                         const TopkEpi<1>::Params& ep, int n_units, cudaStream_t st);
int launch_search_bf16x3(int metric, int ctas, int ares, const CUtensorMap* maps, const GemmShape& shape,
                         const TopkEpi<1>::Params& ep, int n_units, cudaStream_t st);

template <int PREC>
int launch_search_prec(int metric, int ctas, int ares, const CUtensorMap* maps, const GemmShape& shape,
                       const TopkEpi<1>::Params& ep1, int n_units, cudaStream_t st) {
  // TopkEpi<0>::Params and TopkEpi<1>::Params have identical members
  TopkEpi<0>::Params ep0{ep1.cand, ep1.gnorm, ep1.bound, ep1.maxima, ep1.q_pad, ep1.k, ep1.q_sq, ep1.gmax, ep1.n_rows, ep1.n_splits, ep1.kp, ep1.eps_rel};
  if (ctas == 2) {
    if (ares) {
      if constexpr (PREC == 1 || PREC == 2) {   // the two-plane modes cannot keep a query block resident
        return metric == 1 ? launch_nt_gemm<PREC, kGalBN, 2, 1, TopkEpi<1>>(maps, shape, ep1, n_units, st)
                           : launch_nt_gemm<PREC, kGalBN, 2, 1, TopkEpi<0>>(maps, shape, ep0, n_units, st);
      }
    }
    return metric == 1 ? launch_nt_gemm<PREC, kGalBN, 2, 0, TopkEpi<1>>(maps, shape, ep1, n_units, st)
                       : launch_nt_gemm<PREC, kGalBN, 2, 0, TopkEpi<0>>(maps, shape, ep0, n_units, st);
  }
  return metric == 1 ? launch_nt_gemm<PREC, kGalBN, 1, 0, TopkEpi<1>>(maps, shape, ep1, n_units, st)
                     : launch_nt_gemm<PREC, kGalBN, 1, 0, TopkEpi<0>>(maps, shape, ep0, n_units, st);
}

}  // namespace dif
