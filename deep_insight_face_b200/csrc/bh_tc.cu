// Batch-hard mining on the tensor cores for large batches (B >= 512): the B x B distance matrix is produced by
// the tcgen05/TMA NT-GEMM skeleton in 3xTF32 (fp32-exact to ~1e-6) and consumed straight out of TMEM by an
// epilogue that keeps, per anchor, the 4 best positive / negative / overall candidates in registers
// (common/losses.py:42-46,67-71); the matrix never reaches HBM.  A re-rank kernel then recomputes the handful of
// candidates inside the filter's error window in canonical fp32, so the mined indices, tie counts and values are
// bit-identical to the CUDA-core miner and to the oracle.  An anchor whose candidate list could be incomplete
// (fourth candidate still inside the window) is re-scanned canonically in full.
#include <algorithm>

#include "bh_tile.cuh"
#include "nt_gemm.cuh"
#include "prep_rows.cuh"

namespace dif {

constexpr int kMineBN = 256;
constexpr int kMineM = 4;   // candidates kept per class per (anchor, split)

struct BhCand {             // per (split, anchor)
  float pv[kMineM]; int pi[kMineM];   // positives, best first (cosine: ascending, euclid: descending)
  float nv[kMineM]; int ni[kMineM];   // negatives, best first (cosine: descending, euclid: ascending)
  float av[kMineM]; int ai[kMineM];   // overall maxima (euclid filler), descending
  float row_sum; int n_pos;
};

template <bool SMALLER_IS_BETTER>
__device__ __forceinline__ bool better(float a, float b) { return SMALLER_IS_BETTER ? a < b : a > b; }

// insert (v, j) into a best-first list of kMineM; equal values keep the earlier (lower) column first
template <bool SMALLER_IS_BETTER>
__device__ __forceinline__ void cand_insert(float (&v)[kMineM], int (&idx)[kMineM], float x, int j) {
  if (!(idx[kMineM - 1] < 0 || better<SMALLER_IS_BETTER>(x, v[kMineM - 1]))) return;
  v[kMineM - 1] = x;
  idx[kMineM - 1] = j;
#pragma unroll
  for (int s = kMineM - 1; s > 0; --s) {
    const bool sw = idx[s - 1] < 0 || better<SMALLER_IS_BETTER>(v[s], v[s - 1]);
    const float tv = sw ? v[s - 1] : v[s];
    const int ti = sw ? idx[s - 1] : idx[s];
    v[s - 1] = sw ? v[s] : v[s - 1];
    idx[s - 1] = sw ? idx[s] : idx[s - 1];
    v[s] = tv;
    idx[s] = ti;
  }
}

template <bool COSINE>
struct MineEpi {
  struct Params {
    BhCand* cand;            // [n_slots][B]
    const int32_t* labels;   // [B]
    const float* sq;         // [B] canonical sum of squares (euclid)
    const unsigned int* gmax_sq;   // orderable max of sq (euclid): scales the error window like bh_rerank_kernel
    int B;
  };
  // two warps per TMEM lane quarter: the epilogue (about 37 instructions per column) is latency bound with one
  static constexpr int kEpiWarps = 8;
  static constexpr int kEpiThreads = kEpiWarps * 32;
  // labels + sq of the current tile, then the label range [min, max] of each 32-column chunk
  static int smem_bytes(const Params&) { return kMineBN * 8 + 2 * (kMineBN / 32) * 4; }

  const Params& p;
  int* s_lab;
  float* s_sq;
  int* s_lmin;   // [kMineBN / 32] smallest / largest label of each chunk of the current tile: a warp none of whose
  int* s_lmax;   // anchors' labels falls inside a chunk's range has no positives there and skips the label compares
  int e_tid;     // index among the 256 epilogue threads: half * 128 + row
  float pv[kMineM], nv[kMineM], av[kMineM];
  int pi[kMineM], ni[kMineM], ai[kMineM];
  float bp, bn, ba;   // best signed value seen so far per class (positives, negatives, overall): the window's anchor
  float row_sum, my_sq, win;
  int n_pos, my_lab, my_row, tile0;

  __device__ MineEpi(const Params& pp, uint8_t* smem, int row, int half)
      : p(pp), s_lab(reinterpret_cast<int*>(smem)), s_sq(reinterpret_cast<float*>(smem) + kMineBN),
        s_lmin(reinterpret_cast<int*>(smem) + 2 * kMineBN), s_lmax(reinterpret_cast<int*>(smem) + 2 * kMineBN + kMineBN / 32),
        e_tid(half * GEMM_BM + row), bp(-INFINITY), bn(-INFINITY), ba(-INFINITY), row_sum(0.f), my_sq(0.f), win(0.f),
        n_pos(0), my_lab(-1), my_row(0), tile0(0) {}

  __device__ void begin_item(int m_row, int, int) {
    my_row = m_row;
    my_lab = m_row < p.B ? p.labels[m_row] : -1;
    my_sq = (!COSINE && m_row < p.B) ? p.sq[m_row] : 0.f;
    // A candidate only matters if it lies inside the re-rank window (2 eps of bh_rerank_kernel) of the row's best:
    // a column is inserted iff it is within `win` of the best value of its class seen so far in this work item (a
    // list that fills up inside the window is detected by the re-rank, which then re-scans the row canonically).
    if (COSINE) {
      win = 4.1e-5f;
    } else {
      const uint32_t o = *p.gmax_sq;
      win = 4.1e-5f * (my_sq + __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o));
    }
#pragma unroll
    for (int s = 0; s < kMineM; ++s) {
      pv[s] = nv[s] = av[s] = 0.f;
      pi[s] = ni[s] = ai[s] = -1;
    }
    if (COSINE) av[0] = -INFINITY;
    bp = bn = ba = -INFINITY;
    row_sum = 0.f;
    n_pos = 0;
  }

  // all 256 epilogue threads: stage the tile's labels (and squared norms) in shared memory
  __device__ void begin_tile(int col_begin) {
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the previous tile's readers are done
    for (int i = e_tid; i < kMineBN; i += kEpiThreads) {
      const int j = col_begin + i;
      s_lab[i] = j < p.B ? p.labels[j] : -2;
      if (!COSINE) s_sq[i] = j < p.B ? p.sq[j] : 0.f;
    }
    tile0 = col_begin;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (e_tid < kMineBN / 32) {   // PK batches arrive grouped by identity: most chunks hold no positive of a given warp
      int lo = 0x7fffffff, hi = -0x7fffffff;
      for (int i = 0; i < 32; ++i) {
        const int l = s_lab[e_tid * 32 + i];
        if (col_begin + e_tid * 32 + i < p.B) {   // columns past the batch do not count
          lo = min(lo, l);
          hi = max(hi, l);
        }
      }
      s_lmin[e_tid] = lo;
      s_lmax[e_tid] = hi;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }

  __device__ __forceinline__ float dist_of(uint32_t acc_bits, int jl) const {
    const float a = __uint_as_float(acc_bits);
    return COSINE ? a : __fsub_rn(__fadd_rn(my_sq, s_sq[jl]), __fmul_rn(2.f, a));
  }

  // One chunk of 32 columns, in the shape of the gallery's TopkEpi.  "Signed" space: q = -d where smaller is better, so
  // larger q always wins.  Fast path, values only: per column a label compare, two selects and running maxima of the
  // positives' and the negatives' q (plus the row sum / row maximum for the statistics of losses.py:72-80) - no
  // indices, no lists.  Only a column within `win` (the re-rank window) of the best value seen so far can matter, so
  // the chunk is left alone unless its maximum reaches best - win.  Slow path (warp-uniform): the thresholds are first
  // raised by the chunk's own maxima - a column further than `win` below ANY value of its class is dead - then each
  // lane builds the mask of its surviving columns and inserts them (see pick32).  With 32 anchors in lockstep most
  // chunks of a fresh work item still take the slow path, but it now costs one mask pass plus about one insert.
  // CHECK = the chunk straddles the end of the batch (TMA zero fill).
  // NOPOS = no anchor of this warp has a positive in the chunk (label ranges): every column is a negative
  template <bool CHECK, bool NOPOS>
  __device__ __forceinline__ void consume_impl(int col0, const uint32_t (&acc)[32], uint32_t taddr, uint32_t (&pending)[32]) {
    const int base = col0 - tile0;
    constexpr float kSgnN = COSINE ? 1.f : -1.f;           // negatives: cosine keeps the largest, euclid the smallest
    float dv[32];
    float cp = -INFINITY, cn = -INFINITY, ca = -INFINITY, rs = 0.f;
    int np = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      dv[i] = dist_of(acc[i], base + i);
      const bool valid = !CHECK || (col0 + i < p.B);
      const bool same = !NOPOS && s_lab[base + i] == my_lab;   // columns past the batch carry label -2: never "same"
      const float sx = kSgnN * dv[i];
      if (!NOPOS) np += same ? 1 : 0;
      if (!NOPOS) cp = fmaxf(cp, same ? -sx : -INFINITY);
      cn = fmaxf(cn, (same || !valid) ? -INFINITY : sx);
      if (!(NOPOS && COSINE)) ca = fmaxf(ca, valid ? dv[i] : -INFINITY);
      rs += valid ? dv[i] : 0.f;
    }
    if (NOPOS && COSINE) ca = cn;   // all columns are negatives: the chunk's maximum is the negatives' maximum
    row_sum += rs;
    n_pos += np;   // positives of this anchor (itself included) among the columns of this work item
    if (COSINE) av[0] = fmaxf(av[0], ca);   // running row maximum for the max(dists) statistic
    // (a class absent from the chunk has maximum -inf, which must not pass against a still-empty best of -inf)
    const bool hit = (cp > -INFINITY && cp >= bp - win) || (cn > -INFINITY && cn >= bn - win) ||
                     (!COSINE && ca > -INFINITY && ca >= ba - win);
    if (!__any_sync(0xffffffffu, hit)) return;
    tmem_ld_wait(pending);   // warp-uniform: no tcgen05.ld in flight while registers are shuffled below
    bp = fmaxf(bp, cp);
    bn = fmaxf(bn, cn);
    if (!COSINE) ba = fmaxf(ba, ca);
    const float tp = bp - win, tn = bn - win, ta = ba - win;
    uint32_t mask = 0, mask_a = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const bool valid = !CHECK || (col0 + i < p.B);
      const bool same = !NOPOS && s_lab[base + i] == my_lab;
      const float sx = kSgnN * dv[i];
      const bool h = valid && ((same ? -sx : sx) >= (same ? tp : tn));
      mask |= h ? (1u << i) : 0u;
      if (!COSINE) mask_a |= (valid && dv[i] >= ta) ? (1u << i) : 0u;
    }
    // per-lane loop over the lane's own hits (usually one: the chunk's new best).  The value comes out of the
    // register block through a 31-select tree: a TMEM re-read costs a few hundred cycles of latency PER column and the
    // union of 32 lanes' hits is several columns in most chunks of a fresh work item.
    uint32_t todo = mask | mask_a;
    while (todo) {
      const int i = __ffs((int)todo) - 1;
      todo &= todo - 1;
      const float d = pick32(dv, i);
      const int j = col0 + i;
      if ((mask >> i) & 1u) {
        if (!NOPOS && s_lab[base + i] == my_lab) cand_insert<COSINE>(pv, pi, d, j);
        else cand_insert<!COSINE>(nv, ni, d, j);
      }
      if (!COSINE && ((mask_a >> i) & 1u)) cand_insert<false>(av, ai, d, j);
    }
  }

  __device__ __forceinline__ void consume(int col0, const uint32_t (&acc)[32], uint32_t taddr, uint32_t (&pending)[32]) {
    if (col0 >= p.B) return;                       // warp-uniform: the whole chunk is zero fill
    const int ch = (col0 - tile0) >> 5;
    const bool may_have_pos = my_lab >= s_lmin[ch] && my_lab <= s_lmax[ch];
    if (col0 + 32 > p.B) consume_impl<true, false>(col0, acc, taddr, pending);
    else if (__any_sync(0xffffffffu, may_have_pos)) consume_impl<false, false>(col0, acc, taddr, pending);
    else consume_impl<false, true>(col0, acc, taddr, pending);
  }

  __device__ void end_item(int m_row, int slot) {
    if (m_row >= p.B) return;
    BhCand c;
#pragma unroll
    for (int s = 0; s < kMineM; ++s) {
      c.pv[s] = pv[s]; c.pi[s] = pi[s];
      c.nv[s] = nv[s]; c.ni[s] = ni[s];
      c.av[s] = av[s]; c.ai[s] = ai[s];
    }
    c.row_sum = row_sum;
    c.n_pos = n_pos;
    p.cand[(size_t)slot * p.B + m_row] = c;
  }
};

// canonical dist(r, j), one warp (same arithmetic as bh_mine_kernel / the oracle)
template <bool COSINE>
__device__ __forceinline__ float canon_dist(const float* __restrict__ x, const float* __restrict__ aux, int D, int r,
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

// One class (positives / negatives / overall) of one anchor: pick the exact extreme among the candidates inside
// the error window of the best approximate value.  Returns false if a split's list may be incomplete.
template <bool COSINE, bool SMALLER_IS_BETTER>
__device__ __forceinline__ bool rerank_class(const BhCand* __restrict__ cand, int n_slots, int B, int r, int which,
                                             float eps, const float* __restrict__ x, const float* __restrict__ aux,
                                             int D, float& val, int& idx, int& cnt) {
  const int lane = threadIdx.x & 31;
  // best approximate value over all kept candidates
  float best = SMALLER_IS_BETTER ? INFINITY : -INFINITY;
  bool any = false;
  for (int t = lane; t < n_slots * kMineM; t += 32) {
    const BhCand& c = cand[(size_t)(t / kMineM) * B + r];
    const int s = t % kMineM;
    const int ci = which == 0 ? c.pi[s] : (which == 1 ? c.ni[s] : c.ai[s]);
    const float cv = which == 0 ? c.pv[s] : (which == 1 ? c.nv[s] : c.av[s]);
    if (ci >= 0) {
      any = true;
      if (better<SMALLER_IS_BETTER>(cv, best)) best = cv;
    }
  }
  for (int o = 16; o >= 1; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    if (better<SMALLER_IS_BETTER>(ob, best)) best = ob;
  }
  val = SMALLER_IS_BETTER ? INFINITY : -INFINITY;
  idx = -1;
  cnt = 0;
  if (!__any_sync(0xffffffffu, any)) return true;   // the class is empty for this anchor
  const float lim = SMALLER_IS_BETTER ? best + 2.f * eps : best - 2.f * eps;
  bool complete = true;
  for (int t0 = 0; t0 < n_slots * kMineM; t0 += 32) {
    const int t = t0 + lane;
    int ci = -1;
    float cv = 0.f;
    if (t < n_slots * kMineM) {
      const BhCand& c = cand[(size_t)(t / kMineM) * B + r];
      const int s = t % kMineM;
      ci = which == 0 ? c.pi[s] : (which == 1 ? c.ni[s] : c.ai[s]);
      cv = which == 0 ? c.pv[s] : (which == 1 ? c.nv[s] : c.av[s]);
    }
    const bool in_win = ci >= 0 && !better<SMALLER_IS_BETTER>(lim, cv);   // cv at least as good as lim
    if (in_win && (t % kMineM) == kMineM - 1) complete = false;            // a full list reaches into the window
    unsigned todo = __ballot_sync(0xffffffffu, in_win);
    while (todo) {
      const int src = __ffs((int)todo) - 1;
      todo &= todo - 1;
      const int j = __shfl_sync(0xffffffffu, ci, src);
      const float d = canon_dist<COSINE>(x, aux, D, r, j);
      if (better<SMALLER_IS_BETTER>(d, val)) {
        val = d;
        idx = j;
        cnt = 1;
      } else if (d == val) {
        ++cnt;
        idx = j < idx ? j : idx;
      }
    }
  }
  return __all_sync(0xffffffffu, complete);
}

template <bool COSINE>
__device__ __forceinline__ void bh_rerank_row(const BhCand* __restrict__ cand, int n_slots, const float* __restrict__ x,
                                              const int32_t* __restrict__ labels, const float* __restrict__ aux,
                                              const unsigned int* __restrict__ gmax_sq, int B, int D,
                                              BhRec* __restrict__ out, unsigned long long* __restrict__ gmax_key,
                                              const BhFinalize& fin, int r, double* part) {
  const int lane = threadIdx.x & 31;
  // error window of the 3xTF32 pass (measured <= 1.2e-6 of sum |a||b|; 16x margin) plus the fp32 rounding of the sums
  float eps;
  if (COSINE) {
    eps = 2e-5f;
  } else {
    const uint32_t o = *gmax_sq;
    const float gsq = __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
    eps = 2e-5f * (aux[r] + gsq);
  }
  BhRec rec;
  const bool ok_p = rerank_class<COSINE, COSINE>(cand, n_slots, B, r, 0, eps, x, aux, D, rec.pos_val, rec.pos_idx, rec.pos_cnt);
  const bool ok_n = rerank_class<COSINE, !COSINE>(cand, n_slots, B, r, 1, eps, x, aux, D, rec.neg_val, rec.neg_idx, rec.neg_cnt);
  bool ok_a = true;
  rec.all_max = -INFINITY;
  rec.all_idx = -1;
  rec.all_cnt = 0;
  if (!COSINE) ok_a = rerank_class<COSINE, false>(cand, n_slots, B, r, 2, eps, x, aux, D, rec.all_max, rec.all_idx, rec.all_cnt);
  if (!(ok_p && ok_n && ok_a)) {
    // some candidate list may be incomplete: canonical scan of the whole row (rare: duplicates / degenerate batches)
    const int my_lab = labels[r];
    rec.pos_val = COSINE ? INFINITY : -INFINITY; rec.neg_val = COSINE ? -INFINITY : INFINITY; rec.all_max = -INFINITY;
    rec.pos_idx = rec.neg_idx = rec.all_idx = -1;
    rec.pos_cnt = rec.neg_cnt = rec.all_cnt = 0;
    for (int j = 0; j < B; ++j) {
      const float d = canon_dist<COSINE>(x, aux, D, r, j);
      fold<false>(d, j, rec.all_max, rec.all_idx, rec.all_cnt);
      if (labels[j] == my_lab) fold<COSINE>(d, j, rec.pos_val, rec.pos_idx, rec.pos_cnt);
      else fold<!COSINE>(d, j, rec.neg_val, rec.neg_idx, rec.neg_cnt);
    }
  }
  // row sum and number of positives (the anchor itself included): one record per lane, folded by a fixed butterfly
  float rs = 0.f;
  int np_ = 0;
  for (int s = lane; s < n_slots; s += 32) {
    const BhCand& c = cand[(size_t)s * B + r];
    rs += c.row_sum;
    np_ += c.n_pos;
  }
  for (int o = 16; o >= 1; o >>= 1) {
    rs += __shfl_xor_sync(0xffffffffu, rs, o);
    np_ += __shfl_xor_sync(0xffffffffu, np_, o);
  }
  rec.row_sum = rs;
  rec.n_pos = np_;
  rec.pos_sum = 0.f;
  if (COSINE) {   // max(dists) is only a printed statistic for the cosine loss: the filter's value is exact enough
    float am = -INFINITY;
    for (int s = lane; s < n_slots; s += 32) am = fmaxf(am, cand[(size_t)s * B + r].av[0]);
    for (int o = 16; o >= 1; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
    rec.all_max = am;
    rec.all_idx = 0;
    rec.all_cnt = 1;
  }
  if (lane == 0) {
    out[r] = rec;
    if (rec.all_cnt > 0)
      atomicMax(gmax_key, ((unsigned long long)float_orderable(rec.all_max) << 32) | (0xFFFFFFFFu - (unsigned)r));
    if (COSINE && fin.rows) {
      double sums[5];
      bh_finalize_row<COSINE>(rec, r, B, fin.alpha, fin.soft, fin.dloss, 0.f, fin.loss, fin.pos_idx, fin.neg_idx, fin.rows,
                              fin.compact, sums);
      sums[4] = 0.0;   // (the filler share is a squared-L2 matter)
      for (int k = 0; k < 5; ++k) part[k] = sums[k];
    }
  }
}

// The cosine loss needs nothing global to finalize an anchor (its fillers are the constants 1 / -1), so the
// re-rank warp finalizes its anchor on the spot (fin.rows != NULL) and the block leaves one partial-sum record for
// the statistics; the squared-L2 loss waits for max(dists) and keeps the separate bh_finalize_kernel.
constexpr int kRerankWarps = 8;
static_assert(kRerankWarps == kBhFusedFinalizeRows, "one partial-sum record per re-rank block");
template <bool COSINE>
__global__ void __launch_bounds__(kRerankWarps * 32) bh_rerank_kernel(const BhCand* __restrict__ cand, int n_slots,
                                                        const float* __restrict__ x, const int32_t* __restrict__ labels,
                                                        const float* __restrict__ aux, const unsigned int* __restrict__ gmax_sq,
                                                        int B, int D, BhRec* __restrict__ out,
                                                        unsigned long long* __restrict__ gmax_key, BhFinalize fin) {
  __shared__ double s_part[kRerankWarps][5];
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kRerankWarps + (threadIdx.x >> 5);
  if (COSINE && fin.rows) {
    if (lane == 0)   // (the only lane that writes this record, here and at the end of the row)
      for (int k = 0; k < 5; ++k) s_part[threadIdx.x >> 5][k] = 0.0;
    if (r < B) bh_rerank_row<COSINE>(cand, n_slots, x, labels, aux, gmax_sq, B, D, out, gmax_key, fin, r, s_part[threadIdx.x >> 5]);
    __syncthreads();
    if (threadIdx.x < 5) {
      double t = 0.0;
      for (int w = 0; w < kRerankWarps; ++w) t += s_part[w][threadIdx.x];
      fin.partials[(size_t)blockIdx.x * 5 + threadIdx.x] = t;
    }
    return;
  }
  if (r < B) bh_rerank_row<COSINE>(cand, n_slots, x, labels, aux, gmax_sq, B, D, out, gmax_key, fin, r, s_part[0]);
}

struct TcWorkspace {
  float *hi = nullptr, *lo = nullptr;
  unsigned int* gmax = nullptr;
  size_t plane_cap = 0;
  BhCand* cand = nullptr;
  size_t cand_cap = 0;
  int ensure(size_t plane_elems, size_t n_cand) {
    if (plane_elems > plane_cap) {
      retire_device_block(hi); retire_device_block(lo);
      hi = lo = nullptr; plane_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&hi, plane_elems * 4));
      DIF_CUDA_OK(cudaMalloc((void**)&lo, plane_elems * 4));
      plane_cap = plane_elems;
    }
    if (!gmax) DIF_CUDA_OK(cudaMalloc((void**)&gmax, 4));
    if (n_cand > cand_cap) {
      retire_device_block(cand);
      cand = nullptr; cand_cap = 0;
      DIF_CUDA_OK(cudaMalloc((void**)&cand, n_cand * sizeof(BhCand)));
      cand_cap = n_cand;
    }
    return DIF_OK;
  }
};
static thread_local TcWorkspace g_tc;

// Mines every anchor with the tensor-core filter + canonical re-rank and writes ONE merged record per anchor
// (recs [B]) plus aux [B] (inverse norm | sum of squares), exactly what bh_mine_kernel + the split merge produce.
template <bool COSINE>
int bh_mine_tensor(const float* emb, const int32_t* labels, int B, int D, BhRec* recs, float* aux,
                   unsigned long long* gmax_key, const BhFinalize& fin, cudaStream_t st) {
  const int sms = std::max(1, device_sm_count());
  GemmShape shape{};
  shape.m_blocks = (B + GEMM_BM - 1) / GEMM_BM;
  shape.n_tiles = (B + kMineBN - 1) / kMineBN;
  shape.k_chunks = (D + 31) / 32;
  // work item = (128 anchors, tiles_per_split column tiles): the split with the fewest waves x tiles wins, ties go to
  // the longer item (the candidate lists restart with every item and the re-rank reads one record per split)
  {
    long best_cost = -1;
    for (int tps = 1; tps <= shape.n_tiles; ++tps) {
      const int ns = (shape.n_tiles + tps - 1) / tps;
      const long cost = (long)((shape.m_blocks * ns + sms - 1) / sms) * tps;
      if (best_cost < 0 || cost <= best_cost) {
        best_cost = cost;
        shape.tiles_per_split = tps;
        shape.n_splits = ns;
      }
    }
  }
  const int n_slots = 2 * shape.n_splits;   // one candidate record per (column range, epilogue half)
  if (int rc = g_tc.ensure((size_t)B * D, (size_t)n_slots * B)) return rc;
  DIF_CUDA_OK(cudaMemsetAsync(g_tc.gmax, 0, 4, st));
  PrepParams pp{};
  pp.src = emb; pp.n = B; pp.D = D; pp.normalize = COSINE ? 1 : 0; pp.split = 1;
  pp.p0 = g_tc.hi; pp.p1 = g_tc.lo;
  if (COSINE) pp.inv = aux; else { pp.sq = aux; pp.gmax = g_tc.gmax; }
  if (int rc = prep_launch(pp, false, st)) return rc;
  CUtensorMap maps[4];
  if (int rc = make_tmap_2d(&maps[0], g_tc.hi, B, D, (uint64_t)D * 4, GEMM_BM, 32, 0)) return rc;
  if (int rc = make_tmap_2d(&maps[1], g_tc.lo, B, D, (uint64_t)D * 4, GEMM_BM, 32, 0)) return rc;
  if (int rc = make_tmap_2d(&maps[2], g_tc.hi, B, D, (uint64_t)D * 4, kMineBN, 32, 0)) return rc;
  if (int rc = make_tmap_2d(&maps[3], g_tc.lo, B, D, (uint64_t)D * 4, kMineBN, 32, 0)) return rc;
  typename MineEpi<COSINE>::Params ep{g_tc.cand, labels, aux, g_tc.gmax, B};
  if (int rc = launch_nt_gemm<0, kMineBN, 1, 0, MineEpi<COSINE>>(maps, shape, ep, sms, st)) return rc;
  bh_rerank_kernel<COSINE><<<(B + kRerankWarps - 1) / kRerankWarps, kRerankWarps * 32, 0, st>>>(
      g_tc.cand, n_slots, emb, labels, aux, g_tc.gmax, B, D, recs, gmax_key, fin);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

template int bh_mine_tensor<true>(const float*, const int32_t*, int, int, BhRec*, float*, unsigned long long*, const BhFinalize&,
                                  cudaStream_t);
template int bh_mine_tensor<false>(const float*, const int32_t*, int, int, BhRec*, float*, unsigned long long*, const BhFinalize&,
                                   cudaStream_t);

}  // namespace dif
