// 1:N gallery search (new API - the reference only has the 1:1 verify of
// deep_insight_face/predictions.py:104-150 and api.py:94-104).
//
// Pipeline of one dif_gallery_search call (all on `stream`, no host synchronisation):
//   1. prep_rows_kernel       queries -> canonical planes (normalised for cosine)            [HBM-bound]
//   2. nt_gemm_rowscan_kernel tensor-core filter: approximate scores of every (query, row) with a
//                             per-query top-k list per gallery split kept in shared memory; the Q x N
//                             score matrix never exists in HBM                               [tensor-bound]
//   3. rerank_kernel          per query: window of candidates that can still be in the true top-k,
//                             canonical fp32 scores for them, final order (score, row asc);
//                             proves from the error bound of pass 2 that nothing outside the
//                             candidate lists can belong to the top-k, else flags the query    [HBM-bound]
//   4. exact_scan_kernel +    flagged queries only (normally none): canonical brute force over the
//      exact_merge_kernel     whole shard
// The result is therefore bit-identical to oracle/dif_oracle.c:dif_or_gallery_search in every
// precision mode; the mode only changes how fast pass 2 runs and how often pass 4 is needed.
#include <algorithm>
#include <new>
#include <vector>

#include "gallery_epi.cuh"
#include "prep_rows.cuh"

namespace dif {

constexpr int kRerankThreads = 128;
constexpr int kRerankCap = 256;     // candidates inside the 2*eps window before a query is flagged
constexpr int kExactChunks = 256;   // row chunks per flagged query in the exact scan
constexpr int kExactMaxFlagged = 4096;  // flagged queries one exact-scan round has workspace for (rounds cover the rest)
constexpr int kExactThreads = 256;

// ------------------------------------------------------------------------------------------
// Selection helpers on unique 64-bit keys held in shared memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t block_max_u64(uint64_t v, uint64_t* red /* [32] */) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const uint64_t w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[warp] = v;
  __syncthreads();
  uint64_t r = red[0];
  for (int i = 1; i < nw; ++i) r = red[i] > r ? red[i] : r;
  return r;
}

struct RerankParams {
  const uint64_t* cand;  // [rows padded][S][kp] approximate keys from pass 2
  int S, kp, k, n_queries;
  const float* q0;
  const float* q1;
  const float* g0;
  const float* g1;
  const float* q_sq;          // canonical |q|^2 per query (sq-L2 bound)
  const unsigned int* gmax;   // orderable max |g|^2
  int D, metric;
  float eps_rel;
  const int64_t* ids;
  int64_t id_base;
  int64_t n_rows;
  float* out_scores;
  int64_t* out_ids;
  int32_t* out_rows;
  int* flagged_count;
  int* flagged_list;
  int force_flag;
};

__device__ __forceinline__ float exact_score_warp(const float* __restrict__ q0, const float* __restrict__ q1,
                                                  const float* __restrict__ g0, const float* __restrict__ g1, int D,
                                                  int metric) {
  // every operand is rebuilt as p0 + p1 (exact) when a lo plane exists.  Unrolled: the loads of a row do not depend
  // on the fma chain, and a candidate row is a cold 2-4 KB read - one load at a time is one DRAM round trip per
  // 128 bytes (this loop was the serial tail of the sharded search, where every rank re-ranks the full query batch).
  float acc = 0.f;
#pragma unroll 8
  for (int d = (int)(threadIdx.x & 31u); d < D; d += 32) {
    const float a = q1 ? __fadd_rn(q0[d], q1[d]) : q0[d];
    const float b = g1 ? __fadd_rn(g0[d], g1[d]) : g0[d];
    if (metric == 1) {
      acc = __fmaf_rn(a, b, acc);
    } else {
      const float t = __fsub_rn(a, b);
      acc = __fmaf_rn(t, t, acc);
    }
  }
  return canon_tree(acc);
}

__device__ __forceinline__ void emit_result(const RerankParams& p, int q, int slot, uint64_t key) {
  const size_t at = (size_t)q * p.k + slot;
  if (key == 0ull) {
    p.out_scores[at] = 0.f;
    if (p.out_ids) p.out_ids[at] = -1;
    if (p.out_rows) p.out_rows[at] = -1;
    return;
  }
  const uint32_t row = key_index(key);
  const float better = key_score(key);
  p.out_scores[at] = p.metric == 1 ? better : -better;
  if (p.out_ids) p.out_ids[at] = p.ids ? p.ids[row] : p.id_base + (int64_t)row;
  if (p.out_rows) p.out_rows[at] = (int32_t)row;
}

// K7: one block per query.
__global__ void __launch_bounds__(kRerankThreads) rerank_kernel(RerankParams p) {
  extern __shared__ uint64_t sm_keys[];  // [S * kp] keys, then the query [D] floats
  __shared__ uint64_t red[32];
  __shared__ uint64_t ekeys[kRerankCap];
  __shared__ uint32_t rrows[kRerankCap];
  __shared__ int n_r;
  __shared__ int flag;

  const int q = blockIdx.x;
  const int n = p.S * p.kp;
  const int tid = threadIdx.x;
  if (tid == 0) {
    n_r = 0;
    flag = p.force_flag;
  }
  const uint64_t* src = p.cand + (size_t)q * n;
  for (int i = tid; i < n; i += blockDim.x) sm_keys[i] = src[i];
  // the exact query (hi + lo) is staged once: every candidate of the window is scored against it
  float* sm_q = reinterpret_cast<float*>(sm_keys + n);
  for (int d = tid; d < p.D; d += blockDim.x)
    sm_q[d] = p.q1 ? __fadd_rn(p.q0[(size_t)q * p.D + d], p.q1[(size_t)q * p.D + d]) : p.q0[(size_t)q * p.D + d];
  __syncthreads();

  // k-th largest approximate key (keys are unique: the row index is part of the key)
  uint64_t prev = ~0ull;
  for (int r = 0; r < p.k; ++r) {
    uint64_t best = 0ull;
    for (int i = tid; i < n; i += blockDim.x) {
      const uint64_t key = sm_keys[i];
      if (key < prev && key > best) best = key;
    }
    prev = block_max_u64(best, red);
    if (prev == 0ull) break;
  }
  const float eps = window_eps(p.metric, p.eps_rel, p.metric == 0 ? p.q_sq[q] : 1.f,
                               p.metric == 0 ? orderable_to_float(*p.gmax) : 1.f);
  const float lim = prev == 0ull ? -INFINITY : key_score(prev) - 2.f * eps;

  // a full list whose minimum reaches the window may hide a better row that was evicted
  for (int s = tid; s < p.S; s += blockDim.x) {
    uint64_t mn = ~0ull;
    for (int i = 0; i < p.kp; ++i) {
      const uint64_t key = sm_keys[s * p.kp + i];
      mn = key < mn ? key : mn;
    }
    if (mn != 0ull && !(key_score(mn) < lim)) flag = 1;
  }
  for (int i = tid; i < n; i += blockDim.x) {
    const uint64_t key = sm_keys[i];
    if (key != 0ull && key_score(key) >= lim) {
      const int at = atomicAdd(&n_r, 1);
      if (at < kRerankCap) rrows[at] = key_index(key);
      else flag = 1;
    }
  }
  __syncthreads();
  const int nr = min(n_r, kRerankCap);

  // canonical scores of the window
  const int warp = tid >> 5, nw = blockDim.x >> 5;
  for (int i = warp; i < nr; i += nw) {
    const uint32_t row = rrows[i];
    const float s = exact_score_warp(sm_q, nullptr, p.g0 + (size_t)row * p.D, p.g1 ? p.g1 + (size_t)row * p.D : nullptr,
                                     p.D, p.metric);
    if ((tid & 31) == 0) ekeys[i] = make_key(p.metric == 1 ? s : -s, row);
  }
  __syncthreads();
  // rank by counting (|window| is small); unique keys -> unique ranks
  for (int i = tid; i < nr; i += blockDim.x) {
    const uint64_t key = ekeys[i];
    int rank = 0;
    for (int j = 0; j < nr; ++j) rank += ekeys[j] > key;
    if (rank < p.k) emit_result(p, q, rank, key);
  }
  for (int i = nr + tid; i < p.k; i += blockDim.x) emit_result(p, q, i, 0ull);
  if (tid == 0 && flag) p.flagged_list[atomicAdd(p.flagged_count, 1)] = q;
}

// Exact scan of flagged queries: work item = (flagged index, row chunk); 8 warps interleave rows,
// each keeps a top-k list in shared memory; the block merges and writes k canonical keys.
__global__ void __launch_bounds__(kExactThreads) exact_scan_kernel(RerankParams p, uint64_t* ex_keys, int flag_base) {
  extern __shared__ float sm_q[];  // [D]
  __shared__ uint64_t lists[kExactThreads / 32][DIF_MAX_TOPK];
  const int n_flag = max(0, min(*p.flagged_count - flag_base, kExactMaxFlagged));   // this round's share
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int64_t chunk_rows = (p.n_rows + kExactChunks - 1) / kExactChunks;
  for (int item = blockIdx.x; item < n_flag * kExactChunks; item += gridDim.x) {
    const int fi = item / kExactChunks, chunk = item - fi * kExactChunks;
    const int q = p.flagged_list[flag_base + fi];
    __syncthreads();
    for (int d = threadIdx.x; d < p.D; d += blockDim.x)
      sm_q[d] = p.q1 ? __fadd_rn(p.q0[(size_t)q * p.D + d], p.q1[(size_t)q * p.D + d]) : p.q0[(size_t)q * p.D + d];
    if (lane < p.k) lists[warp][lane] = 0ull;
    __syncthreads();
    uint64_t thr_key = 0ull;  // minimum of this warp's list (0 while it is not full)
    int min_slot = 0;
    const int64_t r0 = chunk * chunk_rows, r1 = min(r0 + chunk_rows, p.n_rows);
    for (int64_t r = r0 + warp; r < r1; r += nw) {
      const float s = exact_score_warp(sm_q, nullptr, p.g0 + (size_t)r * p.D, p.g1 ? p.g1 + (size_t)r * p.D : nullptr,
                                       p.D, p.metric);
      const uint64_t key = make_key(p.metric == 1 ? s : -s, (uint32_t)r);
      if (key > thr_key) {  // warp-uniform
        if (lane == 0) lists[warp][min_slot] = key;
        __syncwarp();
        uint64_t mn = lists[warp][0];
        int ms = 0;
        for (int i = 1; i < p.k; ++i) {
          const uint64_t k2 = lists[warp][i];
          if (k2 < mn) {
            mn = k2;
            ms = i;
          }
        }
        thr_key = mn;
        min_slot = ms;
        __syncwarp();
      }
    }
    uint64_t* dst = ex_keys + ((size_t)fi * kExactChunks + chunk) * p.k;
    if (threadIdx.x < p.k) dst[threadIdx.x] = 0ull;
    __syncthreads();
    // merge nw lists by rank counting
    uint64_t* all = &lists[0][0];
    if (threadIdx.x < nw * DIF_MAX_TOPK) {
      const int w = threadIdx.x / DIF_MAX_TOPK, i = threadIdx.x - w * DIF_MAX_TOPK;
      const uint64_t key = i < p.k ? all[threadIdx.x] : 0ull;
      if (key != 0ull) {
        int rank = 0;
        for (int w2 = 0; w2 < nw; ++w2)
          for (int j = 0; j < p.k; ++j) rank += lists[w2][j] > key;
        if (rank < p.k) dst[rank] = key;
      }
    }
  }
}

__global__ void __launch_bounds__(kRerankThreads) exact_merge_kernel(RerankParams p, const uint64_t* ex_keys,
                                                                     int flag_base) {
  extern __shared__ uint64_t sm_keys[];
  const int n_flag = max(0, min(*p.flagged_count - flag_base, kExactMaxFlagged));
  const int n = kExactChunks * p.k;
  for (int fi = blockIdx.x; fi < n_flag; fi += gridDim.x) {
    const int q = p.flagged_list[flag_base + fi];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm_keys[i] = ex_keys[(size_t)fi * n + i];
    __syncthreads();
    int n_valid = 0;
    for (int i = 0; i < n; ++i) n_valid += sm_keys[i] != 0ull;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint64_t key = sm_keys[i];
      if (key == 0ull) continue;
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += sm_keys[j] > key;
      if (rank < p.k) emit_result(p, q, rank, key);
    }
    for (int i = n_valid + threadIdx.x; i < p.k; i += blockDim.x) emit_result(p, q, i, 0ull);
  }
}

// Cross-shard merge: candidates carry (score, global row, id); order (better score, global row asc).
__global__ void __launch_bounds__(128) topk_merge_kernel(const float* scores, const int64_t* grows, const int64_t* ids,
                                                         int world, int n_queries, int k, int metric,
                                                         float* out_scores, int64_t* out_grows, int64_t* out_ids) {
  extern __shared__ uint8_t sm_raw[];
  const int n = world * k;
  uint32_t* so = reinterpret_cast<uint32_t*>(sm_raw);                 // orderable "better" score
  int64_t* sr = reinterpret_cast<int64_t*>(sm_raw + ((n * 4 + 7) & ~7));  // global row, -1 = empty
  for (int q = blockIdx.x; q < n_queries; q += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int w = i / k, j = i - w * k;
      const size_t at = ((size_t)w * n_queries + q) * k + j;
      const float s = scores[at];
      so[i] = float_orderable(metric == 1 ? s : -s);
      sr[i] = grows[at];
    }
    __syncthreads();
    int n_valid = 0;
    for (int i = 0; i < n; ++i) n_valid += sr[i] >= 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (sr[i] < 0) continue;
      int rank = 0;
      for (int j = 0; j < n; ++j)
        rank += sr[j] >= 0 && (so[j] > so[i] || (so[j] == so[i] && (sr[j] < sr[i] || (sr[j] == sr[i] && j < i))));
      if (rank < k) {
        const int w = i / k, jj = i - w * k;
        const size_t at = ((size_t)w * n_queries + q) * k + jj;
        out_scores[(size_t)q * k + rank] = scores[at];
        out_grows[(size_t)q * k + rank] = sr[i];
        if (out_ids) out_ids[(size_t)q * k + rank] = ids ? ids[at] : sr[i];
      }
    }
    for (int i = n_valid + threadIdx.x; i < k; i += blockDim.x) {
      out_scores[(size_t)q * k + i] = 0.f;
      out_grows[(size_t)q * k + i] = -1;
      if (out_ids) out_ids[(size_t)q * k + i] = -1;
    }
  }
}

__global__ void gather_rows_kernel(const float* p0, const float* p1, int64_t row0, int64_t n, int D, float* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * D; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t at = row0 * D + i;
    out[i] = p1 ? __fadd_rn(p0[at], p1[at]) : p0[at];
  }
}

}  // namespace dif

// ==========================================================================================
// host side
// ==========================================================================================
using namespace dif;

namespace dif {
__global__ void iota_ids_kernel(int64_t* ids, int64_t n, int64_t base) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) ids[i] = base + i;
}

// Order-preserving delete, one chunk: every kept row r of [r_begin, r_end) goes to stage row
// r - (#removed rows below r) - new_begin.  One warp per row, rows are `words` 4-byte words.
__global__ void __launch_bounds__(256) compact_rows_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ stage,
                                                           int words, const int64_t* __restrict__ removed, int64_t n_removed,
                                                           int64_t r_begin, int64_t r_end, int64_t new_begin) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), stride = (int64_t)gridDim.x * 8;
  for (int64_t r = r_begin + warp0; r < r_end; r += stride) {
    int64_t lo = 0, hi = n_removed;   // first index with removed[idx] >= r
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (removed[mid] < r) lo = mid + 1;
      else hi = mid;
    }
    if (lo < n_removed && removed[lo] == r) continue;
    const uint32_t* s = src + (size_t)r * words;
    uint32_t* d = stage + (size_t)(r - lo - new_begin) * words;
    for (int w = lane; w < words; w += 32) d[w] = s[w];
  }
}
}  // namespace dif

#include "gallery_state.cuh"

namespace {

template <class T>
int dev_alloc(T** p, size_t n) {
  DIF_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
  return DIF_OK;
}

int ensure_query_ws(dif_gallery* g, int Q, int S, int kp) {
  const int q_pad = (int)round_up(Q, GEMM_BM * 2);
  if (q_pad > g->q_cap) {
    cudaFree(g->q0); cudaFree(g->q1); cudaFree(g->qb); cudaFree(g->qb1); cudaFree(g->qsq); cudaFree(g->flagged); cudaFree(g->bound);
    g->q0 = g->q1 = nullptr; g->qb = g->qb1 = nullptr; g->qsq = nullptr; g->flagged = nullptr; g->bound = nullptr;
    g->q_cap = 0;
    if (int rc = dev_alloc(&g->q0, (size_t)q_pad * g->D)) return rc;
    if (g->precision == DIF_PREC_TF32X3)
      if (int rc = dev_alloc(&g->q1, (size_t)q_pad * g->D)) return rc;
    if (g->precision == DIF_PREC_BF16 || g->precision == DIF_PREC_BF16X3)
      if (int rc = dev_alloc(&g->qb, (size_t)q_pad * g->D)) return rc;
    if (g->precision == DIF_PREC_BF16X3)
      if (int rc = dev_alloc(&g->qb1, (size_t)q_pad * g->D)) return rc;
    if (int rc = dev_alloc(&g->qsq, (size_t)q_pad)) return rc;
    if (int rc = dev_alloc(&g->flagged, (size_t)q_pad + 1)) return rc;
    if (int rc = dev_alloc(&g->bound, (size_t)q_pad)) return rc;
    DIF_CUDA_OK(cudaMemset(g->qsq, 0, (size_t)q_pad * 4));
    g->q_cap = q_pad;
  }
  const size_t need_max = (size_t)g->q_cap * S;
  if (need_max > g->maxima_elems) {
    cudaFree(g->maxima);
    g->maxima = nullptr;
    g->maxima_elems = 0;
    if (int rc = dev_alloc(&g->maxima, need_max)) return rc;
    g->maxima_elems = need_max;
  }
  const size_t need = (size_t)q_pad * S * kp;
  if (need > g->cand_elems) {
    cudaFree(g->cand);
    g->cand = nullptr;
    g->cand_elems = 0;
    if (int rc = dev_alloc(&g->cand, need)) return rc;
    g->cand_elems = need;
  }
  return DIF_OK;
}

// resident-A schedule whenever the query block (all of K) fits beside a >= 3-stage B ring
bool ares_fits(int precision, int k_chunks, int kp) {
  GemmSmemPlan plan;
  const int epi = TopkEpi<1>::smem_bytes_kp(kp);
  bool ok = false;
  if (precision == DIF_PREC_BF16) ok = plan_gemm_smem<1, kGalBN, 2, 1>(k_chunks, epi, &plan);
  else if (precision == DIF_PREC_TF32X1) ok = plan_gemm_smem<2, kGalBN, 2, 1>(k_chunks, epi, &plan);
  return ok && plan.stages >= 3;
}

}  // namespace

extern "C" {

dif_gallery_t* dif_gallery_create(int device, int64_t capacity_rows, int dim, int metric, int precision) {
  if (dif_init(device) != DIF_OK) return nullptr;
  if (capacity_rows <= 0 || capacity_rows >= (int64_t)0x7FFFFF00 || dim < 32 || dim % 4 != 0 || dim > 8192 ||
      (metric != DIF_METRIC_SQL2 && metric != DIF_METRIC_COSINE) || precision < 0 || precision > 3 ||
      ((precision == DIF_PREC_BF16 || precision == DIF_PREC_BF16X3) && (dim % 8 != 0 || dim < 64))) {
    set_error("dif_gallery_create: invalid argument (capacity %lld, dim %d (multiple of 4 and >= 32; of 8 and >= 64 for "
              "bf16), metric %d, precision %d)", (long long)capacity_rows, dim, metric, precision);
    return nullptr;
  }
  dif_gallery* g = new (std::nothrow) dif_gallery();
  if (!g) return nullptr;
  g->device = device;
  g->capacity = capacity_rows;
  g->D = dim;
  g->metric = metric;
  g->precision = precision;
  const size_t elems = (size_t)capacity_rows * dim;
  bool ok = cudaMalloc((void**)&g->g0, elems * 4) == cudaSuccess;
  if (ok && precision == DIF_PREC_TF32X3) ok = cudaMalloc((void**)&g->g1, elems * 4) == cudaSuccess;
  if (ok && (precision == DIF_PREC_BF16 || precision == DIF_PREC_BF16X3)) ok = cudaMalloc((void**)&g->gb, elems * 2) == cudaSuccess;
  if (ok && precision == DIF_PREC_BF16X3) ok = cudaMalloc((void**)&g->gb1, elems * 2) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&g->gsq, ((size_t)capacity_rows + kGalBN) * 4) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&g->gmax, 4) == cudaSuccess;
  ok = ok && cudaMemset(g->gsq, 0, ((size_t)capacity_rows + kGalBN) * 4) == cudaSuccess;
  ok = ok && cudaMemset(g->gmax, 0, 4) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&g->own_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreate(&g->ev0) == cudaSuccess && cudaEventCreate(&g->ev1) == cudaSuccess;
  for (cudaEvent_t& e : g->ev_phase) ok = ok && cudaEventCreate(&e) == cudaSuccess;
  if (!ok) {
    set_error("dif_gallery_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    dif_gallery_destroy(g);
    return nullptr;
  }
  return g;
}

void dif_gallery_destroy(dif_gallery_t* g) {
  if (!g) return;
  cudaFree(g->g0); cudaFree(g->g1); cudaFree(g->gb); cudaFree(g->gb1); cudaFree(g->gsq); cudaFree(g->gmax); cudaFree(g->ids);
  cudaFree(g->q0); cudaFree(g->q1); cudaFree(g->qb); cudaFree(g->qb1); cudaFree(g->qsq); cudaFree(g->cand); cudaFree(g->flagged);
  cudaFree(g->bound); cudaFree(g->maxima); cudaFree(g->ex_keys); cudaFree(g->d_stage);
  dif::shard_state_destroy(g->shard);
  if (g->h_pin) cudaFreeHost(g->h_pin);
  if (g->own_stream) cudaStreamDestroy(g->own_stream);
  if (g->ev0) cudaEventDestroy(g->ev0);
  if (g->ev1) cudaEventDestroy(g->ev1);
  for (cudaEvent_t e : g->ev_phase)
    if (e) cudaEventDestroy(e);
  delete g;
}

int dif_gallery_set_option(dif_gallery_t* g, const char* name, int value) {
  DIF_REQUIRE(g && name, DIF_ERR_INVALID, "dif_gallery_set_option: null argument");
  if (!strcmp(name, "gemm_ctas")) {
    DIF_REQUIRE(value == 1 || value == 2, DIF_ERR_INVALID, "gemm_ctas must be 1 or 2");
    g->opt_ctas = value;
  } else if (!strcmp(name, "force_fallback")) {
    g->opt_force_fallback = value != 0;
  } else if (!strcmp(name, "resident_queries")) {
    g->opt_resident = value < 0 ? -1 : (value != 0);
  } else if (!strcmp(name, "l2_prefetch")) {
    g->opt_l2_prefetch = value != 0;
  } else if (!strcmp(name, "splits")) {
    g->opt_splits = value < 0 ? 0 : value;
  } else {
    DIF_REQUIRE(false, DIF_ERR_INVALID, "unknown option '%s'", name);
  }
  return DIF_OK;
}

static int gallery_append(dif_gallery_t* g, const float* rows, bool synth, uint64_t seed, int64_t row0,
                          const int64_t* ids, int64_t n, cudaStream_t st) {
  DIF_REQUIRE(g, DIF_ERR_INVALID, "null gallery");
  DIF_REQUIRE(n >= 0 && g->size + n <= g->capacity, DIF_ERR_CAPACITY, "gallery full: %lld + %lld > capacity %lld",
              (long long)g->size, (long long)n, (long long)g->capacity);
  if (n == 0) return DIF_OK;
  DIF_REQUIRE(synth || rows, DIF_ERR_INVALID, "rows is NULL");
  PrepParams pp{};
  pp.src = rows;
  pp.seed = seed;
  pp.row0 = row0;
  pp.n = n;
  pp.D = g->D;
  pp.normalize = g->metric == DIF_METRIC_COSINE;
  pp.split = g->precision == DIF_PREC_TF32X3;
  pp.p0 = g->g0 + (size_t)g->size * g->D;
  pp.p1 = g->g1 ? g->g1 + (size_t)g->size * g->D : nullptr;
  pp.pb = g->gb ? g->gb + (size_t)g->size * g->D : nullptr;
  pp.pb1 = g->gb1 ? g->gb1 + (size_t)g->size * g->D : nullptr;
  pp.sq = g->gsq + g->size;
  pp.gmax = g->gmax;
  if (int rc = prep_launch(pp, synth, st)) return rc;
  if (ids || g->has_ids) {
    if (!g->ids) DIF_CUDA_OK(cudaMalloc((void**)&g->ids, (size_t)g->capacity * 8));
    if (!g->has_ids && g->size > 0) {
      // earlier rows were enrolled without ids: their default ids (id_base + row) become explicit now
      iota_ids_kernel<<<(unsigned)std::min<int64_t>((g->size + 255) / 256, 148 * 8), 256, 0, st>>>(g->ids, g->size, g->id_base);
      DIF_LAUNCH_OK();
    }
    if (ids) {
      DIF_CUDA_OK(cudaMemcpyAsync(g->ids + g->size, ids, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
      g->user_ids = true;
    } else {
      DIF_REQUIRE(!g->user_ids, DIF_ERR_STATE, "this gallery was built with explicit ids; ids must be given on every add");
      // default ids were frozen by a remove: rows enrolled without ids continue the id_base + n sequence, n = rows
      // ever enrolled, so a new row never reuses the id of a surviving one
      iota_ids_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(g->ids + g->size, n,
                                                                                           g->id_base + g->enrolled);
      DIF_LAUNCH_OK();
    }
    g->has_ids = true;
  }
  g->enrolled += n;
  g->size += n;
  return DIF_OK;
}

int dif_gallery_add(dif_gallery_t* g, const float* rows, const int64_t* ids, int64_t n, void* stream) {
  return gallery_append(g, rows, false, 0, 0, ids, n, static_cast<cudaStream_t>(stream));
}

int dif_gallery_fill_synth(dif_gallery_t* g, uint64_t seed, int64_t row0, int64_t n, void* stream) {
  return gallery_append(g, nullptr, true, seed, row0, nullptr, n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
namespace dif {
int gallery_ensure_stage(dif_gallery* g, size_t bytes) {
  if (bytes > g->h_pin_bytes) {
    if (g->h_pin) cudaFreeHost(g->h_pin);
    g->h_pin = nullptr;
    g->h_pin_bytes = 0;
    DIF_CUDA_OK(cudaMallocHost(&g->h_pin, bytes));
    g->h_pin_bytes = bytes;
  }
  if (bytes > g->d_stage_bytes) {
    cudaFree(g->d_stage);
    g->d_stage = nullptr;
    g->d_stage_bytes = 0;
    DIF_CUDA_OK(cudaMalloc(&g->d_stage, bytes));
    g->d_stage_bytes = bytes;
  }
  return DIF_OK;
}
}  // namespace dif
static int ensure_stage(dif_gallery_t* g, size_t bytes) { return dif::gallery_ensure_stage(g, bytes); }
extern "C" {

int dif_gallery_add_host(dif_gallery_t* g, const float* rows_host, const int64_t* ids_host, int64_t n) {
  DIF_REQUIRE(g && (rows_host || n == 0), DIF_ERR_INVALID, "dif_gallery_add_host: null argument");
  const int64_t chunk = std::max<int64_t>(1, (int64_t)(64u << 20) / (g->D * 4));
  for (int64_t done = 0; done < n; done += chunk) {
    const int64_t m = std::min(chunk, n - done);
    const size_t rb = (size_t)m * g->D * 4, ib = ids_host ? (size_t)m * 8 : 0;
    if (int rc = ensure_stage(g, rb + ib)) return rc;
    memcpy(g->h_pin, rows_host + (size_t)done * g->D, rb);
    if (ids_host) memcpy((char*)g->h_pin + rb, ids_host + done, ib);
    DIF_CUDA_OK(cudaMemcpyAsync(g->d_stage, g->h_pin, rb + ib, cudaMemcpyHostToDevice, g->own_stream));
    if (int rc = gallery_append(g, (const float*)g->d_stage, false, 0, 0,
                                ids_host ? (const int64_t*)((char*)g->d_stage + rb) : nullptr, m, g->own_stream))
      return rc;
    DIF_CUDA_OK(cudaStreamSynchronize(g->own_stream));
  }
  return DIF_OK;
}

int dif_gallery_remove(dif_gallery_t* g, const int64_t* rows_host, int64_t n, void* stream) {
  DIF_REQUIRE(g && (rows_host || n == 0) && n >= 0, DIF_ERR_INVALID, "dif_gallery_remove: invalid argument");
  if (n == 0) return DIF_OK;
  for (int64_t i = 0; i < n; ++i)
    DIF_REQUIRE(rows_host[i] >= 0 && rows_host[i] < g->size && (i == 0 || rows_host[i] > rows_host[i - 1]), DIF_ERR_INVALID,
                "dif_gallery_remove: rows must be strictly ascending and below the gallery size %lld (entry %lld = %lld)",
                (long long)g->size, (long long)i, (long long)rows_host[i]);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the sorted row list lives behind the compaction chunk in the handle's device staging block: no allocation per call
  const int64_t chunk = std::max<int64_t>(1, (int64_t)(32u << 20) / (g->D * 4));
  const size_t list_off = ((size_t)chunk * g->D * 4 + 255) & ~(size_t)255;
  if (int rc = ensure_stage(g, list_off + (size_t)n * 8)) return rc;
  int64_t* removed = reinterpret_cast<int64_t*>(static_cast<char*>(g->d_stage) + list_off);
  DIF_CUDA_OK(cudaMemcpyAsync(removed, rows_host, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  if (!g->has_ids) {   // ids were the row numbers: make them explicit before rows move
    if (!g->ids) DIF_CUDA_OK(cudaMalloc((void**)&g->ids, (size_t)g->capacity * 8));
    iota_ids_kernel<<<(unsigned)std::min<int64_t>((g->size + 255) / 256, 148 * 8), 256, 0, st>>>(g->ids, g->size, g->id_base);
    DIF_LAUNCH_OK();
    g->has_ids = true;
  }
  const int64_t first = rows_host[0];
  struct Arr {
    void* base;
    int words;   // 4-byte words per row
  } arrs[6] = {{g->g0, g->D}, {g->g1, g->D}, {g->gb, g->D / 2}, {g->gb1, g->D / 2}, {g->gsq, 1}, {g->ids, 2}};
  for (const Arr& a : arrs) {
    if (!a.base) continue;
    for (int64_t r0 = first; r0 < g->size; r0 += chunk) {
      const int64_t r1 = std::min(g->size, r0 + chunk);
      const int64_t before = std::lower_bound(rows_host, rows_host + n, r0) - rows_host;
      const int64_t inside = (std::lower_bound(rows_host, rows_host + n, r1) - rows_host) - before;
      const int64_t kept = (r1 - r0) - inside, new_begin = r0 - before;
      if (kept == 0) continue;
      compact_rows_kernel<<<(unsigned)std::min<int64_t>((r1 - r0 + 7) / 8, 148 * 16), 256, 0, st>>>(
          (const uint32_t*)a.base, (uint32_t*)g->d_stage, a.words, removed, n, r0, r1, new_begin);
      DIF_LAUNCH_OK();
      // the destination ends at or before r1: it may overlap this chunk's (already staged) source, never a later one
      DIF_CUDA_OK(cudaMemcpyAsync((char*)a.base + (size_t)new_begin * a.words * 4, g->d_stage, (size_t)kept * a.words * 4,
                                  cudaMemcpyDeviceToDevice, st));
    }
  }
  DIF_CUDA_OK(cudaMemsetAsync(g->gsq + (g->size - n), 0, (size_t)n * 4, st));
  DIF_CUDA_OK(cudaStreamSynchronize(st));   // rows_host (pageable) and the staging block may be reused on return
  g->size -= n;
  return DIF_OK;
}

int dif_gallery_get_ids(dif_gallery_t* g, int64_t row0, int64_t n, int64_t* out, void* stream) {
  DIF_REQUIRE(g && out && row0 >= 0 && n >= 0 && row0 + n <= g->size, DIF_ERR_INVALID, "dif_gallery_get_ids: range");
  if (n == 0) return DIF_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g->has_ids) {
    DIF_CUDA_OK(cudaMemcpyAsync(out, g->ids + row0, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
  } else {
    iota_ids_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(out, n, g->id_base + row0);
    DIF_LAUNCH_OK();
  }
  return DIF_OK;
}

int dif_gallery_set_id_base(dif_gallery_t* g, int64_t id_base) {
  DIF_REQUIRE(g, DIF_ERR_INVALID, "null gallery");
  g->id_base = id_base;
  return DIF_OK;
}

int64_t dif_gallery_size(const dif_gallery_t* g) { return g ? g->size : -1; }

int dif_gallery_reset(dif_gallery_t* g) {
  DIF_REQUIRE(g, DIF_ERR_INVALID, "null gallery");
  g->size = 0;
  g->enrolled = 0;
  g->has_ids = false;
  g->user_ids = false;
  DIF_CUDA_OK(cudaMemset(g->gsq, 0, ((size_t)g->capacity + kGalBN) * 4));
  DIF_CUDA_OK(cudaMemset(g->gmax, 0, 4));
  return DIF_OK;
}

int dif_gallery_search(dif_gallery_t* g, const float* queries, int n_queries, int k, float* scores, int64_t* ids,
                       int32_t* rows, void* stream) {
  DIF_REQUIRE(g && queries && scores && ids, DIF_ERR_INVALID, "dif_gallery_search: null argument");
  return dif::gallery_search_impl(g, queries, n_queries, k, scores, ids, rows, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

int dif::gallery_search_impl(dif_gallery* g, const float* queries, int n_queries, int k, float* scores, int64_t* ids,
                             int32_t* rows, cudaStream_t st) {
  DIF_REQUIRE(g && queries && scores, DIF_ERR_INVALID, "gallery search: null argument");
  DIF_REQUIRE(n_queries > 0 && k >= 1 && k <= DIF_MAX_TOPK, DIF_ERR_INVALID,
              "dif_gallery_search: n_queries %d, k %d (1..%d)", n_queries, k, DIF_MAX_TOPK);
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  const int D = g->D, kp = k;
  const int ctas = g->opt_ctas;
  const int64_t launches0 = dif_launch_count();

  // work decomposition: items = splits x query blocks, a whole number of items per persistent unit
  GemmShape shape{};
  shape.m_blocks = (n_queries + GEMM_BM * ctas - 1) / (GEMM_BM * ctas);
  shape.n_tiles = (int)((g->size + kGalBN - 1) / kGalBN);
  const bool bf = g->precision == DIF_PREC_BF16 || g->precision == DIF_PREC_BF16X3;
  const int chunk = bf ? 64 : 32;
  shape.k_chunks = (D + chunk - 1) / chunk;
  const int units = std::max(1, device_sm_count() / ctas);
  int splits = 1;
  if (shape.n_tiles > 0) {
    const int rounds = std::max(1, std::min(8, (shape.n_tiles * shape.m_blocks) / (units * 8)));
    splits = std::max(1, (units * rounds + shape.m_blocks - 1) / shape.m_blocks);
    // Each split keeps only k candidates: a split that scans a large share of the shard fills its list with
    // rows inside the re-rank window and flags the query for the exact scan.  >= 32 splits keep every list's
    // k-th best well below the window for the wide (bf16 / TF32) filters even at 10^7 rows per shard.
    splits = std::max(splits, 32);
    splits = std::min(splits, shape.n_tiles);
    splits = std::min(splits, 512);
    if (g->opt_splits > 0) splits = std::min(g->opt_splits, shape.n_tiles);
  }
  shape.l2_prefetch = g->opt_l2_prefetch;
  shape.n_splits = splits;
  shape.tiles_per_split = shape.n_tiles > 0 ? (shape.n_tiles + splits - 1) / splits : 0;
  if (shape.tiles_per_split > 0) shape.n_splits = splits = (shape.n_tiles + shape.tiles_per_split - 1) / shape.tiles_per_split;

  if (int rc = ensure_query_ws(g, n_queries, splits, kp)) return rc;

  DIF_CUDA_OK(cudaEventRecord(g->ev_phase[0], st));
  g->phase_sharded = false;
  // 1. queries -> canonical planes
  PrepParams pp{};
  pp.src = queries;
  pp.n = n_queries;
  pp.D = D;
  pp.normalize = g->metric == DIF_METRIC_COSINE;
  pp.split = g->precision == DIF_PREC_TF32X3;
  pp.p0 = g->q0;
  pp.p1 = g->q1;
  pp.pb = g->qb;
  pp.pb1 = g->qb1;
  pp.sq = g->qsq;
  pp.gmax = nullptr;
  if (int rc = prep_launch(pp, false, st)) return rc;
  DIF_CUDA_OK(cudaMemsetAsync(g->flagged, 0, sizeof(int), st));

  // 2. tensor-core filter
  if (shape.n_tiles > 0) {
    CUtensorMap maps[4];
    const void* a0 = bf ? (const void*)g->qb : (const void*)g->q0;
    const void* b0 = bf ? (const void*)g->gb : (const void*)g->g0;
    const int esz = bf ? 2 : 4;
    const uint32_t bcols = (uint32_t)(128 / esz);
    if (int rc = make_tmap_2d(&maps[0], a0, n_queries, D, (uint64_t)D * esz, GEMM_BM, bcols, bf)) return rc;
    if (int rc = make_tmap_2d(&maps[2], b0, g->size, D, (uint64_t)D * esz, kGalBN / ctas, bcols, bf)) return rc;
    maps[1] = maps[0];
    maps[3] = maps[2];
    if (g->precision == DIF_PREC_TF32X3) {
      if (int rc = make_tmap_2d(&maps[1], g->q1, n_queries, D, (uint64_t)D * 4, GEMM_BM, bcols, 0)) return rc;
      if (int rc = make_tmap_2d(&maps[3], g->g1, g->size, D, (uint64_t)D * 4, kGalBN / ctas, bcols, 0)) return rc;
    } else if (g->precision == DIF_PREC_BF16X3) {
      if (int rc = make_tmap_2d(&maps[1], g->qb1, n_queries, D, (uint64_t)D * 2, GEMM_BM, bcols, 1)) return rc;
      if (int rc = make_tmap_2d(&maps[3], g->gb1, g->size, D, (uint64_t)D * 2, kGalBN / ctas, bcols, 1)) return rc;
    }
    DIF_CUDA_OK(cudaMemsetAsync(g->bound, 0, (size_t)g->q_cap * sizeof(unsigned int), st));
    DIF_CUDA_OK(cudaMemsetAsync(g->maxima, 0, (size_t)g->q_cap * splits * sizeof(unsigned int), st));
    TopkEpi<1>::Params ep{g->cand, g->gsq, g->bound, g->maxima, g->q_cap, k, g->qsq, g->gmax, (int)g->size, splits, kp,
                          mode_eps(g->precision)};
    const int ares = ctas == 2 && g->opt_resident != 0 && ares_fits(g->precision, shape.k_chunks, kp);
    g->stats[4] = ares;
    DIF_CUDA_OK(cudaEventRecord(g->ev0, st));
    const int rc = g->precision == DIF_PREC_TF32X3 ? launch_search_tf32x3(g->metric, ctas, ares, maps, shape, ep, units, st)
                   : g->precision == DIF_PREC_BF16 ? launch_search_bf16(g->metric, ctas, ares, maps, shape, ep, units, st)
                   : g->precision == DIF_PREC_BF16X3 ? launch_search_bf16x3(g->metric, ctas, 0, maps, shape, ep, units, st)
                                                   : launch_search_tf32x1(g->metric, ctas, ares, maps, shape, ep, units, st);
    if (rc) return rc;
    DIF_CUDA_OK(cudaEventRecord(g->ev1, st));
  } else {
    DIF_CUDA_OK(cudaMemsetAsync(g->cand, 0, (size_t)round_up(n_queries, GEMM_BM * 2) * splits * kp * 8, st));
  }

  // 3. window + canonical re-rank
  RerankParams rp{};
  rp.cand = g->cand;
  rp.S = splits;
  rp.kp = kp;
  rp.k = k;
  rp.n_queries = n_queries;
  rp.q0 = g->q0;
  rp.q1 = g->q1;
  rp.g0 = g->g0;
  rp.g1 = g->g1;
  rp.q_sq = g->qsq;
  rp.gmax = g->gmax;
  rp.D = D;
  rp.metric = g->metric;
  rp.eps_rel = mode_eps(g->precision);
  rp.ids = g->has_ids ? g->ids : nullptr;
  rp.id_base = g->id_base;
  rp.n_rows = g->size;
  rp.out_scores = scores;
  rp.out_ids = ids;
  rp.out_rows = rows;
  rp.flagged_count = g->flagged;
  rp.flagged_list = g->flagged + 1;
  rp.force_flag = g->opt_force_fallback;
  const size_t rr_smem = (size_t)splits * kp * 8 + (size_t)D * 4;
  DIF_REQUIRE(rr_smem <= 160 * 1024, DIF_ERR_CAPACITY, "candidate lists too large for the re-rank kernel");
  if (rr_smem > 40 * 1024)
    DIF_CUDA_OK(cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rr_smem));
  rerank_kernel<<<n_queries, kRerankThreads, rr_smem, st>>>(rp);
  DIF_LAUNCH_OK();
  DIF_CUDA_OK(cudaEventRecord(g->ev_phase[1], st));

  // 4. exact path for flagged queries (grids sized for the hardware, work read from the device counter)
  if (g->size > 0) {
    const size_t need = (size_t)std::min(g->q_cap, kExactMaxFlagged) * kExactChunks * k;
    if (need > g->ex_elems) {
      // allocated lazily but before it can be needed: the flagged count is only known on the device
      cudaFree(g->ex_keys);
      g->ex_keys = nullptr;
      g->ex_elems = 0;
      if (int rc = dev_alloc(&g->ex_keys, need)) return rc;
      g->ex_elems = need;
    }
    const size_t em_smem = (size_t)kExactChunks * k * 8;
    if (em_smem > 40 * 1024)
      DIF_CUDA_OK(cudaFuncSetAttribute(exact_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)em_smem));
    // the flagged count lives on the device: launch enough rounds of (scan, merge) to cover every query; a round
    // whose share is empty returns at once
    for (int base = 0; base < n_queries; base += kExactMaxFlagged) {
      exact_scan_kernel<<<device_sm_count() * 4, kExactThreads, (size_t)D * 4, st>>>(rp, g->ex_keys, base);
      DIF_LAUNCH_OK();
      exact_merge_kernel<<<device_sm_count() * 2, kRerankThreads, em_smem, st>>>(rp, g->ex_keys, base);
      DIF_LAUNCH_OK();
    }
  }
  DIF_CUDA_OK(cudaEventRecord(g->ev_phase[2], st));
  g->stats[1] = dif_launch_count() - launches0;
  g->stats[2] = splits;
  g->stats[3] = kp;
  return DIF_OK;
}

extern "C" {

int dif_gallery_search_host(dif_gallery_t* g, const float* queries_host, int n_queries, int k, float* scores_host,
                            int64_t* ids_host, int32_t* rows_host) {
  DIF_REQUIRE(g && queries_host && scores_host && ids_host, DIF_ERR_INVALID, "dif_gallery_search_host: null argument");
  DIF_REQUIRE(n_queries > 0 && k >= 1 && k <= DIF_MAX_TOPK, DIF_ERR_INVALID, "n_queries %d, k %d", n_queries, k);
  const size_t qb = (size_t)n_queries * g->D * 4;
  const size_t nk = (size_t)n_queries * k;
  const size_t ob = nk * (4 + 8 + 4);
  const size_t qb_al = (qb + 255) & ~(size_t)255;
  if (int rc = ensure_stage(g, qb_al + ob + 256)) return rc;
  char* hp = (char*)g->h_pin;
  char* dp = (char*)g->d_stage;
  cudaStream_t st = g->own_stream;
  cudaPointerAttributes attr{};
  const bool pinned = cudaPointerGetAttributes(&attr, queries_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  if (!pinned) cudaGetLastError();   // (older drivers report pageable memory as an error)
  if (pinned) {
    // page-locked caller buffer (cudaHostAlloc / cudaHostRegister / torch pin_memory): DMA straight from it
    DIF_CUDA_OK(cudaMemcpyAsync(dp, queries_host, qb, cudaMemcpyHostToDevice, st));
  } else {
    // staged in 1 MB pieces so the DMA of piece i overlaps the host memcpy of piece i + 1
    const size_t piece = (size_t)1 << 20;
    for (size_t off = 0; off < qb; off += piece) {
      const size_t n = std::min(piece, qb - off);
      memcpy(hp + off, (const char*)queries_host + off, n);
      DIF_CUDA_OK(cudaMemcpyAsync(dp + off, hp + off, n, cudaMemcpyHostToDevice, st));
    }
  }
  int64_t* d_ids = (int64_t*)(dp + qb_al);
  float* d_scores = (float*)(dp + qb_al + nk * 8);
  int32_t* d_rows = (int32_t*)(dp + qb_al + nk * 12);
  if (int rc = dif_gallery_search(g, (const float*)dp, n_queries, k, d_scores, d_ids, d_rows, st)) return rc;
  DIF_CUDA_OK(cudaMemcpyAsync(hp + qb_al, dp + qb_al, ob, cudaMemcpyDeviceToHost, st));
  DIF_CUDA_OK(cudaStreamSynchronize(st));

  memcpy(ids_host, hp + qb_al, nk * 8);
  memcpy(scores_host, hp + qb_al + nk * 8, nk * 4);
  if (rows_host) memcpy(rows_host, hp + qb_al + nk * 12, nk * 4);
  return DIF_OK;
}

int dif_gallery_last_stats(const dif_gallery_t* g, int64_t out[6]) {
  DIF_REQUIRE(g && out, DIF_ERR_INVALID, "null argument");
  int n_flag = 0;
  if (g->flagged) DIF_CUDA_OK(cudaMemcpy(&n_flag, g->flagged, sizeof(int), cudaMemcpyDeviceToHost));
  out[0] = n_flag;
  out[1] = g->stats[1];
  out[2] = g->stats[2];
  out[3] = g->stats[3];
  out[4] = g->stats[4];
  out[5] = 0;   // (reserved; every flagged query is verified by the exact scan rounds)
  return DIF_OK;
}

int dif_gallery_last_kernel_ms(dif_gallery_t* g, float* ms) {
  DIF_REQUIRE(g && ms, DIF_ERR_INVALID, "null argument");
  DIF_CUDA_OK(cudaEventElapsedTime(ms, g->ev0, g->ev1));
  return DIF_OK;
}

int dif_gallery_last_phase_ms(dif_gallery_t* g, float out[5]) {
  DIF_REQUIRE(g && out, DIF_ERR_INVALID, "null argument");
  for (int i = 0; i < 5; ++i) out[i] = 0.f;
  if (g->size > 0) {
    DIF_CUDA_OK(cudaEventElapsedTime(&out[0], g->ev_phase[0], g->ev0));     // query prep + workspace memsets
    DIF_CUDA_OK(cudaEventElapsedTime(&out[1], g->ev0, g->ev1));             // tensor-core filter
    DIF_CUDA_OK(cudaEventElapsedTime(&out[2], g->ev1, g->ev_phase[1]));     // window + canonical re-rank
  }
  DIF_CUDA_OK(cudaEventElapsedTime(&out[3], g->ev_phase[1], g->ev_phase[2]));   // exact path for flagged queries
  if (g->phase_sharded) DIF_CUDA_OK(cudaEventElapsedTime(&out[4], g->ev_phase[2], g->ev_phase[3]));   // exchange + merge
  return DIF_OK;
}

int dif_gallery_get_rows(dif_gallery_t* g, int64_t row0, int64_t n, float* out, void* stream) {
  DIF_REQUIRE(g && out && row0 >= 0 && n >= 0 && row0 + n <= g->size, DIF_ERR_INVALID, "dif_gallery_get_rows: range");
  if (n == 0) return DIF_OK;
  gather_rows_kernel<<<(unsigned)std::min<int64_t>((n * g->D + 255) / 256, 148 * 8), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(g->g0, g->g1, row0, n, g->D, out);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

int dif_topk_merge(const float* scores, const int64_t* grows, const int64_t* ids, int world, int n_queries, int k,
                   int metric, float* out_scores, int64_t* out_grows, int64_t* out_ids, void* stream) {
  DIF_REQUIRE(scores && grows && out_scores && out_grows, DIF_ERR_INVALID, "dif_topk_merge: null argument");
  DIF_REQUIRE(world >= 1 && world <= 64 && n_queries > 0 && k >= 1 && k <= DIF_MAX_TOPK, DIF_ERR_INVALID,
              "dif_topk_merge: world %d, n_queries %d, k %d", world, n_queries, k);
  const int n = world * k;
  const size_t smem = ((n * 4 + 7) & ~7) + (size_t)n * 8;
  topk_merge_kernel<<<std::min(n_queries, 148 * 8), 128, smem, static_cast<cudaStream_t>(stream)>>>(
      scores, grows, ids, world, n_queries, k, metric, out_scores, out_grows, out_ids);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

// raw synthetic rows (no normalisation): out[n*D] = synth_value(seed, row, d, D); rows_idx NULL -> row0 + r
int dif_synth_fill(float* out, uint64_t seed, int64_t row0, const int64_t* rows_idx, int64_t n, int D, void* stream);

}  // extern "C"

namespace dif {
__global__ void synth_fill_kernel(float* out, uint64_t seed, int64_t row0, const int64_t* rows_idx, int64_t n, int D) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * D; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D;
    const int d = (int)(i - r * D);
    const int64_t row = rows_idx ? rows_idx[r] : row0 + r;
    out[i] = synth_value(seed, (uint64_t)row, (uint64_t)d, (uint64_t)D);
  }
}
}  // namespace dif

extern "C" int dif_synth_fill(float* out, uint64_t seed, int64_t row0, const int64_t* rows_idx, int64_t n, int D,
                              void* stream) {
  DIF_REQUIRE(out && n >= 0 && D > 0, DIF_ERR_INVALID, "dif_synth_fill: invalid argument");
  if (n == 0) return DIF_OK;
  dif::synth_fill_kernel<<<(unsigned)std::min<int64_t>((n * D + 255) / 256, 148 * 16), 256, 0,
                           static_cast<cudaStream_t>(stream)>>>(out, seed, row0, rows_idx, n, D);
  DIF_LAUNCH_OK();
  return DIF_OK;
}
