// Row-wise pair kernels of the verification / siamese / explicit-triplet paths (all HBM- or
// latency-bound; one warp per row, canonical fp32 reductions of dif_canon.cuh):
//   dif_pair_distance      deep_insight_face/evaluation/utility.py:52-66   (distance)
//   dif_fold_mean          deep_insight_face/evaluation/utility.py:98-102,144-147 (train-fold mean)
//   dif_threshold_sweep    deep_insight_face/evaluation/utility.py:36-49,69-77 (all thresholds, all folds, one pass)
//   dif_triplet_apn        deep_insight_face/networks/triplet.py:16-46     (triplet_loss)
//   dif_euclidean_distance deep_insight_face/networks/siamese.py:22-24
//   dif_contrastive_loss   deep_insight_face/networks/siamese.py:32-39
#include <algorithm>

#include "dif_canon.cuh"
#include "dif_common.cuh"

namespace dif {

constexpr float kPi = 3.14159265358979323846f;

// ------------------------------------------------------------------------------------------
// pair distance: out[r] = sum((a-b)^2)  |  arccos(a.b / (|a||b|)) / pi, optional mean subtraction
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pair_distance_kernel(const float* __restrict__ e1, const float* __restrict__ e2,
                                                            int64_t n, int D, int metric,
                                                            const float* __restrict__ mean, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += warps) {
    const float* a = e1 + r * D;
    const float* b = e2 + r * D;
    float acc = 0.f, na = 0.f, nb = 0.f;
    for (int d = lane; d < D; d += 32) {
      float x = a[d], y = b[d];
      if (mean) {
        const float m = mean[d];
        x = __fsub_rn(x, m);
        y = __fsub_rn(y, m);
      }
      if (metric == 0) {
        const float t = __fsub_rn(x, y);
        acc = __fmaf_rn(t, t, acc);
      } else {
        acc = __fmaf_rn(x, y, acc);
        na = __fmaf_rn(x, x, na);
        nb = __fmaf_rn(y, y, nb);
      }
    }
    acc = canon_tree(acc);
    if (metric == 0) {
      if (lane == 0) out[r] = acc;
    } else {
      na = canon_tree(na);
      nb = canon_tree(nb);
      if (lane == 0) {
        const float sim = __fdiv_rn(acc, __fmul_rn(__fsqrt_rn(na), __fsqrt_rn(nb)));
        out[r] = __fdiv_rn(acosf(sim), kPi);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// train-fold mean: mean[f][d] = mean over rows NOT in fold f of both embedding sets.
// Folds are the contiguous KFold(shuffle=False) ranges; one block per fold sums its own rows
// (fixed order, fp64 accumulation), a second kernel turns fold sums into leave-one-fold-out means.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fold_sum_kernel(const float* __restrict__ e1, const float* __restrict__ e2,
                                                       const int64_t* __restrict__ fold_begin, int D,
                                                       double* __restrict__ sums /* [F][D] */) {
  const int f = blockIdx.x;
  const int64_t r0 = fold_begin[f], r1 = fold_begin[f + 1];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) s += (double)e1[r * D + d] + (double)e2[r * D + d];
    sums[(size_t)f * D + d] = s;
  }
}
__global__ void fold_mean_kernel(const double* __restrict__ sums, const int64_t* __restrict__ fold_begin, int F, int D,
                                 float* __restrict__ mean /* [F][D] */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F * D) return;
  const int f = i / D, d = i - f * D;
  double s = 0.0;
  for (int g = 0; g < F; ++g)
    if (g != f) s += sums[(size_t)g * D + d];
  const int64_t n_train = fold_begin[F] - (fold_begin[f + 1] - fold_begin[f]);
  mean[i] = n_train > 0 ? (float)(s / (2.0 * (double)n_train)) : 0.f;
}

// ------------------------------------------------------------------------------------------
// threshold sweep: predict = (double)dist < thr[t] (np.less against float64 thresholds).
// For ascending thresholds the first t with dist < thr[t] is found by binary search; a histogram of
// that index per (fold, issame) followed by a prefix sum yields tp/fp for every threshold at once:
//   tp[f][t] = #{i in fold f, same, first(i) <= t},  fn = n_same - tp,  tn = n_diff - fp.
// Unsorted thresholds take the direct O(N*T) kernel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sweep_hist_kernel(const float* __restrict__ dist,
                                                         const uint8_t* __restrict__ issame,
                                                         const int32_t* __restrict__ fold, int64_t n, int F,
                                                         const double* __restrict__ thr, int T,
                                                         unsigned int* __restrict__ hist /* [F][2][T+1] */) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = (double)dist[i];
    int lo = 0, hi = T;  // first t in [0, T] with d < thr[t]; T = none
    if (d != d) {
      lo = T;  // NaN < x is false for every threshold
    } else {
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (d < thr[mid]) hi = mid;
        else lo = mid + 1;
      }
    }
    const int f = fold ? fold[i] : 0;
    if (f < 0 || f >= F) continue;
    atomicAdd(&hist[((size_t)f * 2 + (issame[i] ? 1 : 0)) * (T + 1) + lo], 1u);
  }
}
// Same histogram with block-private bins in shared memory (and the thresholds staged beside them): every distance
// costs a binary search in shared memory and one shared-memory atomic; a block adds its non-empty bins to the global
// histogram once at the end.  The global-atomic form above serialises when the distances crowd into a few bins -
// 1M pairs that all exceed the last threshold are 1M atomics on two addresses (0.35 ms; 14 us here).
__global__ void __launch_bounds__(512) sweep_hist_smem_kernel(const float* __restrict__ dist,
                                                              const uint8_t* __restrict__ issame,
                                                              const int32_t* __restrict__ fold, int64_t n, int F,
                                                              const double* __restrict__ thr, int T,
                                                              unsigned int* __restrict__ hist /* [F][2][T+1] */) {
  extern __shared__ double s_thr_raw[];
  double* s_thr = s_thr_raw;                                          // [T]
  unsigned int* s_h = reinterpret_cast<unsigned int*>(s_thr + T);     // [F][2][T+1]
  const int bins = F * 2 * (T + 1);
  for (int t = threadIdx.x; t < T; t += blockDim.x) s_thr[t] = thr[t];
  for (int b = threadIdx.x; b < bins; b += blockDim.x) s_h[b] = 0u;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = (double)dist[i];
    int lo = 0, hi = T;  // first t in [0, T] with d < thr[t]; T = none
    if (d != d) {
      lo = T;
    } else {
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (d < s_thr[mid]) hi = mid;
        else lo = mid + 1;
      }
    }
    const int f = fold ? fold[i] : 0;
    if (f < 0 || f >= F) continue;
    atomicAdd(&s_h[(f * 2 + (issame[i] ? 1 : 0)) * (T + 1) + lo], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < bins; b += blockDim.x) {
    const unsigned int v = s_h[b];
    if (v) atomicAdd(&hist[b], v);
  }
}
// one block per (fold, class): inclusive scan of the histogram -> counts
__global__ void __launch_bounds__(256) sweep_scan_kernel(const unsigned int* __restrict__ hist, int T,
                                                         int64_t* __restrict__ counts /* [F][T][4] */) {
  __shared__ long long warp_tot[8];
  __shared__ long long carry_s;
  const int f = blockIdx.x >> 1, same = blockIdx.x & 1;
  const unsigned int* h = hist + (size_t)blockIdx.x * (T + 1);
  // total of this (fold, class) including the "never predicted" bin T
  long long part = 0;
  for (int t = threadIdx.x; t <= T; t += blockDim.x) part += h[t];
  for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = part;
  __syncthreads();
  long long total = 0;
  for (int w = 0; w < 8; ++w) total += warp_tot[w];
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < T; base += blockDim.x) {
    const int t = base + threadIdx.x;
    long long v = t < T ? (long long)h[t] : 0;
    // block inclusive scan
    long long x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    long long off = carry_s;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += warp_tot[w];
    const long long incl = x + off;
    if (t < T) {
      int64_t* c = counts + ((size_t)f * T + t) * 4;
      if (same) {
        c[0] = incl;          // tp
        c[3] = total - incl;  // fn
      } else {
        c[1] = incl;          // fp
        c[2] = total - incl;  // tn
      }
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = incl;
    __syncthreads();
  }
}
// direct form for arbitrary threshold order: one block per (fold, threshold)
__global__ void __launch_bounds__(256) sweep_direct_kernel(const float* __restrict__ dist,
                                                           const uint8_t* __restrict__ issame,
                                                           const int32_t* __restrict__ fold, int64_t n,
                                                           const double* __restrict__ thr, int T,
                                                           int64_t* __restrict__ counts) {
  __shared__ long long red[4][8];
  const int f = blockIdx.x / T, t = blockIdx.x - f * T;
  const double th = thr[t];
  long long c[4] = {0, 0, 0, 0};
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    if (fold && fold[i] != f) continue;
    const bool pred = (double)dist[i] < th, same = issame[i] != 0;
    c[0] += pred && same;
    c[1] += pred && !same;
    c[2] += !pred && !same;
    c[3] += !pred && same;
  }
  for (int j = 0; j < 4; ++j) {
    for (int o = 16; o >= 1; o >>= 1) c[j] += __shfl_xor_sync(0xffffffffu, c[j], o);
    if ((threadIdx.x & 31) == 0) red[j][threadIdx.x >> 5] = c[j];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    long long s = 0;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    counts[((size_t)f * T + t) * 4 + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------
// explicit (anchor | positive | negative) triplet loss on [B, 3D] rows, fwd + optional bwd
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) triplet_apn_kernel(const float* __restrict__ y, int B, int D, float alpha,
                                                          float* __restrict__ loss, const float* __restrict__ dloss,
                                                          float* __restrict__ dy) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (blockDim.x >> 5);
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < B; r += warps) {
    const float* a = y + (size_t)r * 3 * D;
    const float* p = a + D;
    const float* ng = a + 2 * D;
    float dp = 0.f, dn = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float tp = __fsub_rn(a[d], p[d]), tn = __fsub_rn(a[d], ng[d]);
      dp = __fmaf_rn(tp, tp, dp);
      dn = __fmaf_rn(tn, tn, dn);
    }
    dp = canon_tree(dp);
    dn = canon_tree(dn);
    const float basic = __fadd_rn(__fsub_rn(dp, dn), alpha);
    if (lane == 0) loss[r] = fmaxf(basic, 0.f);
    if (dy) {
      // tf.maximum passes the gradient to its first argument when basic >= 0
      const float g = basic >= 0.f ? (dloss ? dloss[r] : 1.f) : 0.f;
      float* ga = dy + (size_t)r * 3 * D;
      for (int d = lane; d < D; d += 32) {
        const float av = a[d], pv = p[d], nv = ng[d];
        ga[d] = 2.f * g * (nv - pv);           // d/da [(a-p)^2 - (a-n)^2]
        ga[D + d] = -2.f * g * (av - pv);
        ga[2 * D + d] = 2.f * g * (av - nv);
      }
    }
  }
}

__global__ void __launch_bounds__(256) euclid_dist_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                          int B, int D, float eps, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (blockDim.x >> 5);
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < B; r += warps) {
    const float s = canon_sqdist_warp(x + (size_t)r * D, y + (size_t)r * D, D);
    if (lane == 0) out[r] = __fsqrt_rn(fmaxf(s, eps));
  }
}

// mean(y * d^2 + (1 - y) * max(margin - d, 0)^2); single block, fixed summation order
__global__ void __launch_bounds__(256) contrastive_kernel(const float* __restrict__ yt, const float* __restrict__ dist,
                                                          int B, float margin, float* __restrict__ out,
                                                          float* __restrict__ dd) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float y = yt[i], d = dist[i];
    const float m = fmaxf(margin - d, 0.f);
    acc += (double)(y * d * d + (1.f - y) * m * m);
    if (dd) dd[i] = (2.f * y * d - 2.f * (1.f - y) * m) / (float)B;
  }
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    out[0] = (float)(s / (double)B);
  }
}

static int row_grid(int64_t rows) {
  const int sms = std::max(1, device_sm_count());
  return (int)std::max<int64_t>(1, std::min<int64_t>((rows + 7) / 8, (int64_t)sms * 16));
}

// pinned staging shared by the *_host entry points of this file (grown on demand, per thread)
struct HostStage {
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
static thread_local HostStage g_stage;

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace dif

using namespace dif;

// ---------------------------------------------------------------------------------------------
// Embedding head: tf.nn.l2_normalize(x, axis=1) = x * rsqrt(max(sum x^2, 1e-12))
// (deep_insight_face/networks/inceptionv3.py:305, networks/triplet.py:138) and its backward
// dx = inv * (g - y (y . g)), a plain scaling where the squared norm was clamped.  One warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2_normalize_kernel(const float* __restrict__ x, int64_t n, int D,
                                                           float* __restrict__ y, float* __restrict__ inv_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * 8;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += warps) {
    const float* s = x + r * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = __fmaf_rn(s[d], s[d], acc);
    const float inv = canon_inv_norm(canon_tree(acc));
    for (int d = lane; d < D; d += 32) y[r * D + d] = __fmul_rn(s[d], inv);
    if (lane == 0 && inv_out) inv_out[r] = inv;
  }
}

__global__ void __launch_bounds__(256) l2_normalize_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                               const float* __restrict__ inv, int64_t n, int D,
                                                               float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * 8;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += warps) {
    const float dot = canon_dot_warp(g + r * D, y + r * D, D);
    const float iv = inv[r];
    const bool clamped = iv >= 0.99e6f;   // 1 / sqrt(1e-12): max() passed no gradient to the norm
    for (int d = lane; d < D; d += 32) {
      const float gv = g[r * D + d];
      dx[r * D + d] = clamped ? iv * gv : iv * (gv - y[r * D + d] * dot);
    }
  }
}

extern "C" {

int dif_l2_normalize(const float* x, int64_t n, int D, float* y, float* inv_norm, void* stream) {
  DIF_REQUIRE(x && y && n >= 0 && D >= 1, DIF_ERR_INVALID, "dif_l2_normalize: invalid argument");
  if (n == 0) return DIF_OK;
  l2_normalize_kernel<<<(unsigned)std::min<int64_t>((n + 7) / 8, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, n, D, y, inv_norm);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

int dif_l2_normalize_bwd(const float* g, const float* y, const float* inv_norm, int64_t n, int D, float* dx, void* stream) {
  DIF_REQUIRE(g && y && inv_norm && dx && n >= 0 && D >= 1, DIF_ERR_INVALID, "dif_l2_normalize_bwd: invalid argument");
  if (n == 0) return DIF_OK;
  l2_normalize_bwd_kernel<<<(unsigned)std::min<int64_t>((n + 7) / 8, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      g, y, inv_norm, n, D, dx);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

int dif_pair_distance(const float* e1, const float* e2, int64_t N, int D, int metric, const float* mean, float* out,
                      void* stream) {
  DIF_REQUIRE(e1 && e2 && out && N >= 0 && D > 0, DIF_ERR_INVALID, "dif_pair_distance: invalid argument");
  // deep_insight_face/evaluation/utility.py:64 raises RuntimeError('Undefined distance metric %d')
  DIF_REQUIRE(metric == 0 || metric == 1, DIF_ERR_INVALID, "Undefined distance metric %d", metric);
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  if (N == 0) return DIF_OK;
  pair_distance_kernel<<<row_grid(N), 256, 0, static_cast<cudaStream_t>(stream)>>>(e1, e2, N, D, metric, mean, out);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

int dif_pair_distance_host(const float* e1_host, const float* e2_host, int64_t N, int D, int metric,
                           const float* mean_host, float* out_host) {
  DIF_REQUIRE(e1_host && e2_host && out_host && N >= 0 && D > 0, DIF_ERR_INVALID, "dif_pair_distance_host: invalid argument");
  DIF_REQUIRE(metric == 0 || metric == 1, DIF_ERR_INVALID, "Undefined distance metric %d", metric);
  if (N == 0) return DIF_OK;
  const size_t eb = al256((size_t)N * D * 4), mb = al256((size_t)D * 4), ob = al256((size_t)N * 4);
  if (int rc = g_stage.ensure(2 * eb + mb + ob)) return rc;
  char* h = (char*)g_stage.h;
  char* d = (char*)g_stage.d;
  memcpy(h, e1_host, (size_t)N * D * 4);
  memcpy(h + eb, e2_host, (size_t)N * D * 4);
  if (mean_host) memcpy(h + 2 * eb, mean_host, (size_t)D * 4);
  DIF_CUDA_OK(cudaMemcpyAsync(d, h, 2 * eb + mb, cudaMemcpyHostToDevice, g_stage.st));
  if (int rc = dif_pair_distance((const float*)d, (const float*)(d + eb), N, D, metric,
                                 mean_host ? (const float*)(d + 2 * eb) : nullptr, (float*)(d + 2 * eb + mb), g_stage.st))
    return rc;
  DIF_CUDA_OK(cudaMemcpyAsync(h + 2 * eb + mb, d + 2 * eb + mb, (size_t)N * 4, cudaMemcpyDeviceToHost, g_stage.st));
  DIF_CUDA_OK(cudaStreamSynchronize(g_stage.st));
  memcpy(out_host, h + 2 * eb + mb, (size_t)N * 4);
  return DIF_OK;
}

int dif_fold_mean(const float* e1, const float* e2, const int64_t* fold_begin, int n_folds, int D, double* workspace,
                  float* mean, void* stream) {
  DIF_REQUIRE(e1 && e2 && fold_begin && workspace && mean && n_folds >= 1 && D > 0, DIF_ERR_INVALID,
              "dif_fold_mean: invalid argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  fold_sum_kernel<<<n_folds, 256, 0, st>>>(e1, e2, fold_begin, D, workspace);
  DIF_LAUNCH_OK();
  fold_mean_kernel<<<(n_folds * D + 255) / 256, 256, 0, st>>>(workspace, fold_begin, n_folds, D, mean);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

int dif_threshold_sweep(const float* dist, const uint8_t* issame, const int32_t* fold, int64_t N, int n_folds,
                        const double* thresholds, int T, int ascending, unsigned int* workspace, int64_t* counts,
                        void* stream) {
  DIF_REQUIRE(dist && issame && thresholds && counts && N >= 0 && T >= 1 && n_folds >= 1, DIF_ERR_INVALID,
              "dif_threshold_sweep: invalid argument");
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ascending) {
    DIF_REQUIRE(workspace, DIF_ERR_INVALID, "dif_threshold_sweep: workspace [n_folds*2*(T+1)] u32 is required");
    const size_t hb = (size_t)n_folds * 2 * (T + 1) * sizeof(unsigned int);
    DIF_CUDA_OK(cudaMemsetAsync(workspace, 0, hb, st));
    if (N > 0) {
      const size_t sm_bytes = (size_t)T * 8 + hb;
      if (sm_bytes <= 200 * 1024 && N >= 4096) {   // block-private bins fit one SM's shared memory
        static bool configured = false;
        if (!configured) {
          DIF_CUDA_OK(cudaFuncSetAttribute(sweep_hist_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
          configured = true;
        }
        const int per_sm = sm_bytes <= 100 * 1024 ? 2 : 1;
        const int blocks = (int)std::min<int64_t>((N + 511) / 512, (int64_t)device_sm_count() * per_sm);
        sweep_hist_smem_kernel<<<blocks, 512, sm_bytes, st>>>(dist, issame, fold, N, n_folds, thresholds, T, workspace);
      } else {
        const int blocks = (int)std::min<int64_t>((N + 255) / 256, (int64_t)device_sm_count() * 8);
        sweep_hist_kernel<<<blocks, 256, 0, st>>>(dist, issame, fold, N, n_folds, thresholds, T, workspace);
      }
      DIF_LAUNCH_OK();
    }
    sweep_scan_kernel<<<n_folds * 2, 256, 0, st>>>(workspace, T, counts);
    DIF_LAUNCH_OK();
  } else {
    sweep_direct_kernel<<<n_folds * T, 256, 0, st>>>(dist, issame, fold, N, thresholds, T, counts);
    DIF_LAUNCH_OK();
  }
  return DIF_OK;
}

int dif_threshold_sweep_host(const float* dist_host, const uint8_t* issame_host, const int32_t* fold_host, int64_t N,
                             int n_folds, const double* thresholds_host, int T, int64_t* counts_host) {
  DIF_REQUIRE(dist_host && issame_host && thresholds_host && counts_host && N >= 0 && T >= 1 && n_folds >= 1,
              DIF_ERR_INVALID, "dif_threshold_sweep_host: invalid argument");
  bool asc = true;
  for (int t = 1; t < T; ++t) asc = asc && thresholds_host[t - 1] <= thresholds_host[t];
  const size_t db = al256((size_t)N * 4), sb = al256((size_t)N), fb = al256((size_t)N * 4), tb = al256((size_t)T * 8);
  const size_t wb = al256((size_t)n_folds * 2 * (T + 1) * 4), cb = al256((size_t)n_folds * T * 4 * 8);
  if (int rc = g_stage.ensure(db + sb + fb + tb + wb + cb)) return rc;
  char* h = (char*)g_stage.h;
  char* d = (char*)g_stage.d;
  memcpy(h, dist_host, (size_t)N * 4);
  memcpy(h + db, issame_host, (size_t)N);
  if (fold_host) memcpy(h + db + sb, fold_host, (size_t)N * 4);
  memcpy(h + db + sb + fb, thresholds_host, (size_t)T * 8);
  DIF_CUDA_OK(cudaMemcpyAsync(d, h, db + sb + fb + tb, cudaMemcpyHostToDevice, g_stage.st));
  char* dw = d + db + sb + fb + tb;
  if (int rc = dif_threshold_sweep((const float*)d, (const uint8_t*)(d + db),
                                   fold_host ? (const int32_t*)(d + db + sb) : nullptr, N, n_folds,
                                   (const double*)(d + db + sb + fb), T, asc ? 1 : 0, (unsigned int*)dw,
                                   (int64_t*)(dw + wb), g_stage.st))
    return rc;
  DIF_CUDA_OK(cudaMemcpyAsync(h + db + sb + fb + tb + wb, dw + wb, (size_t)n_folds * T * 4 * 8, cudaMemcpyDeviceToHost,
                              g_stage.st));
  DIF_CUDA_OK(cudaStreamSynchronize(g_stage.st));
  memcpy(counts_host, h + db + sb + fb + tb + wb, (size_t)n_folds * T * 4 * 8);
  return DIF_OK;
}

int dif_triplet_apn(const float* y_pred, int B, int D, float alpha, float* loss, const float* dloss, float* dy,
                    void* stream) {
  DIF_REQUIRE(y_pred && loss && B >= 0 && D > 0, DIF_ERR_INVALID, "dif_triplet_apn: invalid argument");
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  if (B == 0) return DIF_OK;
  triplet_apn_kernel<<<row_grid(B), 256, 0, static_cast<cudaStream_t>(stream)>>>(y_pred, B, D, alpha, loss, dloss, dy);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

int dif_euclidean_distance(const float* x, const float* y, int B, int D, float eps, float* out, void* stream) {
  DIF_REQUIRE(x && y && out && B >= 0 && D > 0, DIF_ERR_INVALID, "dif_euclidean_distance: invalid argument");
  DIF_REQUIRE(device_sm_count() > 0, DIF_ERR_STATE, "dif_init has not succeeded on this process");
  if (B == 0) return DIF_OK;
  euclid_dist_kernel<<<row_grid(B), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, B, D, eps, out);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

int dif_contrastive_loss(const float* y_true, const float* dist, int B, float margin, float* out, float* dd,
                         void* stream) {
  DIF_REQUIRE(y_true && dist && out && B > 0, DIF_ERR_INVALID, "dif_contrastive_loss: invalid argument");
  contrastive_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(y_true, dist, B, margin, out, dd);
  DIF_LAUNCH_OK();
  return DIF_OK;
}

}  // extern "C"
