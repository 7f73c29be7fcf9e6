/* dif_oracle.c - CPU oracle for the embedding-space distance path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library; nothing under deep_insight_face_b200/ does.
 *
 * What it restates (paths relative to the reference repository root):
 *   - pair distances           deep_insight_face/evaluation/utility.py:52-66  (distance)
 *   - threshold counts         deep_insight_face/evaluation/utility.py:36-49  (calculate_accuracy)
 *                              deep_insight_face/evaluation/utility.py:69-77  (calculate_val_far)
 *   - row normalisation        tf.nn.l2_normalize as used by common/losses.py:39 and
 *                              networks/inceptionv3.py:305: x * rsqrt(max(sum(x^2), 1e-12))
 *   - Gram / distance matrix   common/losses.py:40 (cosine) and :63-65 (squared L2)
 *   - 1:N top-k search         ABSENT from the reference (predictions.py:126 is 1:1 only); the
 *                              semantics are this build's: k best rows by (score best-first,
 *                              row index ascending).  Pinned as a ranking by the reference's
 *                              distance() (utility.py:52-66, tests/golden/make_golden_gallery.py);
 *                              the tie rule is this build's.
 *
 * Canonical fp32 arithmetic.  The GPU library promises bit-identical "exact" results, so every
 * reduction here uses the same fixed order as csrc/dif_canon.cuh:
 *   32 strided chains (chain l takes d = l, l+32, ...) of single-rounded fma, starting at +0,
 *   combined by the butterfly t[i] += t[i ^ o], o = 16, 8, 4, 2, 1.
 * Build with -ffp-contract=off (no implicit contraction) and -mfma (fmaf -> one instruction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

static inline float tree32(float t[32]) {
  for (int o = 16; o >= 1; o >>= 1) {
    float n[32];
    for (int i = 0; i < 32; ++i) n[i] = t[i] + t[i ^ o];
    memcpy(t, n, sizeof(n));
  }
  return t[0];
}

EXPORT float dif_or_canon_dot(const float* a, const float* b, int D) {
  float t[32];
  for (int l = 0; l < 32; ++l) t[l] = 0.f;
  int d = 0;
  for (; d + 32 <= D; d += 32)
    for (int l = 0; l < 32; ++l) t[l] = fmaf(a[d + l], b[d + l], t[l]);
  for (int l = 0; d + l < D; ++l) t[l] = fmaf(a[d + l], b[d + l], t[l]);
  return tree32(t);
}

EXPORT float dif_or_canon_sqdist(const float* a, const float* b, int D) {
  float t[32];
  for (int l = 0; l < 32; ++l) t[l] = 0.f;
  int d = 0;
  for (; d + 32 <= D; d += 32)
    for (int l = 0; l < 32; ++l) {
      const float x = a[d + l] - b[d + l];
      t[l] = fmaf(x, x, t[l]);
    }
  for (int l = 0; d + l < D; ++l) {
    const float x = a[d + l] - b[d + l];
    t[l] = fmaf(x, x, t[l]);
  }
  return tree32(t);
}

EXPORT float dif_or_inv_norm(float ss) { return 1.0f / sqrtf(fmaxf(ss, 1e-12f)); }

/* tf.nn.l2_normalize(x, 1) in canonical arithmetic; out may alias x */
EXPORT void dif_or_normalize_rows(const float* x, int64_t n, int D, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    const float* xr = x + r * (int64_t)D;
    const float inv = dif_or_inv_norm(dif_or_canon_dot(xr, xr, D));
    for (int d = 0; d < D; ++d) out[r * (int64_t)D + d] = xr[d] * inv;
  }
}

EXPORT void dif_or_row_sqnorm(const float* x, int64_t n, int D, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) out[r] = dif_or_canon_dot(x + r * (int64_t)D, x + r * (int64_t)D, D);
}

/* out[i*nb + j] = canon_dot(a_i, b_j)  (metric 1)  or canon_sqdist(a_i, b_j) (metric 0) */
EXPORT void dif_or_cross(const float* a, int64_t na, const float* b, int64_t nb, int D, int metric, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < na; ++i)
    for (int64_t j = 0; j < nb; ++j)
      out[i * nb + j] = metric == 1 ? dif_or_canon_dot(a + i * (int64_t)D, b + j * (int64_t)D, D)
                                    : dif_or_canon_sqdist(a + i * (int64_t)D, b + j * (int64_t)D, D);
}

/* ---- synthetic data: identical integer arithmetic in csrc/dif_canon.cuh:synth_value --------- */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
EXPORT float dif_or_synth_value(uint64_t seed, uint64_t row, uint64_t col, uint64_t dim) {
  const uint64_t h = splitmix64(seed * 0xD1342543DE82EF95ull + row * dim + col);
  const int s = (int)(h & 0xFFFF) + (int)((h >> 16) & 0xFFFF) + (int)((h >> 32) & 0xFFFF) + (int)((h >> 48) & 0xFFFF) -
                131070;
  return (float)s * 2.6428998e-05f;
}
EXPORT void dif_or_synth_rows(uint64_t seed, int64_t row0, int64_t n, int D, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r)
    for (int d = 0; d < D; ++d)
      out[r * (int64_t)D + d] = dif_or_synth_value(seed, (uint64_t)(row0 + r), (uint64_t)d, (uint64_t)D);
}

/* ---- 1:N search ----------------------------------------------------------------------------
 * gallery / queries hold the canonical stored rows (already normalised for cosine).
 * metric 1: score = canon_dot, larger is better; metric 0: score = canon_sqdist, smaller is better.
 * Output per query: k entries ordered best-first, ties by ascending row; slots past n_rows: row -1, score 0. */
typedef struct {
  float better; /* larger wins */
  int64_t row;
} cand_t;

static inline int cand_before(const cand_t* a, const cand_t* b) {
  if (a->better > b->better) return 1;
  if (a->better < b->better) return 0;
  return a->row < b->row;
}

EXPORT void dif_or_gallery_search(const float* gallery, int64_t n_rows, int D, int metric, const float* queries,
                                  int n_queries, int k, float* scores, int64_t* rows) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int q = 0; q < n_queries; ++q) {
    cand_t best[64];
    int nb = 0;
    const float* qv = queries + (int64_t)q * D;
    for (int64_t r = 0; r < n_rows; ++r) {
      const float* g = gallery + r * (int64_t)D;
      cand_t c;
      c.better = metric == 1 ? dif_or_canon_dot(qv, g, D) : -dif_or_canon_sqdist(qv, g, D);
      c.row = r;
      if (nb == k && !cand_before(&c, &best[nb - 1])) continue;
      int p = nb < k ? nb : k - 1;
      while (p > 0 && cand_before(&c, &best[p - 1])) {
        best[p] = best[p - 1];
        --p;
      }
      best[p] = c;
      if (nb < k) ++nb;
    }
    for (int i = 0; i < k; ++i) {
      if (i < nb) {
        scores[(int64_t)q * k + i] = metric == 1 ? best[i].better : -best[i].better;
        rows[(int64_t)q * k + i] = best[i].row;
      } else {
        scores[(int64_t)q * k + i] = 0.f;
        rows[(int64_t)q * k + i] = -1;
      }
    }
  }
}

/* merge `world` shard results [world][Q][k] (global rows, -1 = empty) into [Q][k] */
EXPORT void dif_or_topk_merge(const float* scores, const int64_t* grows, int world, int n_queries, int k, int metric,
                              float* out_scores, int64_t* out_grows) {
  for (int q = 0; q < n_queries; ++q) {
    cand_t best[64];
    int nb = 0;
    for (int w = 0; w < world; ++w)
      for (int i = 0; i < k; ++i) {
        const int64_t at = ((int64_t)w * n_queries + q) * k + i;
        if (grows[at] < 0) continue;
        cand_t c;
        c.better = metric == 1 ? scores[at] : -scores[at];
        c.row = grows[at];
        if (nb == k && !cand_before(&c, &best[nb - 1])) continue;
        int p = nb < k ? nb : k - 1;
        while (p > 0 && cand_before(&c, &best[p - 1])) {
          best[p] = best[p - 1];
          --p;
        }
        best[p] = c;
        if (nb < k) ++nb;
      }
    for (int i = 0; i < k; ++i) {
      out_scores[(int64_t)q * k + i] = i < nb ? (metric == 1 ? best[i].better : -best[i].better) : 0.f;
      out_grows[(int64_t)q * k + i] = i < nb ? best[i].row : -1;
    }
  }
}

/* ---- pair verification ---------------------------------------------------------------------
 * utility.py:52-66.  metric 0: sum((a-b)^2) in canonical order; metric 1: arccos(dot/(|a||b|))/pi. */
EXPORT void dif_or_pair_distance(const float* e1, const float* e2, int64_t n, int D, int metric, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    const float* a = e1 + r * (int64_t)D;
    const float* b = e2 + r * (int64_t)D;
    if (metric == 0) {
      out[r] = dif_or_canon_sqdist(a, b, D);
    } else {
      const float dot = dif_or_canon_dot(a, b, D);
      const float na = sqrtf(dif_or_canon_dot(a, a, D));
      const float nb = sqrtf(dif_or_canon_dot(b, b, D));
      out[r] = acosf(dot / (na * nb)) / 3.14159265358979323846f;
    }
  }
}

/* utility.py:36-49 / :69-77 for T thresholds: counts[t] = {tp, fp, tn, fn}, predict = dist < thr.
 * np.less(dist, threshold) compares the float32 distances with the FLOAT64 thresholds of np.arange: the
 * comparison is done in double (a distance that equals a threshold after rounding to float32 can still be
 * strictly below the float64 threshold). */
EXPORT void dif_or_threshold_sweep(const float* dist, const uint8_t* issame, const uint8_t* select, int64_t n,
                                   const double* thr, int T, int64_t* counts) {
  for (int t = 0; t < T; ++t) {
    int64_t tp = 0, fp = 0, tn = 0, fn = 0;
    for (int64_t i = 0; i < n; ++i) {
      if (select && !select[i]) continue;
      const int pred = (double)dist[i] < thr[t];
      const int same = issame[i] != 0;
      tp += pred && same;
      fp += pred && !same;
      tn += !pred && !same;
      fn += !pred && same;
    }
    counts[4 * t + 0] = tp;
    counts[4 * t + 1] = fp;
    counts[4 * t + 2] = tn;
    counts[4 * t + 3] = fn;
  }
}

EXPORT int dif_or_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
