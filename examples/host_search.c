/* A plain-C99 host driving the path through include/dif_b200.h - no Python, no torch: what a non-Python maintainer
 * of a face-recognition service would link.  Self-checking (no oracle needed): every result is compared with a
 * double-precision loop written out here.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/host_search.c -Ldeep_insight_face_b200 -ldif_b200 \
 *       -Wl,-rpath,$PWD/deep_insight_face_b200 -Wl,--allow-shlib-undefined -lm -o examples/host_search
 *
 * 1. 1:N search (the 1:N extension of predictions.py:104-150): enrol 20 000 x 128 rows from host memory, probe 64
 *    noisy copies, top-5 by cosine; the best hit must be the copied row and its score the cosine computed here.
 * 2. pair distances (evaluation/utility.py:52-66), both metrics, 1 000 pairs.
 * 3. one batch-hard triplet step (common/losses.py:33-51) on an 18 x 4 batch: mined columns must carry the right
 *    labels and the per-anchor loss must be max(neg - pos + alpha, 0) of the cosines recomputed here.
 * Exit code 0 and "host_search OK" on success; 3 with the library's message when there is no B200 (no CPU fallback). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "dif_b200.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static float unit_noise(void) { /* sum of four uniforms, centred: roughly normal, deterministic */
  float s = 0.f;
  for (int i = 0; i < 4; ++i) {
    rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
    s += (float)((rng_state >> 40) & 0xFFFFFF) / 16777216.0f;
  }
  return (s - 2.0f) * 1.7320508f;
}

static double cosine(const float* a, const float* b, int D) {
  double ab = 0, aa = 0, bb = 0;
  for (int d = 0; d < D; ++d) {
    ab += (double)a[d] * b[d];
    aa += (double)a[d] * a[d];
    bb += (double)b[d] * b[d];
  }
  return ab / sqrt(aa * bb);
}

#define CHECK(call)                                                                          \
  do {                                                                                       \
    int rc_ = (call);                                                                        \
    if (rc_ != DIF_OK) {                                                                     \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, dif_last_error());                 \
      return rc_ == DIF_ERR_NO_DEVICE ? 3 : 1;                                               \
    }                                                                                        \
  } while (0)
#define REQUIRE(cond, ...)                                                                   \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      fprintf(stderr, "check failed: %s: ", #cond);                                          \
      fprintf(stderr, __VA_ARGS__);                                                          \
      fprintf(stderr, "\n");                                                                 \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

int main(void) {
  CHECK(dif_init(0));
  printf("%s\n", dif_version());

  /* ---- 1. gallery search from host buffers */
  enum { N = 20000, D = 128, Q = 64, K = 5 };
  float* rows = (float*)malloc(sizeof(float) * N * D);
  float* queries = (float*)malloc(sizeof(float) * Q * D);
  int64_t pick[Q];
  for (size_t i = 0; i < (size_t)N * D; ++i) rows[i] = unit_noise();
  for (int q = 0; q < Q; ++q) {
    pick[q] = (int64_t)((q * 311 + 17) % N);
    for (int d = 0; d < D; ++d) queries[q * D + d] = rows[pick[q] * D + d] + 0.3f * unit_noise();
  }
  dif_gallery_t* g = dif_gallery_create(0, N, D, DIF_METRIC_COSINE, DIF_PREC_TF32X3);
  REQUIRE(g != NULL, "%s", dif_last_error());
  CHECK(dif_gallery_add_host(g, rows, NULL, N));
  float scores[Q * K];
  int64_t ids[Q * K];
  CHECK(dif_gallery_search_host(g, queries, Q, K, scores, ids, NULL));
  for (int q = 0; q < Q; ++q) {
    REQUIRE(ids[q * K] == pick[q], "query %d: best id %lld, expected %lld", q, (long long)ids[q * K], (long long)pick[q]);
    const double want = cosine(queries + q * D, rows + pick[q] * D, D);
    REQUIRE(fabs(scores[q * K] - want) <= 1e-5, "query %d: score %.7f, cosine %.7f", q, scores[q * K], want);
    for (int j = 1; j < K; ++j) {
      REQUIRE(scores[q * K + j] <= scores[q * K + j - 1], "query %d: scores not descending at %d", q, j);
      const double w = cosine(queries + q * D, rows + ids[q * K + j] * D, D);
      REQUIRE(fabs(scores[q * K + j] - w) <= 1e-5, "query %d rank %d: score %.7f, cosine %.7f", q, j, scores[q * K + j], w);
    }
  }
  dif_gallery_destroy(g);

  /* ---- 2. pair distances */
  enum { P = 1000 };
  float dist[P];
  CHECK(dif_pair_distance_host(rows, rows + (size_t)P * D, P, D, DIF_METRIC_SQL2, NULL, dist));
  for (int i = 0; i < P; ++i) {
    double want = 0;
    for (int d = 0; d < D; ++d) {
      const double t = (double)rows[(size_t)i * D + d] - rows[(size_t)(P + i) * D + d];
      want += t * t;
    }
    REQUIRE(fabs(dist[i] - want) <= 1e-5 * want, "pair %d: squared distance %.6f, expected %.6f", i, dist[i], want);
  }
  CHECK(dif_pair_distance_host(rows, rows + (size_t)P * D, P, D, DIF_METRIC_COSINE, NULL, dist));
  for (int i = 0; i < P; ++i) {
    const double want = acos(cosine(rows + (size_t)i * D, rows + (size_t)(P + i) * D, D)) / 3.14159265358979323846;
    REQUIRE(fabs(dist[i] - want) <= 1e-5, "pair %d: arccos distance %.7f, expected %.7f", i, dist[i], want);
  }

  /* ---- 3. one batch-hard triplet step, cosine variant, alpha = 0.35 */
  enum { B = 72, PER = 4 };
  const float alpha = 0.35f;
  float emb[B * D], loss[B], stats[4], grad[B * D];
  int32_t labels[B], pos_idx[B], neg_idx[B];
  for (int i = 0; i < B; ++i) {
    labels[i] = i / PER;
    for (int d = 0; d < D; ++d) emb[i * D + d] = rows[(size_t)(1000 + labels[i]) * D + d] + 0.8f * unit_noise();
  }
  CHECK(dif_batch_hard_host(emb, labels, B, D, DIF_LOSS_BH_COSINE, alpha, loss, pos_idx, neg_idx, stats, NULL, grad,
                            DIF_PREC_TF32X3));
  for (int i = 0; i < B; ++i) {
    REQUIRE(pos_idx[i] >= 0 && pos_idx[i] < B && labels[pos_idx[i]] == labels[i], "anchor %d: positive %d", i, pos_idx[i]);
    REQUIRE(neg_idx[i] >= 0 && neg_idx[i] < B && labels[neg_idx[i]] != labels[i], "anchor %d: negative %d", i, neg_idx[i]);
    /* losses.py:42-51: hardest positive = the LEAST similar sample of the identity, hardest negative = the MOST similar other */
    double hp = 2, hn = -2;
    for (int j = 0; j < B; ++j) {
      const double c = cosine(emb + i * D, emb + j * D, D);
      if (labels[j] == labels[i]) {
        if (c < hp) hp = c;
      } else if (c > hn) {
        hn = c;
      }
    }
    const double want = hn - hp + alpha > 0 ? hn - hp + alpha : 0;
    REQUIRE(fabs(loss[i] - want) <= 1e-5, "anchor %d: loss %.7f, expected %.7f", i, loss[i], want);
  }
  double gsum = 0;
  for (int i = 0; i < B * D; ++i) gsum += fabs(grad[i]);
  REQUIRE(isfinite(gsum) && gsum > 0, "gradient sum %g", gsum);

  free(rows);
  free(queries);
  printf("host_search OK: %d queries x %d rows, %d pairs, one %d-sample triplet step\n", Q, N, P, B);
  return 0;
}
