/* TEST INFRASTRUCTURE: a CPU stand-in for the few entry points examples/host_search.c calls, so that the example's
 * own logic (argument order, buffer sizes, its double-precision checks and tolerances) can be exercised in the build
 * container, which has no GPU.  It includes include/dif_b200.h, so a signature that drifts from the header does not
 * compile.  The arithmetic comes from oracle/libdif_oracle.so (canonical fp32, what the CUDA library reproduces bit
 * for bit).  Never shipped, never loaded by the product: tests/test_host_example_cpu.py builds it into a temp dir. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "dif_b200.h"

void dif_or_normalize_rows(const float* x, int64_t n, int D, float* out);
float dif_or_canon_dot(const float* a, const float* b, int D);
void dif_or_gallery_search(const float* gallery, int64_t n_rows, int D, int metric, const float* queries, int n_queries,
                           int k, float* scores, int64_t* rows);
void dif_or_pair_distance(const float* e1, const float* e2, int64_t n, int D, int metric, float* out);

struct dif_gallery {
  float* rows;
  int64_t size, capacity;
  int D, metric;
};

int dif_init(int device) { return device == 0 ? DIF_OK : DIF_ERR_INVALID; }
const char* dif_last_error(void) { return "fake library"; }
const char* dif_version(void) { return "fake libdif_b200 (CPU stand-in for tests)"; }

dif_gallery_t* dif_gallery_create(int device, int64_t capacity_rows, int dim, int metric, int precision) {
  (void)device;
  (void)precision;
  dif_gallery_t* g = (dif_gallery_t*)calloc(1, sizeof(*g));
  g->rows = (float*)malloc(sizeof(float) * (size_t)capacity_rows * dim);
  g->capacity = capacity_rows;
  g->D = dim;
  g->metric = metric;
  return g;
}
void dif_gallery_destroy(dif_gallery_t* g) {
  free(g->rows);
  free(g);
}
int dif_gallery_add_host(dif_gallery_t* g, const float* rows_host, const int64_t* ids_host, int64_t n) {
  if (ids_host || g->size + n > g->capacity) return DIF_ERR_INVALID;
  float* dst = g->rows + (size_t)g->size * g->D;
  if (g->metric == DIF_METRIC_COSINE)
    dif_or_normalize_rows(rows_host, n, g->D, dst);
  else
    memcpy(dst, rows_host, sizeof(float) * (size_t)n * g->D);
  g->size += n;
  return DIF_OK;
}
int dif_gallery_search_host(dif_gallery_t* g, const float* queries_host, int n_queries, int k, float* scores_host,
                            int64_t* ids_host, int32_t* rows_host) {
  float* q = (float*)malloc(sizeof(float) * (size_t)n_queries * g->D);
  if (g->metric == DIF_METRIC_COSINE)
    dif_or_normalize_rows(queries_host, n_queries, g->D, q);
  else
    memcpy(q, queries_host, sizeof(float) * (size_t)n_queries * g->D);
  dif_or_gallery_search(g->rows, g->size, g->D, g->metric, q, n_queries, k, scores_host, ids_host);
  if (rows_host)
    for (int i = 0; i < n_queries * k; ++i) rows_host[i] = (int32_t)ids_host[i];
  free(q);
  return DIF_OK;
}
int dif_pair_distance_host(const float* e1_host, const float* e2_host, int64_t N, int D, int metric, const float* mean_host,
                           float* out_host) {
  if (mean_host) return DIF_ERR_INVALID;
  dif_or_pair_distance(e1_host, e2_host, N, D, metric, out_host);
  return DIF_OK;
}
/* common/losses.py:33-51 on the canonical cosine matrix; gradient: a nonzero placeholder (the example only checks that
 * it is finite and not all zero) */
int dif_batch_hard_host(const float* emb_host, const int32_t* labels_host, int B, int D, int variant, float alpha,
                        float* loss_host, int32_t* pos_idx_host, int32_t* neg_idx_host, float* stats_host,
                        const float* dloss_host, float* demb_host, int precision) {
  (void)precision;
  if (variant != DIF_LOSS_BH_COSINE || dloss_host) return DIF_ERR_INVALID;
  float* n = (float*)malloc(sizeof(float) * (size_t)B * D);
  dif_or_normalize_rows(emb_host, B, D, n);
  for (int i = 0; i < B; ++i) {
    float hp = 1.f, hn = -1.f;
    int pi = -1, ni = -1;
    for (int j = 0; j < B; ++j) {
      const float c = dif_or_canon_dot(n + (size_t)i * D, n + (size_t)j * D, D);
      if (labels_host[j] == labels_host[i]) {
        if (c < hp) hp = c, pi = j;
      } else if (c > hn) {
        hn = c, ni = j;
      }
    }
    loss_host[i] = fmaxf(hn - hp + alpha, 0.f);
    pos_idx_host[i] = pi;
    neg_idx_host[i] = ni;
  }
  for (int i = 0; i < 4; ++i) stats_host[i] = 0.f;
  if (demb_host)
    for (int i = 0; i < B * D; ++i) demb_host[i] = 1e-3f;
  free(n);
  return DIF_OK;
}
